// gpmm.cuh -- GPMM construction on the device (included by update.cu).
//
// Replaces GPMMTriangleMesh3D(reference, relativeTolerance).Gaussian / GaussianMixture (api/gpmm/GPMMHelper.scala:99-129)
// -> GPMM.construct (:39-54) -> scalismo LowRankGaussianProcess.approximateGPCholesky:
//   pivoted Cholesky L (3M x k) of the matrix-valued kernel on the reference points until
//   trace(residual) <= relTol * trace(K);  (V, d) = svd(L^T L);  basis = L V diag(d^-1/2),  variance = d.
// The kernels of these constructors are DiagonalKernel(scalar kernel, 3), i.e. K = Ks (x) I3 with
//   ks(x, y) = sum_q scaling_q exp(-|x - y|^2 / sigma_q^2)          [scalismo GaussianKernel: sigma^2, not 2 sigma^2]
// so the 3M x 3M factorisation is the scalar M x M one taken three times: a pivot (point p, dimension 0) leaves the
// residual diagonal of (p, 1), (p, 2) untouched and maximal, so the columns come in triples of the same point, and after
// c columns of triple k the trace is 3 T(k) - c (T(k) - T(k+1)) with T the scalar traces -- which also reproduces a rank
// that is not a multiple of three (tests/test_oracle_gpmm.py proves this against the literal 3M x 3M algorithm).
//   pivoted Cholesky   one column per step: argmax of the residual diagonal (lowest index on ties), the kernel column
//                      minus the projection on the previous columns, diagonal update; all scalars stay on the device,
//                      the host looks at the trace history once per batch of steps
//   KL basis           instead of forming L^T L and its eigenvectors: one-sided Jacobi (Hestenes) on the tall factor
//                      itself, L = U Sigma V^T, so U is the orthonormal basis and Sigma^2 the variances -- rotations of
//                      column pairs in round-robin order, one CTA per pair, until no pair is rotated
// The degenerate (triple) eigenvalues make the basis unique only up to a rotation inside each eigenspace; what is
// defined -- and tested -- is the rank, the variances and the covariance basis diag(variance) basis^T = L L^T.
#pragma once

namespace gingr {

struct GaussKernelDev {
  int n;
  double inv_sigma2[8];
  double scaling[8];
};

__device__ __forceinline__ double gpmm_kernel(const GaussKernelDev& kp, double dx, double dy, double dz) {
  const double d2 = dx * dx + dy * dy + dz * dz;
  double s = 0.0;
  for (int q = 0; q < kp.n; ++q) s += kp.scaling[q] * exp(-d2 * kp.inv_sigma2[q]);
  return s;
}

__global__ void gpmm_init_diag_kernel(int M, GaussKernelDev kp, double* __restrict__ d, uint8_t* __restrict__ done) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  d[i] = gpmm_kernel(kp, 0.0, 0.0, 0.0);
  done[i] = 0;
}

// per-block (max residual, lowest index) over the points not yet chosen, and the block's share of the residual trace
__global__ void __launch_bounds__(256) gpmm_argmax_partial_kernel(int M, const double* __restrict__ d, const uint8_t* __restrict__ done,
                                                                  double* __restrict__ pmax, int* __restrict__ pidx,
                                                                  double* __restrict__ psum) {
  __shared__ double smax[256], ssum[256];
  __shared__ int sidx[256];
  double best = -1.0, sum = 0.0;
  int bi = 0x7fffffff;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < M; i += gridDim.x * 256) {
    if (done[i]) continue;
    const double v = d[i];
    sum += v;
    if (v > best || (v == best && i < bi)) { best = v; bi = i; }
  }
  smax[threadIdx.x] = best; sidx[threadIdx.x] = bi; ssum[threadIdx.x] = sum;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      const double v = smax[threadIdx.x + o];
      const int j = sidx[threadIdx.x + o];
      if (v > smax[threadIdx.x] || (v == smax[threadIdx.x] && j < sidx[threadIdx.x])) { smax[threadIdx.x] = v; sidx[threadIdx.x] = j; }
      ssum[threadIdx.x] += ssum[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { pmax[blockIdx.x] = smax[0]; pidx[blockIdx.x] = sidx[0]; psum[blockIdx.x] = ssum[0]; }
}

// pivot of step k, its residual, and T(k) = residual trace before the step (fixed block order: deterministic)
__global__ void gpmm_argmax_final_kernel(int nb, int k, const double* __restrict__ pmax, const int* __restrict__ pidx,
                                         const double* __restrict__ psum, int* __restrict__ pivots, double* __restrict__ dpiv,
                                         double* __restrict__ trace) {
  if (threadIdx.x != 0) return;
  double best = -1.0, sum = 0.0;
  int bi = 0x7fffffff;
  for (int b = 0; b < nb; ++b) {
    sum += psum[b];
    if (pmax[b] > best || (pmax[b] == best && pidx[b] < bi)) { best = pmax[b]; bi = pidx[b]; }
  }
  pivots[k] = bi;
  dpiv[k] = best;
  trace[k] = sum;
}

// column k of the scalar factor (column-major Ls[k][M]): (ks(x_i, x_p) - sum_{j<k} Ls[j][i] Ls[j][p]) / sqrt(d_p) for the
// points not yet chosen, sqrt(d_p) at the pivot, 0 at earlier pivots; residual diagonal updated
__global__ void __launch_bounds__(256) gpmm_column_kernel(int M, int k, GaussKernelDev kp, const double* __restrict__ pts /*AoS*/,
                                                          const int* __restrict__ pivots, const double* __restrict__ dpiv,
                                                          double* __restrict__ Ls, double* __restrict__ d,
                                                          uint8_t* __restrict__ done) {
  extern __shared__ double lp[];   // Ls[j][p], j < k
  const int p = pivots[k];
  const double dp = dpiv[k];
  for (int j = threadIdx.x; j < k; j += 256) lp[j] = Ls[(size_t)j * M + p];
  __syncthreads();
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= M) return;
  double out = 0.0;
  if (dp > 0.0) {
    const double piv = sqrt(dp);
    if (i == p) {
      out = piv;
    } else if (!done[i]) {
      double v = gpmm_kernel(kp, pts[3 * i] - pts[3 * p], pts[3 * i + 1] - pts[3 * p + 1], pts[3 * i + 2] - pts[3 * p + 2]);
      for (int j = 0; j < k; ++j) v -= Ls[(size_t)j * M + i] * lp[j];
      out = v / piv;
      d[i] -= out * out;
    }
  }
  Ls[(size_t)k * M + i] = out;
  if (i == p) done[i] = 1;
}

// ---- one-sided Jacobi on the columns of A (column-major [n][M]) --------------------------------------------------
// round-robin pairing (circle method) of n_even players, player n_even - 1 fixed
__device__ __forceinline__ void jacobi_pair(int n_even, int round, int m, int& a, int& b) {
  const int nm1 = n_even - 1;
  if (m == 0) { a = nm1; b = round % nm1; }
  else { a = (round + m) % nm1; b = (round - m + nm1) % nm1; }
  if (a > b) { const int t = a; a = b; b = t; }
}

__global__ void __launch_bounds__(256) gpmm_jacobi_round_kernel(int M, int n, int n_even, int round, double tol,
                                                                double* __restrict__ A, int* __restrict__ rotated) {
  __shared__ double red[3][256];
  __shared__ double cs[2];
  int a, b;
  jacobi_pair(n_even, round, blockIdx.x, a, b);
  if (b >= n) return;   // the dummy player of an odd n
  double* ca = A + (size_t)a * M;
  double* cb = A + (size_t)b * M;
  double saa = 0.0, sbb = 0.0, sab = 0.0;
  for (int i = threadIdx.x; i < M; i += 256) {
    const double x = ca[i], y = cb[i];
    saa = fma(x, x, saa); sbb = fma(y, y, sbb); sab = fma(x, y, sab);
  }
  red[0][threadIdx.x] = saa; red[1][threadIdx.x] = sbb; red[2][threadIdx.x] = sab;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o)
      for (int q = 0; q < 3; ++q) red[q][threadIdx.x] += red[q][threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double alpha = red[0][0], beta = red[1][0], gamma = red[2][0];
    double c = 1.0, s = 0.0;
    if (fabs(gamma) > tol * sqrt(alpha * beta) && gamma != 0.0) {
      const double zeta = (beta - alpha) / (2.0 * gamma);
      const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
      c = 1.0 / sqrt(1.0 + t * t);
      s = c * t;
      atomicExch(rotated, 1);
    }
    cs[0] = c; cs[1] = s;
  }
  __syncthreads();
  const double c = cs[0], s = cs[1];
  if (s == 0.0) return;
  for (int i = threadIdx.x; i < M; i += 256) {
    const double x = ca[i], y = cb[i];
    ca[i] = c * x - s * y;
    cb[i] = s * x + c * y;
  }
}

__global__ void __launch_bounds__(256) gpmm_colnorm_kernel(int M, const double* __restrict__ A, double* __restrict__ norm2) {
  __shared__ double red[256];
  const double* c = A + (size_t)blockIdx.x * M;
  double s = 0.0;
  for (int i = threadIdx.x; i < M; i += 256) s = fma(c[i], c[i], s);
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) norm2[blockIdx.x] = red[0];
}

// model column c = (source set, scalar column j, dimension d):  phi[3 i + d][c] = A_set[j][i] / sqrt(norm2_set[j])
struct GpmmColumn { int set, j, d; };
__global__ void gpmm_assemble_kernel(int M, int r, int rp, const GpmmColumn* __restrict__ cols, const double* __restrict__ A0,
                                     const double* __restrict__ n0, const double* __restrict__ A1, const double* __restrict__ n1,
                                     double* __restrict__ phi) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int c = blockIdx.y;
  if (i >= M || c >= r) return;
  const GpmmColumn col = cols[c];
  const double* A = col.set == 0 ? A0 : A1;
  const double* nn = col.set == 0 ? n0 : n1;
  phi[((size_t)3 * i + col.d) * rp + c] = A[(size_t)col.j * M + i] / sqrt(nn[col.j]);
}

}  // namespace gingr

using namespace gingr;

// Hestenes sweeps until no pair is rotated; returns the squared column norms (host)
static int32_t gpmm_jacobi(gingr_ctx* ctx, int M, int n, double* d_A, int* d_flag, double* d_norm2, std::vector<double>* norm2,
                           int* sweeps_out) {
  cudaStream_t st = ctx->stream;
  const int n_even = (n + 1) & ~1;
  // a pair counts as orthogonal below the rounding level of its own dot product (length-M sums in FP64)
  const double tol = std::max(1e-14, 8.0 * 2.220446049250313e-16 * sqrt((double)M));
  int sweeps = 0;
  if (n >= 2) {
    for (; sweeps < 40; ++sweeps) {
      GINGR_CUDA_TRY(ctx, cudaMemsetAsync(d_flag, 0, sizeof(int), st));
      for (int round = 0; round < n_even - 1; ++round) {
        gpmm_jacobi_round_kernel<<<n_even / 2, 256, 0, st>>>(M, n, n_even, round, tol, d_A, d_flag);
        GINGR_LAUNCHED(ctx);
      }
      int h = 0;
      GINGR_CUDA_TRY(ctx, cudaMemcpyAsync(&h, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
      GINGR_CUDA_TRY(ctx, cudaStreamSynchronize(st));
      if (!h) { ++sweeps; break; }
    }
  }
  gpmm_colnorm_kernel<<<n, 256, 0, st>>>(M, d_A, d_norm2);
  GINGR_LAUNCHED(ctx);
  norm2->resize(n);
  GINGR_CUDA_TRY(ctx, cudaMemcpyAsync(norm2->data(), d_norm2, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
  GINGR_CUDA_TRY(ctx, cudaStreamSynchronize(st));
  if (sweeps_out) *sweeps_out = sweeps;
  return GINGR_OK;
}

extern "C" {

int32_t gingr_gpmm_gaussian_mixture(gingr_ctx* ctx, int32_t M, const double* ref_pts, const int32_t* tri, int32_t T,
                                    int32_t n_kernels, const double* sigma, const double* scaling, double rel_tol,
                                    int32_t max_rank, gingr_model** out, int32_t* rank_out) {
  if (!ctx || !ref_pts || !out || M <= 0 || T < 0 || (T > 0 && !tri) || n_kernels <= 0 || n_kernels > 8 || !sigma || !scaling ||
      !(rel_tol >= 0.0) || max_rank < 0)
    return gingr_fail(ctx, GINGR_ERR_ARG, "gingr_gpmm_gaussian_mixture: bad argument (1..8 kernels, relTol >= 0)");
  if (ctx->nranks != 1) return gingr_fail(ctx, GINGR_ERR_UNSUPPORTED, "gingr_gpmm_gaussian_mixture: single-GPU entry point (build before sharding)");
  GaussKernelDev kp;
  kp.n = n_kernels;
  for (int q = 0; q < n_kernels; ++q) {
    if (!(sigma[q] > 0.0) || !(scaling[q] > 0.0)) return gingr_fail(ctx, GINGR_ERR_ARG, "gingr_gpmm_gaussian_mixture: sigma and scaling must be positive");
    kp.inv_sigma2[q] = 1.0 / (sigma[q] * sigma[q]);
    kp.scaling[q] = scaling[q];
  }
  GINGR_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const long long full_cap = max_rank > 0 ? std::min<long long>(max_rank, 3LL * M) : 3LL * M;
  // scalar columns that may be needed (6000: the previous columns' pivot row is staged in 48 KB of shared memory)
  const int kcap = (int)std::min<long long>(std::min<long long>((full_cap + 2) / 3, M), 6000);
  const int nb = std::min(128, ceil_div(M, 256));
  DevBuf<double> pts, d, Ls, Lt, pmax, psum, dpiv, trace, norm2a, norm2b;
  DevBuf<int> pidx, pivots, flag;
  DevBuf<uint8_t> done;
  DevBuf<GpmmColumn> dcols;
  gingr_model* m = nullptr;
  int32_t rc = GINGR_OK;
  auto A = [&](cudaError_t e) { if (e != cudaSuccess && rc == GINGR_OK) { gingr_set_error(ctx, cudaGetErrorString(e)); rc = GINGR_ERR_CUDA; } };
  auto cleanup = [&]() {
    pts.release(); d.release(); Ls.release(); Lt.release(); pmax.release(); psum.release(); dpiv.release(); trace.release();
    norm2a.release(); norm2b.release(); pidx.release(); pivots.release(); flag.release(); done.release(); dcols.release();
  };
  // the factor grows in chunks of columns so that a small rank does not allocate M x kcap
  int kalloc = std::min(kcap, 256);
  A(pts.alloc((size_t)3 * M)); A(d.alloc(M)); A(done.alloc(M)); A(Ls.alloc((size_t)kalloc * M));
  A(pmax.alloc(nb)); A(psum.alloc(nb)); A(pidx.alloc(nb)); A(pivots.alloc(kcap + 1)); A(dpiv.alloc(kcap + 1));
  A(trace.alloc(kcap + 2)); A(flag.alloc(4));
  if (rc != GINGR_OK) { cleanup(); return rc; }
  A(cudaMemcpyAsync(pts.p, ref_pts, sizeof(double) * 3 * (size_t)M, cudaMemcpyHostToDevice, st));
  gpmm_init_diag_kernel<<<ceil_div(M, 256), 256, 0, st>>>(M, kp, d.p, done.p);
  GINGR_LAUNCHED(ctx);
  // ---- pivoted Cholesky: batches of steps, the trace history decides where the reference's loop would have stopped
  std::vector<double> htrace;
  int k = 0;            // scalar steps done
  int rank = -1;        // full rank 3 kk + c once known
  double tol = 0.0;
  while (rc == GINGR_OK && rank < 0) {
    const int batch = std::min(32, kcap - k);
    if (k + batch > kalloc) {   // grow the factor
      const int knew = std::min(kcap, std::max(kalloc * 2, k + batch));
      DevBuf<double> bigger;
      A(bigger.alloc((size_t)knew * M));
      if (rc == GINGR_OK) A(cudaMemcpyAsync(bigger.p, Ls.p, sizeof(double) * (size_t)k * M, cudaMemcpyDeviceToDevice, st));
      if (rc == GINGR_OK) A(cudaStreamSynchronize(st));
      if (rc != GINGR_OK) { bigger.release(); break; }
      Ls.release();
      Ls = bigger;
      kalloc = knew;
    }
    for (int s = 0; s < batch; ++s) {
      gpmm_argmax_partial_kernel<<<nb, 256, 0, st>>>(M, d.p, done.p, pmax.p, pidx.p, psum.p);
      gpmm_argmax_final_kernel<<<1, 32, 0, st>>>(nb, k + s, pmax.p, pidx.p, psum.p, pivots.p, dpiv.p, trace.p);
      gpmm_column_kernel<<<ceil_div(M, 256), 256, sizeof(double) * (size_t)std::max(k + s, 1), st>>>(M, k + s, kp, pts.p, pivots.p,
                                                                                                    dpiv.p, Ls.p, d.p, done.p);
      ctx->launches += 3;
    }
    k += batch;
    // T(k) after the batch
    gpmm_argmax_partial_kernel<<<nb, 256, 0, st>>>(M, d.p, done.p, pmax.p, pidx.p, psum.p);
    gpmm_argmax_final_kernel<<<1, 32, 0, st>>>(nb, k, pmax.p, pidx.p, psum.p, pivots.p, dpiv.p, trace.p);
    ctx->launches += 2;
    htrace.resize(k + 1);
    A(cudaMemcpyAsync(htrace.data(), trace.p, sizeof(double) * (k + 1), cudaMemcpyDeviceToHost, st));
    A(cudaStreamSynchronize(st));
    A(cudaGetLastError());
    if (rc != GINGR_OK) break;
    tol = rel_tol * 3.0 * htrace[0];
    // the reference's loop: while (cols < n && trace > tolerance) add a column
    for (int kk = 0; kk < k && rank < 0; ++kk)
      for (int c = 0; c < 3; ++c) {
        const double tr_before = 3.0 * htrace[kk] - c * (htrace[kk] - htrace[kk + 1]);
        const long long cols = 3LL * kk + c;
        if (!(tr_before > tol) || cols >= full_cap) { rank = (int)cols; break; }
      }
    if (rank < 0 && (k >= kcap || !(3.0 * htrace[k] > tol))) rank = (int)std::min<long long>(3LL * k, full_cap);
    if (rank < 0 && (long long)3 * k >= full_cap) rank = (int)full_cap;
  }
  if (rc == GINGR_OK && rank <= 0) rc = gingr_fail(ctx, GINGR_ERR_ARG, "gingr_gpmm_gaussian_mixture: the tolerance leaves an empty model");
  if (rc != GINGR_OK) { cleanup(); return rc; }
  // dims < c use k1 scalar columns, dims >= c use k1 - 1 (c == 0: all dims use rank / 3)
  const int cpart = rank % 3;
  const int k_hi = (rank + 2) / 3, k_lo = rank / 3;
  // ---- KL basis: one-sided Jacobi on the factor (a second, truncated copy when the last triple is incomplete)
  std::vector<double> n_hi, n_lo;
  int sweeps = 0;
  A(norm2a.alloc(std::max(k_hi, 1))); A(norm2b.alloc(std::max(k_lo, 1)));
  if (cpart != 0 && k_lo > 0) {
    A(Lt.alloc((size_t)k_lo * M));
    if (rc == GINGR_OK) A(cudaMemcpyAsync(Lt.p, Ls.p, sizeof(double) * (size_t)k_lo * M, cudaMemcpyDeviceToDevice, st));
  }
  if (rc == GINGR_OK) rc = gpmm_jacobi(ctx, M, k_hi, Ls.p, flag.p, norm2a.p, &n_hi, &sweeps);
  if (rc == GINGR_OK && cpart != 0 && k_lo > 0) rc = gpmm_jacobi(ctx, M, k_lo, Lt.p, flag.p, norm2b.p, &n_lo, nullptr);
  if (rc != GINGR_OK) { cleanup(); return rc; }
  // ---- columns of the model, sorted by variance (descending; ties: dimension, then scalar column)
  std::vector<GpmmColumn> cols;
  std::vector<double> lam;
  for (int dd = 0; dd < 3; ++dd) {
    const bool hi = cpart == 0 || dd < cpart;
    const int kd = hi ? k_hi : k_lo;
    for (int j = 0; j < kd; ++j) {
      cols.push_back(GpmmColumn{hi ? 0 : 1, j, dd});
      lam.push_back(hi ? n_hi[j] : n_lo[j]);
    }
  }
  std::vector<int> order(cols.size());
  for (size_t q = 0; q < order.size(); ++q) order[q] = (int)q;
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return lam[x] > lam[y]; });
  const int r = (int)cols.size();
  std::vector<GpmmColumn> scols(r);
  std::vector<double> var(r);
  for (int q = 0; q < r; ++q) { scols[q] = cols[order[q]]; var[q] = lam[order[q]]; }
  // ---- the model handle
  m = new gingr_model();
  m->ctx = ctx;
  m->M = M; m->r = r; m->rp = (r + 7) / 8 * 8; m->m0 = 0; m->Ml = M;
  A(m->ref.alloc((size_t)3 * M)); A(m->mean.alloc((size_t)3 * M)); A(m->sqrt_lambda.alloc(m->rp));
  A(m->phi.alloc((size_t)3 * M * m->rp)); A(dcols.alloc(r));
  if (rc == GINGR_OK) {
    std::vector<double> sl(m->rp, 0.0);
    for (int q = 0; q < r; ++q) sl[q] = sqrt(var[q]);
    A(cudaMemcpyAsync(m->ref.p, pts.p, sizeof(double) * 3 * (size_t)M, cudaMemcpyDeviceToDevice, st));
    A(cudaMemsetAsync(m->mean.p, 0, sizeof(double) * 3 * (size_t)M, st));
    A(cudaMemsetAsync(m->phi.p, 0, sizeof(double) * (size_t)3 * M * m->rp, st));
    A(cudaMemcpyAsync(m->sqrt_lambda.p, sl.data(), sizeof(double) * m->rp, cudaMemcpyHostToDevice, st));
    A(cudaMemcpyAsync(dcols.p, scols.data(), sizeof(GpmmColumn) * r, cudaMemcpyHostToDevice, st));
    gpmm_assemble_kernel<<<dim3(ceil_div(M, 256), r), 256, 0, st>>>(M, r, m->rp, dcols.p, Ls.p, norm2a.p, Lt.p ? Lt.p : Ls.p,
                                                                    norm2b.p, m->phi.p);
    GINGR_LAUNCHED(ctx);
    A(cudaGetLastError());
    A(cudaStreamSynchronize(st));   // sl / scols are host vectors
  }
  cleanup();
  if (rc == GINGR_OK) rc = model_upload_topology(ctx, m, tri, T);
  if (rc == GINGR_OK) rc = model_build_constants(ctx, m);
  if (rc != GINGR_OK) { gingr_model_destroy(m); return rc; }
  *out = m;
  if (rank_out) *rank_out = r;
  (void)sweeps;
  return GINGR_OK;
}

// The model as scalismo stores it: meanVector [3M], basisMatrix column-major [3M x r] (leading dimension
// ld_basis >= 3M), variance [r].  Any output may be NULL.
int32_t gingr_model_download(gingr_ctx* ctx, const gingr_model* model, int32_t* M_out, int32_t* r_out, double* ref_pts,
                             double* mean, double* basis, int64_t ld_basis, double* variance) {
  if (!ctx || !model) return gingr_fail(ctx, GINGR_ERR_ARG, "gingr_model_download: bad argument");
  if (ctx->nranks != 1) return gingr_fail(ctx, GINGR_ERR_UNSUPPORTED, "gingr_model_download: single-GPU entry point");
  const int M = model->M, r = model->r, rp = model->rp;
  if (M_out) *M_out = M;
  if (r_out) *r_out = r;
  if (basis && ld_basis < 3 * (int64_t)M) return gingr_fail(ctx, GINGR_ERR_ARG, "gingr_model_download: ld_basis < 3 M");
  GINGR_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  if (ref_pts) GINGR_CUDA_TRY(ctx, cudaMemcpyAsync(ref_pts, model->ref.p, sizeof(double) * 3 * (size_t)M, cudaMemcpyDeviceToHost, st));
  if (mean) GINGR_CUDA_TRY(ctx, cudaMemcpyAsync(mean, model->mean.p, sizeof(double) * 3 * (size_t)M, cudaMemcpyDeviceToHost, st));
  std::vector<double> sl, rows;
  if (variance) {
    sl.resize(rp);
    GINGR_CUDA_TRY(ctx, cudaMemcpyAsync(sl.data(), model->sqrt_lambda.p, sizeof(double) * rp, cudaMemcpyDeviceToHost, st));
  }
  if (basis) {
    rows.resize((size_t)3 * M * rp);
    GINGR_CUDA_TRY(ctx, cudaMemcpyAsync(rows.data(), model->phi.p, sizeof(double) * rows.size(), cudaMemcpyDeviceToHost, st));
  }
  GINGR_CUDA_TRY(ctx, cudaStreamSynchronize(st));
  if (variance)
    for (int a = 0; a < r; ++a) variance[a] = sl[a] * sl[a];
  if (basis)
    for (int a = 0; a < r; ++a)
      for (size_t k = 0; k < (size_t)3 * M; ++k) basis[(size_t)a * ld_basis + k] = rows[k * rp + a];
  return GINGR_OK;
}

}  // extern "C"
