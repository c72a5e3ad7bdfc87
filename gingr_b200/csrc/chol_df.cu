// chol_df.cu -- K3b: the r x r Cholesky factorisation as ONE persistent data-flow kernel (sm_100a).
//
// Replaces `Minv = breeze.linalg.pinv(M)` of scalismo's regression (call sites
// GingrAlgorithm.scala:300, :215, :236) together with chol.cu's back substitution.  The factorisation is the one
// replicated, latency-bound step of a multi-GPU iteration, so what counts is its CRITICAL PATH, not its flops
// (r^3/3 = 2.7 GFLOP at r = 2000 is 0.1 ms of DMMA): the previous form (one panel kernel + one or two trailing-update
// kernels per 64-column step, 32 dependent steps of ~47 us) is replaced by a tile data-flow:
//
//   * the lower triangle is cut into 64 x 64 tiles; tiles are handed out in column-major order by an atomic counter
//     to a persistent grid (one CTA per SM); a CTA that owns tile (i, j) keeps its accumulator in registers,
//     subtracts X_ik X_jk^T (DMMA.8x8x4) for every k < j as soon as the two operand tiles are flagged final
//     (left-looking, operands double-buffered through cp.async), then
//   * diagonal tile: factorises the 64 x 64 block inside the CTA (two in-register 32 x 32 warp factorisations that
//     eliminate TWO columns per dependent rsqrt, the off-diagonal block by substitution, everything else as small
//     products), also forms L_jj^-1, publishes both and raises the tile's flag;
//   * off-diagonal tile: waits for L_jj^-1 and finishes with X_ij = A_ij L_jj^-T as a (triangular) DMMA product --
//     no substitution outside the diagonal tiles.
//   Because a CTA only ever waits for tiles with a smaller ticket, and every ticket that was drawn belongs to a
//   running CTA, the spin waits cannot deadlock whatever else occupies the GPU.  No launch boundary, no grid-wide
//   barrier: between two diagonal factorisations the critical path is  flag -> one tile product -> flag.
//
// Rows below the square part (right-hand sides stored as extra rows) ride along, which performs the forward
// substitution L^-1 b inside the factorisation (as chol.cu's kernels did).
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "posterior.cuh"

namespace gingr {

namespace {

constexpr int TB = 64;                 // tile edge
constexpr int TP = TB + 4;             // shared pitch, == 4 (mod 16): conflict-free DMMA fragment loads
constexpr int DF_THREADS = 128;
constexpr int TILE_DOUBLES = TB * TP;  // 4352
// shared memory: 5 tile buffers (2 stages x 2 operands + the tile's own input), diag(L) and the sub-diagonal entries
// L[2s+1][2s] of the in-CTA factorisation
constexpr size_t DF_SMEM = (size_t)(5 * TILE_DOUBLES + TB + 32) * sizeof(double) + 64;

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

#ifdef DF_TIMING
// tuning build only (tools/build_variant.sh dft chol_df.cu -DDF_TIMING): per tile 8 globaltimer stamps (ns) and, for
// diagonal tiles, 8 clock64 phase stamps of the in-CTA factorisation; read with gingr_debug_chol_df_timing
__device__ unsigned long long df_stamps[64 * 64 * 16];
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define DF_STAMP(tile, k) do { if (threadIdx.x == 0 && (tile) < 64 * 64) df_stamps[(tile) * 16 + (k)] = gtime(); } while (0)
#define DF_CLOCK(tile, k) do { if (threadIdx.x == 0 && (tile) < 64 * 64) df_stamps[(tile) * 16 + 8 + (k)] = (unsigned long long)clock64(); } while (0)
#else
#define DF_STAMP(tile, k) ((void)0)
#define DF_CLOCK(tile, k) ((void)0)
#endif

struct DfParams {
  double* A;
  int ld, n, nrows, nb, nbr, ntiles;
  int* sync;      // [0] ticket counter, [1] finished CTAs, [2 ...] flags[nbr][nb]
  double* linv;   // [nb][64][64] inverses of the diagonal blocks
  int* info;
};

// 64 rows x 64 columns of A starting at (row0, col0) -> s (pitch TP); rows >= rows_valid and columns >= ld are zero-filled
__device__ __forceinline__ void load_tile_async(double* s, const double* __restrict__ A, int ld, int row0, int col0,
                                                int rows_valid, int tid) {
#pragma unroll 4
  for (int e = tid; e < TB * (TB / 2); e += DF_THREADS) {
    const int i = e >> 5, c = (e & 31) * 2;
    double* dst = s + i * TP + c;
    if (i < rows_valid && col0 + c < ld) cp_async16(dst, A + (size_t)(row0 + i) * ld + col0 + c);   // never past the row pitch
    else *reinterpret_cast<double2*>(dst) = make_double2(0.0, 0.0);
  }
}

// acc += I K^T over the 64 columns of the two shared tiles (warp tile 32 x 32: acc[i][j] = rows wm*32+i*8+g, cols wn*32+j*8+2t)
__device__ __forceinline__ void mma_tile(double (&acc)[4][4][2], const double* __restrict__ sI,
                                         const double* __restrict__ sK, int wm, int wn, int g, int t) {
#pragma unroll 4
  for (int c4 = 0; c4 < TB / 4; ++c4) {
    double af[4], bf[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) af[i] = sI[(wm * 32 + i * 8 + g) * TP + c4 * 4 + t];
#pragma unroll
    for (int j = 0; j < 4; ++j) bf[j] = sK[(wn * 32 + j * 8 + g) * TP + c4 * 4 + t];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
  }
}

// X = V Z^T with Z lower triangular (Z = L^-1): column c of X only needs k <= c, so a fragment column block jj stops at
// k4 < wn*8 + jj*2 + 2
__device__ __forceinline__ void mma_tile_lower(double (&acc)[4][4][2], const double* __restrict__ sI,
                                               const double* __restrict__ sK, int wm, int wn, int g, int t) {
#pragma unroll 2
  for (int c4 = 0; c4 < TB / 4; ++c4) {
    if (c4 >= wn * 8 + 8) break;
    double af[4], bf[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) af[i] = sI[(wm * 32 + i * 8 + g) * TP + c4 * 4 + t];
#pragma unroll
    for (int j = 0; j < 4; ++j) bf[j] = sK[(wn * 32 + j * 8 + g) * TP + c4 * 4 + t];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (c4 < wn * 8 + j * 2 + 2) {
#pragma unroll
        for (int i = 0; i < 4; ++i) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// In-CTA factorisation of a 64 x 64 block, L AND L^-1 in one pass (all 128 threads).
//
// What bounds it (tools/lat_bench.cu on a B200): a DFMA issues once per ~6 cycles from a single warp (8 cycles dependent),
// a 64-bit shuffle costs 26, a shared-memory exchange 35, the branch-free rsqrt 49.  So (1) the trailing updates must be
// spread over all four warps, and (2) the dependent chain pivot -> rsqrt -> scale -> exchange -> update is walked as few
// times as possible.  Layout: lane holds rows `lane` and `lane + 32`; warp g holds the column PAIRS 4 lp + g (lp < 8), i.e.
// 32 matrix entries per thread in registers.  Step s eliminates the columns j = 2 s and j + 1 at once: with p = a_jj,
// q = a_{j+1,j}, t = a_{j+1,j+1} the second pivot is det / p, det = p t - q^2, so rsqrt(p) and rsqrt(det) are independent
// (one chain per TWO columns; det loses what t - (q rs)^2 loses).  The scaled columns go through shared memory (G2, one
// double2 per row and step), one CTA barrier per step, and every thread updates its own 2 x 16 entries.
// L^-1 rides along for free: a row that has been eliminated (row j after its pivot step) continues as row j of the identity
// appended below the block -- rows below the square part receive X = I L^-T under exactly the same column operations, and
// row m of the identity is zero before step m/2 -- so every thread always carries two ACTIVE rows, and at the end
//   G2[r][s] = (L[r][2s], L[r][2s+1])  for r > 2s + 1,      G2[r][s] = (Z[2s][r], Z[2s+1][r])  for r <= 2s + 1   (Z = L^-1)
// with L[2s+1][2s] in sSub[s] and the diagonal of L in sDiag.
// The chain is executed by all warps on their own registers (no divergent block around it, so that ptxas can interleave it
// with the updates); only the owner warp of the pair stores.
// ---------------------------------------------------------------------------------------------
constexpr int G2P = 33;   // double2 pitch of G2: 132 words == 4 (mod 32), conflict-free LDS.128 down a column

__device__ __forceinline__ double rsq_nr(double d) {   // MUFU.RSQ64H seed + one cubic step: rsqrt() without its slow-path branch
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(d));
  const double e = fma(-d * y0, y0, 1.0);
  const double h = fma(e, 0.375, 0.5);
  return fma(y0 * e, h, y0);
}

// sT: the block (lower triangle, pitch TP, identity-padded beyond the valid part).  On return G2 / sDiag / sSub hold L and
// Z = L^-1 as described above (read them with potrf_L / potrf_Z).  Returns true if a pivot was not positive and finite.
__device__ __forceinline__ double potrf_L(const double2* __restrict__ G2, const double* __restrict__ sDiag,
                                          const double* __restrict__ sSub, int r, int c) {   // r >= c
  if (r == c) return sDiag[c];
  if (r == c + 1 && !(c & 1)) return sSub[c >> 1];
  const double2 v = G2[r * G2P + (c >> 1)];
  return (c & 1) ? v.y : v.x;
}
__device__ __forceinline__ double potrf_Z(const double2* __restrict__ G2, int c, int k) {   // Z[c][k], zero for k > c
  if (k > c) return 0.0;
  const double2 v = G2[k * G2P + (c >> 1)];
  return (c & 1) ? v.y : v.x;
}
__device__ __forceinline__ bool potrf64(const double* __restrict__ sT, double2* __restrict__ G2, double* __restrict__ sDiag,
                                        double* __restrict__ sSub, int tid, int tile_id) {
  const int lane = tid & 31, g = tid >> 5;
  (void)tile_id;
  DF_CLOCK(tile_id, 0);
  double a0[16], a1[16];   // rows lane / lane + 32, columns 2 (4 lp + g) + e at index 2 lp + e
#pragma unroll
  for (int lp = 0; lp < 8; ++lp)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int col = 2 * (4 * lp + g) + e;
      a0[2 * lp + e] = (col <= lane) ? sT[lane * TP + col] : 0.0;
      a1[2 * lp + e] = (col <= lane + 32) ? sT[(lane + 32) * TP + col] : 0.0;
    }
  bool bad = false;
#pragma unroll
  for (int s = 0; s < 32; ++s) {
    const int lp0 = s >> 2, go = s & 3;        // local index and owner warp of the pair
    const int j = 2 * s, hf = j >> 5, pl = j & 31;
    // ---- the chain: pivot block -> two independent rsqrt -> scaled columns of both rows ----------------------
    const double v0 = hf ? a1[2 * lp0] : a0[2 * lp0];
    const double v1 = hf ? a1[2 * lp0 + 1] : a0[2 * lp0 + 1];
    const double p = __shfl_sync(0xffffffffu, v0, pl);
    const double q = __shfl_sync(0xffffffffu, v0, pl + 1);
    const double t = __shfl_sync(0xffffffffu, v1, pl + 1);
    const double det = fma(p, t, -q * q);
    const double rs1 = rsq_nr(p), rsd = rsq_nr(det);
    const double sp = p * rs1;         // L[j][j]
    const double l21 = q * rs1;        // L[j+1][j]
    const double i22 = rsd * sp;       // 1 / L[j+1][j+1]
    double f0a = a0[2 * lp0] * rs1, f0b = a1[2 * lp0] * rs1;
    double f1a = fma(-f0a, l21, a0[2 * lp0 + 1]) * i22, f1b = fma(-f0b, l21, a1[2 * lp0 + 1]) * i22;
    // rows j and j + 1 turn into rows of the appended identity: their factors are those of e_j and e_{j+1}
    const bool sw0 = (lane == pl), sw1 = (lane == pl + 1);
    if (hf == 0) {
      if (sw0) { f0a = rs1; f1a = -rs1 * l21 * i22; }
      if (sw1) { f0a = 0.0; f1a = i22; }
    } else {
      if (sw0) { f0b = rs1; f1b = -rs1 * l21 * i22; }
      if (sw1) { f0b = 0.0; f1b = i22; }
    }
    if (g == go) {
      if (!(p > 0.0) || !(p < INFINITY) || !(det > 0.0) || !(det < INFINITY)) bad = true;
      G2[lane * G2P + s] = make_double2(f0a, f1a);
      G2[(lane + 32) * G2P + s] = make_double2(f0b, f1b);
      if (lane == 0) {
        sDiag[j] = sp;
        sDiag[j + 1] = fma(-l21, l21, t) * i22;
        sSub[s] = l21;
      }
    }
    __syncthreads();
    if (s == 31) break;
    // ---- trailing update of the own pairs > s (rows that just switched restart from zero) ------------------
    const double2 Fa = G2[lane * G2P + s], Fb = G2[(lane + 32) * G2P + s];
    const bool za = (hf == 0) && (sw0 || sw1), zb = (hf == 1) && (sw0 || sw1);
#pragma unroll
    for (int lp = lp0; lp < 8; ++lp) {
      if (lp > lp0 || g > go) {
        const int P = 4 * lp + g;
        const double2 c0 = G2[(2 * P) * G2P + s], c1 = G2[(2 * P + 1) * G2P + s];
        const double x0 = za ? 0.0 : a0[2 * lp], x1 = za ? 0.0 : a0[2 * lp + 1];
        const double y0 = zb ? 0.0 : a1[2 * lp], y1 = zb ? 0.0 : a1[2 * lp + 1];
        a0[2 * lp] = fma(-Fa.y, c0.y, fma(-Fa.x, c0.x, x0));
        a0[2 * lp + 1] = fma(-Fa.y, c1.y, fma(-Fa.x, c1.x, x1));
        a1[2 * lp] = fma(-Fb.y, c0.y, fma(-Fb.x, c0.x, y0));
        a1[2 * lp + 1] = fma(-Fb.y, c1.y, fma(-Fb.x, c1.x, y1));
      }
    }
  }
  DF_CLOCK(tile_id, 1);
  return __syncthreads_or(bad ? 1 : 0) != 0;
}


__global__ void __launch_bounds__(DF_THREADS, 1) chol_df_kernel(DfParams P) {
  extern __shared__ __align__(16) double dsm[];
  double* buf0 = dsm;                       // stage 0, row operand   | potrf: Z = L^-1        | trsm: Z
  double* buf1 = dsm + TILE_DOUBLES;        // stage 0, column operand| potrf: extra rows
  double* buf2 = dsm + 2 * TILE_DOUBLES;    // stage 1, row operand   | potrf: column exchange G2
  double* buf3 = dsm + 3 * TILE_DOUBLES;    // stage 1, column operand
  double* sA0 = dsm + 4 * TILE_DOUBLES;     // the tile's own input, later its updated value / L
  double* sDiag = dsm + 5 * TILE_DOUBLES;   // [64]
  double* sSub = sDiag + TB;                // [32]
  __shared__ int s_tile, s_pre[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 1, wn = warp & 1, g = lane >> 2, t = lane & 3;
  int* flags = P.sync + 2;
  const double* __restrict__ A = P.A;
  const int ld = P.ld;

  for (;;) {
    __syncthreads();   // s_tile / shared buffers of the previous tile are free
    if (tid == 0) s_tile = atomicAdd(&P.sync[0], 1);
    __syncthreads();
    int tt = s_tile;
    if (tt >= P.ntiles) break;
    int j = 0;
    while (tt >= P.nbr - j) { tt -= P.nbr - j; ++j; }
    const int i = j + tt;
    const int i0 = i * TB, j0 = j * TB;
    const int jb = min(TB, P.n - j0);              // valid columns of this block column
    const int rcnt = min(TB, P.nrows - i0);        // valid rows of this tile
    const bool diag = (i == j);
    const int tile_id = i * P.nb + j;
    (void)tile_id;
    DF_STAMP(tile_id, 0);

    // the tile's own input (its group completes long before it is needed)
    load_tile_async(sA0, A, ld, i0, j0, rcnt, tid);
    cp_async_commit();

    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

    // ---- left-looking updates: S = sum_k X_ik X_jk^T ------------------------------------------------------
    if (j > 0) {
      if (tid == 0) {
        while (ld_acquire(&flags[i * P.nb + 0]) == 0) {}
        if (!diag) while (ld_acquire(&flags[j * P.nb + 0]) == 0) {}
      }
      __syncthreads();
      load_tile_async(buf0, A, ld, i0, 0, rcnt, tid);
      if (!diag) load_tile_async(buf1, A, ld, j0, 0, TB, tid);
      cp_async_commit();
      for (int k = 0; k < j; ++k) {
        const int st = k & 1;
        double* sI = st ? buf2 : buf0;
        double* sK = diag ? sI : (st ? buf3 : buf1);
        const bool more = (k + 1 < j);
        if (tid == 0) {
          int pre = 0;
          if (more) pre = (ld_acquire(&flags[i * P.nb + k + 1]) != 0) && (diag || ld_acquire(&flags[j * P.nb + k + 1]) != 0);
          s_pre[st] = pre;        // two slots: a slow reader of step k cannot see the value of step k + 1
        }
        cp_async_wait_all();
        __syncthreads();          // stage k landed; every thread is past the product of step k - 1; s_pre visible
        const bool pre = s_pre[st] != 0;
        if (more && pre) {
          load_tile_async(st ? buf0 : buf2, A, ld, i0, (k + 1) * TB, rcnt, tid);
          if (!diag) load_tile_async(st ? buf1 : buf3, A, ld, j0, (k + 1) * TB, TB, tid);
          cp_async_commit();
        }
        if (!(diag && wm == 0 && wn == 1)) mma_tile(acc, sI, sK, wm, wn, g, t);   // the upper-right quarter of a diagonal tile is never used
        if (more && !pre) {
          if (tid == 0) {
            while (ld_acquire(&flags[i * P.nb + k + 1]) == 0) {}
            if (!diag) while (ld_acquire(&flags[j * P.nb + k + 1]) == 0) {}
          }
          __syncthreads();
          load_tile_async(st ? buf0 : buf2, A, ld, i0, (k + 1) * TB, rcnt, tid);
          if (!diag) load_tile_async(st ? buf1 : buf3, A, ld, j0, (k + 1) * TB, TB, tid);
          cp_async_commit();
        }
      }
    }
    cp_async_wait_all();
    __syncthreads();   // sA0 landed; the stage buffers are free
    DF_STAMP(tile_id, 1);

    // ---- sA0 <- A0 - S, masked to the valid part ---------------------------------------------------------------
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int row = wm * 32 + a * 8 + g, col = wn * 32 + b * 8 + 2 * t;
        double2* p = reinterpret_cast<double2*>(sA0 + row * TP + col);
        double2 v = *p;
        v.x = (row < rcnt && col < jb) ? v.x - acc[a][b][0] : 0.0;
        v.y = (row < rcnt && col + 1 < jb) ? v.y - acc[a][b][1] : 0.0;
        *p = v;
      }
    __syncthreads();

    if (diag) {
      // rows >= jb of a diagonal tile lie below the square part (right-hand sides): set them aside, pad with identity
      const int ecnt = rcnt - jb;     // > 0 only in the last block column
      if (jb < TB) {
        for (int e = tid; e < TB * TB; e += DF_THREADS) {
          const int r = e >> 6, c = e & 63;
          if (r >= jb) {
            buf1[r * TP + c] = sA0[r * TP + c];
            sA0[r * TP + c] = (r == c) ? 1.0 : 0.0;
          }
        }
        __syncthreads();
      }
      DF_STAMP(tile_id, 2);
      const double2* G2 = reinterpret_cast<const double2*>(buf2);
      const bool bad = potrf64(sA0, reinterpret_cast<double2*>(buf2), sDiag, sSub, tid, tile_id);
      DF_STAMP(tile_id, 3);
      if (bad && tid == 0) P.info[0] = 1;
      // publish Z = L^-1 first: it is all the waiting tiles of this block column need
      double* Z = P.linv + (size_t)j * TB * TB;
#pragma unroll 4
      for (int e = tid; e < TB * (TB / 2); e += DF_THREADS) {
        const int c = e >> 5, k = (e & 31) * 2;
        *reinterpret_cast<double2*>(Z + c * TB + k) = make_double2(potrf_Z(G2, c, k), potrf_Z(G2, c, k + 1));
      }
      __syncthreads();
      if (tid == 0) {
        __threadfence();
        st_release(&flags[i * P.nb + j], 1);
      }
      DF_STAMP(tile_id, 4);
      // ... then L and the solved extra rows, which nothing inside this kernel reads (a diagonal tile is never an operand)
      for (int e = tid; e < TB * TB; e += DF_THREADS) {
        const int r = e >> 6, c = e & 63;
        if (r < jb && c <= r) P.A[(size_t)(i0 + r) * ld + j0 + c] = potrf_L(G2, sDiag, sSub, r, c);
      }
      if (ecnt > 0) {
        for (int e = tid; e < ecnt * TB; e += DF_THREADS) {
          const int r = jb + (e >> 6), c = e & 63;
          if (c < jb) {
            const double* er = buf1 + r * TP;
            double s0 = 0.0;
            for (int k = 0; k <= c; ++k) s0 = fma(er[k], potrf_Z(G2, c, k), s0);
            P.A[(size_t)(i0 + r) * ld + j0 + c] = s0;
          }
        }
      }
      continue;   // the flag is up already
    } else {
      // X = V Z^T with Z = L_jj^-1
      if (tid == 0) while (ld_acquire(&flags[j * P.nb + j]) == 0) {}
      __syncthreads();
      DF_STAMP(tile_id, 2);
      {
        const double* Z = P.linv + (size_t)j * TB * TB;
#pragma unroll 4
        for (int e = tid; e < TB * (TB / 2); e += DF_THREADS) {
          const int r = e >> 5, c = (e & 31) * 2;
          cp_async16(buf0 + r * TP + c, Z + r * TB + c);
        }
        cp_async_commit();
      }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
      cp_async_wait_all();
      __syncthreads();
      mma_tile_lower(acc, sA0, buf0, wm, wn, g, t);
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int row = wm * 32 + a * 8 + g, col = wn * 32 + b * 8 + 2 * t;
          if (row < rcnt) {
            double* p = P.A + (size_t)(i0 + row) * ld + j0 + col;
            if (col + 1 < jb) *reinterpret_cast<double2*>(p) = make_double2(acc[a][b][0], acc[a][b][1]);
            else if (col < jb) p[0] = acc[a][b][0];
          }
        }
    }
    __syncthreads();   // every thread's stores are issued
    if (tid == 0) {
      __threadfence();
      st_release(&flags[i * P.nb + j], 1);
    }
    DF_STAMP(tile_id, 4);
  }
  // ---- the last CTA out resets the tickets and flags for the next launch (no memset node in the captured graph) ----
  __shared__ int s_last;
  if (tid == 0) s_last = (atomicAdd(&P.sync[1], 1) == (int)gridDim.x - 1);
  __syncthreads();
  if (s_last) {
    for (int e = tid; e < P.nbr * P.nb; e += DF_THREADS) flags[e] = 0;
    if (tid == 0) { P.sync[0] = 0; P.sync[1] = 0; }
  }
}

}  // namespace

int32_t CholWs::alloc(gingr_ctx* ctx, int n, int nrows) {
  const int nb = ceil_div(n, TB), nbr = ceil_div(nrows, TB);
  const size_t ns = 2 + (size_t)nb * nbr;
  if (sync.n < ns) {
    GINGR_CUDA_TRY(ctx, sync.alloc(ns));
    GINGR_CUDA_TRY(ctx, cudaMemsetAsync(sync.p, 0, sizeof(int) * ns, ctx->stream));
  }
  GINGR_CUDA_TRY(ctx, linv.alloc((size_t)nb * TB * TB));
  cap_n = std::max(cap_n, n);
  cap_nrows = std::max(cap_nrows, nrows);
  return GINGR_OK;
}

void CholWs::release() {
  sync.release();
  linv.release();
  cap_n = cap_nrows = 0;
}

int32_t cholesky_df_enqueue(gingr_ctx* ctx, int n, int nrows, double* d_A, int ld, int* d_info, CholWs& ws) {
  static thread_local int attr_device = -1;
  if (attr_device != ctx->device) {
    GINGR_CUDA_TRY(ctx, cudaFuncSetAttribute(chol_df_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DF_SMEM));
    attr_device = ctx->device;
  }
  DfParams P;
  P.A = d_A;
  P.ld = ld;
  P.n = n;
  P.nrows = nrows;
  P.nb = ceil_div(n, TB);
  P.nbr = ceil_div(nrows, TB);
  P.ntiles = 0;
  for (int j = 0; j < P.nb; ++j) P.ntiles += P.nbr - j;
  if ((size_t)(2 + (size_t)P.nb * P.nbr) > ws.sync.n || (size_t)P.nb * TB * TB > ws.linv.n)
    return gingr_fail(ctx, GINGR_ERR_ARG, "cholesky: workspace too small");
  if ((ld & 1) != 0 || (((uintptr_t)d_A) & 15) != 0) return gingr_fail(ctx, GINGR_ERR_ARG, "cholesky: matrix must be 16-byte aligned with an even pitch");
  P.sync = ws.sync.p;
  P.linv = ws.linv.p;
  P.info = d_info;
  const int grid = std::min(P.ntiles, ctx->num_sms);
  chol_df_kernel<<<grid, DF_THREADS, DF_SMEM, ctx->stream>>>(P);
  GINGR_LAUNCHED(ctx);
  GINGR_CUDA_TRY(ctx, cudaGetLastError());
  return GINGR_OK;
}

}  // namespace gingr

#ifdef DF_TIMING
extern "C" GINGR_API int32_t gingr_debug_chol_df_timing(unsigned long long* out, int32_t count) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, gingr::df_stamps, sizeof(unsigned long long) * (size_t)count);
  return 0;
}
#endif
