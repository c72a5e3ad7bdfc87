/*
 * gingr_oracle.c -- CPU restatement of the GiNGR `update()` hot path (loops part).
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (gingr_b200/, libgingr_cuda.so) never links, imports or falls back to anything here.
 *
 * PARITY UNPINNED: the reference (unibas-gravis/GiNGR, pure Scala 3 on scalismo 1.0-RC1 +
 * Breeze) ships no golden vectors or known-answer tests (src/test/scala/DummyTest.scala.scala:1-3
 * is `assert(1 > 0)`), and neither a JVM nor the scalismo/Breeze jars exist in the build
 * container, so the reference itself cannot be run to pin this restatement.  Every function
 * below cites the reference file:line it restates; lines whose semantics live in scalismo
 * (not vendored) are marked "scalismo-recalled".
 *
 * All arithmetic is IEEE double, like the reference (Breeze DenseMatrix[Double], scalismo Point3D).
 * Point arrays are row-major [n][3].  Matrices named P are row-major [M][N]
 * (rows = moving/fit points i, columns = target points j).
 *
 * Build: see oracle/Makefile  (gcc -O2 -fopenmp -shared -fPIC; NO -ffast-math).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))

ORACLE_API int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

ORACLE_API void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* scalismo EuclideanVector3D.norm2 = x*x + y*y + z*z  (scalismo-recalled) */
static inline double norm2_3(double dx, double dy, double dz) { return dx * dx + dy * dy + dz * dz; }

/* ------------------------------------------------------------------------------------------
 * CPD soft-assignment matrix, literal.   CPD.scala:54-75
 *   gaussKernel(x, y, sigma2) = exp(-(x - y).norm2 / (2.0 * sigma2))          CPD.scala:55-57
 *   P(i,j) = gaussKernel(target_j, fit_i, sigma2)                              CPD.scala:64-68
 *   c = w/(1-w) * pow(2 pi sigma2, 3/2) * (M/N)                                CPD.scala:69-70
 *   den_j = sum_i P(i,j) + c ;  P = P ./ den                                   CPD.scala:71-74
 * Breeze sum(P, Axis._0) accumulates down a column in row order i = 0..M-1.
 * ------------------------------------------------------------------------------------------ */
ORACLE_API void oracle_cpd_P(int M, int N, const double* fit, const double* target, double sigma2, double w,
                             double* P) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < M; ++i) {
    const double yx = fit[3 * i], yy = fit[3 * i + 1], yz = fit[3 * i + 2];
    for (int j = 0; j < N; ++j) {
      const double dx = target[3 * j] - yx, dy = target[3 * j + 1] - yy, dz = target[3 * j + 2] - yz;
      P[(size_t)i * N + j] = exp(-norm2_3(dx, dy, dz) / (2.0 * sigma2));
    }
  }
  const double c = w / (1 - w) * pow(2.0 * M_PI * sigma2, 3.0 / 2.0) * ((double)M / (double)N);
#pragma omp parallel for schedule(static)
  for (int j = 0; j < N; ++j) {
    double s = 0.0;
    for (int i = 0; i < M; ++i) s += P[(size_t)i * N + j];
    const double den = s + c;
    for (int i = 0; i < M; ++i) P[(size_t)i * N + j] /= den;
  }
}

/* ------------------------------------------------------------------------------------------
 * Reductions of a materialised P used by the reference:
 *   P1  = sum(P, Axis._1)   (row sums, j ascending)        CPD.scala:36, :122, :139
 *   Pt1 = sum(P, Axis._0)   (column sums, i ascending)     CPD.scala:140
 *   PX  = P * X             (M x 3)                        CPD.scala:145
 * ------------------------------------------------------------------------------------------ */
ORACLE_API void oracle_P_reductions(int M, int N, const double* P, const double* target, double* P1, double* Pt1,
                                    double* PX) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < M; ++i) {
    double s = 0, ax = 0, ay = 0, az = 0;
    for (int j = 0; j < N; ++j) {
      const double p = P[(size_t)i * N + j];
      s += p;
      ax += p * target[3 * j];
      ay += p * target[3 * j + 1];
      az += p * target[3 * j + 2];
    }
    P1[i] = s;
    PX[3 * i] = ax;
    PX[3 * i + 1] = ay;
    PX[3 * i + 2] = az;
  }
#pragma omp parallel for schedule(static)
  for (int j = 0; j < N; ++j) {
    double s = 0;
    for (int i = 0; i < M; ++i) s += P[(size_t)i * N + j];
    Pt1[j] = s;
  }
}

/* ------------------------------------------------------------------------------------------
 * CPDCorrespondence.estimate, literal.   CPD.scala:32-49
 *   P1inv_i = 1/P1_i ; xscale_j = (P1inv_i * P(i,j)) * x_j ; deform_i = sum_j xscale_j - y_i
 *   td_i = y_i + deform_i
 * ------------------------------------------------------------------------------------------ */
ORACLE_API void oracle_cpd_correspondence(int M, int N, const double* P, const double* fit, const double* target,
                                          double* td) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < M; ++i) {
    double s = 0;
    for (int j = 0; j < N; ++j) s += P[(size_t)i * N + j];
    const double inv = 1.0 / s;
    double ax = 0, ay = 0, az = 0;
    for (int j = 0; j < N; ++j) {
      const double tmp = inv * P[(size_t)i * N + j];
      ax += tmp * target[3 * j];
      ay += tmp * target[3 * j + 1];
      az += tmp * target[3 * j + 2];
    }
    const double dx = ax - fit[3 * i], dy = ay - fit[3 * i + 1], dz = az - fit[3 * i + 2];
    td[3 * i] = fit[3 * i] + dx;
    td[3 * i + 1] = fit[3 * i + 1] + dy;
    td[3 * i + 2] = fit[3 * i + 2] + dz;
  }
}

/* ------------------------------------------------------------------------------------------
 * Streaming E-step: the same quantities as oracle_cpd_P + oracle_P_reductions without storing
 * P (two sweeps over the pairs).  Used for sizes where M*N doubles do not fit, and as the CPU
 * baseline in "algorithmic-minimum" mode.  Same formulas, same summation order per column /
 * per row as the literal version (i ascending for columns, j ascending for rows).
 *   colsum_j = sum_i K_ij ; den_j = colsum_j + c ; Pt1_j = sum_i K_ij/den_j
 *   P1_i = sum_j K_ij/den_j ;  PX_i = sum_j (K_ij/den_j) x_j
 * ------------------------------------------------------------------------------------------ */
ORACLE_API void oracle_cpd_estep_stream(int M, int N, const double* fit, const double* target, double sigma2,
                                        double w, double* P1, double* Pt1, double* PX, double* den_out) {
  const double c = w / (1 - w) * pow(2.0 * M_PI * sigma2, 3.0 / 2.0) * ((double)M / (double)N);
  const double two_s2 = 2.0 * sigma2;
  double* den = den_out ? den_out : (double*)malloc(sizeof(double) * (size_t)N);
#pragma omp parallel for schedule(static)
  for (int j = 0; j < N; ++j) {
    const double xx = target[3 * j], xy = target[3 * j + 1], xz = target[3 * j + 2];
    double s = 0.0;
    for (int i = 0; i < M; ++i) {
      const double dx = xx - fit[3 * i], dy = xy - fit[3 * i + 1], dz = xz - fit[3 * i + 2];
      s += exp(-norm2_3(dx, dy, dz) / two_s2);
    }
    const double d = s + c;
    den[j] = d;
    double t = 0.0;
    for (int i = 0; i < M; ++i) {
      const double dx = xx - fit[3 * i], dy = xy - fit[3 * i + 1], dz = xz - fit[3 * i + 2];
      t += exp(-norm2_3(dx, dy, dz) / two_s2) / d;
    }
    Pt1[j] = t;
  }
#pragma omp parallel for schedule(static)
  for (int i = 0; i < M; ++i) {
    const double yx = fit[3 * i], yy = fit[3 * i + 1], yz = fit[3 * i + 2];
    double s = 0, ax = 0, ay = 0, az = 0;
    for (int j = 0; j < N; ++j) {
      const double dx = target[3 * j] - yx, dy = target[3 * j + 1] - yy, dz = target[3 * j + 2] - yz;
      const double p = exp(-norm2_3(dx, dy, dz) / two_s2) / den[j];
      s += p;
      ax += p * target[3 * j];
      ay += p * target[3 * j + 1];
      az += p * target[3 * j + 2];
    }
    P1[i] = s;
    PX[3 * i] = ax;
    PX[3 * i + 1] = ay;
    PX[3 * i + 2] = az;
  }
  if (!den_out) free(den);
}

/* Cheaper streaming variant for the timed CPU baseline: Pt1_j = colsum_j/den_j (one exp per pair
 * and sweep instead of recomputing the column a second time).  Mathematically identical. */
ORACLE_API void oracle_cpd_estep_fast(int M, int N, const double* fit, const double* target, double sigma2,
                                      double w, double* P1, double* Pt1, double* PX) {
  const double c = w / (1 - w) * pow(2.0 * M_PI * sigma2, 3.0 / 2.0) * ((double)M / (double)N);
  const double two_s2 = 2.0 * sigma2;
  double* inv_den = (double*)malloc(sizeof(double) * (size_t)N);
#pragma omp parallel for schedule(static)
  for (int j = 0; j < N; ++j) {
    const double xx = target[3 * j], xy = target[3 * j + 1], xz = target[3 * j + 2];
    double s = 0.0;
    for (int i = 0; i < M; ++i) {
      const double dx = xx - fit[3 * i], dy = xy - fit[3 * i + 1], dz = xz - fit[3 * i + 2];
      s += exp(-norm2_3(dx, dy, dz) / two_s2);
    }
    Pt1[j] = s / (s + c);
    inv_den[j] = 1.0 / (s + c);
  }
#pragma omp parallel for schedule(static)
  for (int i = 0; i < M; ++i) {
    const double yx = fit[3 * i], yy = fit[3 * i + 1], yz = fit[3 * i + 2];
    double s = 0, ax = 0, ay = 0, az = 0;
    for (int j = 0; j < N; ++j) {
      const double dx = target[3 * j] - yx, dy = target[3 * j + 1] - yy, dz = target[3 * j + 2] - yz;
      const double p = exp(-norm2_3(dx, dy, dz) / two_s2) * inv_den[j];
      s += p;
      ax += p * target[3 * j];
      ay += p * target[3 * j + 1];
      az += p * target[3 * j + 2];
    }
    P1[i] = s;
    PX[3 * i] = ax;
    PX[3 * i + 1] = ay;
    PX[3 * i + 2] = az;
  }
  free(inv_den);
}

/* ------------------------------------------------------------------------------------------
 * computeInitialSigma2.   CPD.scala:81-90
 *   sumDist = sum over reference pm, over target pn of (pn - pm).norm2 ; / (3 N M)
 * ------------------------------------------------------------------------------------------ */
ORACLE_API double oracle_cpd_initial_sigma2(int M, int N, const double* reference, const double* target) {
  double sum = 0.0;
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < N; ++j)
      sum += norm2_3(target[3 * j] - reference[3 * i], target[3 * j + 1] - reference[3 * i + 1],
                     target[3 * j + 2] - reference[3 * i + 2]);
  return sum / (3.0 * N * M);
}

/* ------------------------------------------------------------------------------------------
 * BCPD.computeP, literal (quirks kept).   BCPD.scala:167-184
 *   mvnd = N(y_m, sigma2 I3) ; pdf(x) = (2 pi)^{-3/2} det^{-1/2} exp(-0.5 (x-y)^T Sinv (x-y))
 *          (scalismo MultivariateNormalDistribution.pdf, scalismo-recalled)
 *   e = exp(-s/(2 sigma2) * trace(I3 * Sigma_mm))                BCPD.scala:171  (s, not s^2)
 *   Phi(m,n) = pdf(x_n) * e * alpha_m                            BCPD.scala:173
 *   Pinit = Phi * (1-w)                                          BCPD.scala:176
 *   c = w / N                                                    BCPD.scala:178
 *   den_n = sum_m Pinit(m,n) * (1-w) + c                         BCPD.scala:179  ((1-w) twice)
 *   P = Pinit ./ den                                             BCPD.scala:182
 * ------------------------------------------------------------------------------------------ */
ORACLE_API void oracle_bcpd_P(int M, int N, const double* y, const double* x, const double* sigma_mm,
                              const double* alpha, double sigma2, double s, double w, double* P) {
  const double norm = pow(2.0 * M_PI, -1.5) * pow(sigma2 * sigma2 * sigma2, -0.5);
#pragma omp parallel for schedule(static)
  for (int m = 0; m < M; ++m) {
    const double e = exp(-s / (2 * sigma2) * (3.0 * sigma_mm[m]));
    for (int n = 0; n < N; ++n) {
      const double dx = x[3 * n] - y[3 * m], dy = x[3 * n + 1] - y[3 * m + 1], dz = x[3 * n + 2] - y[3 * m + 2];
      const double pdf = norm * exp(-0.5 * (norm2_3(dx, dy, dz) / sigma2));
      P[(size_t)m * N + n] = pdf * e * alpha[m] * (1 - w);
    }
  }
  const double c = w * 1.0 / (double)N;
#pragma omp parallel for schedule(static)
  for (int n = 0; n < N; ++n) {
    double sum = 0;
    for (int m = 0; m < M; ++m) sum += P[(size_t)m * N + n];
    const double den = sum * (1 - w) + c;
    for (int m = 0; m < M; ++m) P[(size_t)m * N + n] /= den;
  }
}

/* ------------------------------------------------------------------------------------------
 * Nearest vertex, O(M N) scan, FP64, ties -> lowest index (north-star rule; scalismo's KD-tree
 * tie order is unspecified).   ClosestPointRegistrator.scala:133-148 (findClosestPoint)
 * ------------------------------------------------------------------------------------------ */
ORACLE_API void oracle_nearest_vertex(int M, const double* q, int N, const double* p, int32_t* idx, double* d2out) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < M; ++i) {
    double best = INFINITY;
    int32_t bi = -1;
    for (int j = 0; j < N; ++j) {
      const double d = norm2_3(q[3 * i] - p[3 * j], q[3 * i + 1] - p[3 * j + 1], q[3 * i + 2] - p[3 * j + 2]);
      if (d < best) {
        best = d;
        bi = j;
      }
    }
    idx[i] = bi;
    if (d2out) d2out[i] = best;
  }
}

/* ------------------------------------------------------------------------------------------
 * Closest point on a triangle (Ericson, Real-Time Collision Detection 5.1.5: vertex / edge /
 * interior regions).  Stands in for scalismo `closestPointOnSurface` (scalismo-recalled, A6:
 * "exact nearest point on any triangle (vertex/edge/interior cases)").
 * The identical sequence of FP64 operations is used by the CUDA kernel so that the selected
 * triangle and point agree bit-for-bit (both compiled without FMA contraction here).
 * ------------------------------------------------------------------------------------------ */
static inline void closest_on_triangle(const double* p, const double* a, const double* b, const double* c,
                                       double* out) {
  const double abx = b[0] - a[0], aby = b[1] - a[1], abz = b[2] - a[2];
  const double acx = c[0] - a[0], acy = c[1] - a[1], acz = c[2] - a[2];
  const double apx = p[0] - a[0], apy = p[1] - a[1], apz = p[2] - a[2];
  const double d1 = abx * apx + aby * apy + abz * apz;
  const double d2 = acx * apx + acy * apy + acz * apz;
  if (d1 <= 0.0 && d2 <= 0.0) { out[0] = a[0]; out[1] = a[1]; out[2] = a[2]; return; }
  const double bpx = p[0] - b[0], bpy = p[1] - b[1], bpz = p[2] - b[2];
  const double d3 = abx * bpx + aby * bpy + abz * bpz;
  const double d4 = acx * bpx + acy * bpy + acz * bpz;
  if (d3 >= 0.0 && d4 <= d3) { out[0] = b[0]; out[1] = b[1]; out[2] = b[2]; return; }
  const double vc = d1 * d4 - d3 * d2;
  if (vc <= 0.0 && d1 >= 0.0 && d3 <= 0.0) {
    const double v = d1 / (d1 - d3);
    out[0] = a[0] + v * abx; out[1] = a[1] + v * aby; out[2] = a[2] + v * abz; return;
  }
  const double cpx = p[0] - c[0], cpy = p[1] - c[1], cpz = p[2] - c[2];
  const double d5 = abx * cpx + aby * cpy + abz * cpz;
  const double d6 = acx * cpx + acy * cpy + acz * cpz;
  if (d6 >= 0.0 && d5 <= d6) { out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; return; }
  const double vb = d5 * d2 - d1 * d6;
  if (vb <= 0.0 && d2 >= 0.0 && d6 <= 0.0) {
    const double w = d2 / (d2 - d6);
    out[0] = a[0] + w * acx; out[1] = a[1] + w * acy; out[2] = a[2] + w * acz; return;
  }
  const double va = d3 * d6 - d5 * d4;
  if (va <= 0.0 && (d4 - d3) >= 0.0 && (d5 - d6) >= 0.0) {
    const double w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
    out[0] = b[0] + w * (c[0] - b[0]); out[1] = b[1] + w * (c[1] - b[1]); out[2] = b[2] + w * (c[2] - b[2]); return;
  }
  const double denom = 1.0 / (va + vb + vc);
  const double v = vb * denom, w = vc * denom;
  out[0] = a[0] + abx * v + acx * w;
  out[1] = a[1] + aby * v + acy * w;
  out[2] = a[2] + abz * v + acz * w;
}

/* closestPointOnSurface over all triangles, O(M T) scan, ties -> lowest triangle index.
 * ClosestPointRegistrator.scala:82 ; returns point, squared distance and the triangle id. */
ORACLE_API void oracle_closest_on_surface(int M, const double* q, int N, const double* verts, int T,
                                          const int32_t* tri, double* cp, double* d2out, int32_t* tri_out) {
  (void)N;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < M; ++i) {
    double best = INFINITY, bp[3] = {0, 0, 0};
    int32_t bt = -1;
    for (int t = 0; t < T; ++t) {
      double c[3];
      closest_on_triangle(q + 3 * i, verts + 3 * tri[3 * t], verts + 3 * tri[3 * t + 1], verts + 3 * tri[3 * t + 2], c);
      const double d = norm2_3(q[3 * i] - c[0], q[3 * i + 1] - c[1], q[3 * i + 2] - c[2]);
      if (d < best) { best = d; bt = t; bp[0] = c[0]; bp[1] = c[1]; bp[2] = c[2]; }
    }
    cp[3 * i] = bp[0]; cp[3 * i + 1] = bp[1]; cp[3 * i + 2] = bp[2];
    if (d2out) d2out[i] = best;
    if (tri_out) tri_out[i] = bt;
  }
}

/* ------------------------------------------------------------------------------------------
 * Mesh predicates (scalismo-recalled).
 *   cellNormal = normalize((p2-p1) x (p3-p1)); vertexNormal = normalize(mean of adjacent cell
 *   normals); boundary edge = edge in exactly one triangle; pointIsOnBoundary = vertex touches
 *   a boundary edge.
 * ------------------------------------------------------------------------------------------ */
ORACLE_API void oracle_vertex_normals(int N, const double* v, int T, const int32_t* tri, double* normals) {
  double* acc = (double*)calloc((size_t)3 * N, sizeof(double));
  int* cnt = (int*)calloc((size_t)N, sizeof(int));
  for (int t = 0; t < T; ++t) {
    const double* a = v + 3 * tri[3 * t];
    const double* b = v + 3 * tri[3 * t + 1];
    const double* c = v + 3 * tri[3 * t + 2];
    const double ux = b[0] - a[0], uy = b[1] - a[1], uz = b[2] - a[2];
    const double wx = c[0] - a[0], wy = c[1] - a[1], wz = c[2] - a[2];
    double nx = uy * wz - uz * wy, ny = uz * wx - ux * wz, nz = ux * wy - uy * wx;
    const double len = sqrt(nx * nx + ny * ny + nz * nz);
    nx /= len; ny /= len; nz /= len;
    for (int k = 0; k < 3; ++k) {
      const int id = tri[3 * t + k];
      acc[3 * id] += nx; acc[3 * id + 1] += ny; acc[3 * id + 2] += nz;
      cnt[id]++;
    }
  }
  for (int i = 0; i < N; ++i) {
    double nx = acc[3 * i], ny = acc[3 * i + 1], nz = acc[3 * i + 2];
    if (cnt[i] > 0) { nx /= cnt[i]; ny /= cnt[i]; nz /= cnt[i]; }
    const double len = sqrt(nx * nx + ny * ny + nz * nz);
    normals[3 * i] = nx / len; normals[3 * i + 1] = ny / len; normals[3 * i + 2] = nz / len;
  }
  free(acc); free(cnt);
}

typedef struct { int32_t a, b; } edge_t;
static int edge_cmp(const void* x, const void* y) {
  const edge_t* e = (const edge_t*)x; const edge_t* f = (const edge_t*)y;
  if (e->a != f->a) return e->a < f->a ? -1 : 1;
  if (e->b != f->b) return e->b < f->b ? -1 : 1;
  return 0;
}

ORACLE_API void oracle_boundary_vertices(int N, int T, const int32_t* tri, uint8_t* on_boundary) {
  memset(on_boundary, 0, (size_t)N);
  edge_t* e = (edge_t*)malloc(sizeof(edge_t) * (size_t)3 * T);
  for (int t = 0; t < T; ++t)
    for (int k = 0; k < 3; ++k) {
      int32_t a = tri[3 * t + k], b = tri[3 * t + (k + 1) % 3];
      if (a > b) { int32_t s = a; a = b; b = s; }
      e[3 * t + k].a = a; e[3 * t + k].b = b;
    }
  qsort(e, (size_t)3 * T, sizeof(edge_t), edge_cmp);
  for (int i = 0; i < 3 * T;) {
    int j = i + 1;
    while (j < 3 * T && e[j].a == e[i].a && e[j].b == e[i].b) ++j;
    if (j - i == 1) { on_boundary[e[i].a] = 1; on_boundary[e[i].b] = 1; }
    i = j;
  }
  free(e);
}

/* ------------------------------------------------------------------------------------------
 * isClosestPointIntersecting.   ClosestPointRegistrator.scala:62-72
 *   p = mesh.point(id) ; v = p - cp ; all intersections of the INFINITE line p + s v with the
 *   mesh triangles (scalismo getIntersectionPoints, scalismo-recalled A6), drop points == p,
 *   result = min |p - ip| < |v|.
 * Line/triangle test: Moeller-Trumbore without the s >= 0 restriction.  Returns, per query,
 * the minimum distance to an intersection point != p (INFINITY if none).
 * ------------------------------------------------------------------------------------------ */
static void line_mesh_nearest(int M, const double* p, const double* dir, const double* verts, int T, const int32_t* tri,
                              int skip_incident, double* min_dist, double* hit_pt);

ORACLE_API void oracle_line_mesh_min_dist(int M, const double* p, const double* dir, int N, const double* verts, int T,
                                          const int32_t* tri, int skip_incident, double* min_dist) {
  (void)N;
  line_mesh_nearest(M, p, dir, verts, T, tri, skip_incident, min_dist, NULL);
}

/* ClosestPointAlongNormalTriangleMesh3D (ClosestPointRegistrator.scala:105-110): nearest intersection != p of the
 * line (p, n) with the TARGET mesh: intersectingPoints.minBy(ip => (p - ip).norm); minBy keeps the FIRST minimum,
 * i.e. the lowest triangle index in this scan order.  hit_pt[i] = p_i and min_dist = INFINITY when there is none. */
ORACLE_API void oracle_line_mesh_nearest(int M, const double* p, const double* dir, int N, const double* verts, int T,
                                         const int32_t* tri, double* min_dist, double* hit_pt) {
  (void)N;
  line_mesh_nearest(M, p, dir, verts, T, tri, 0, min_dist, hit_pt);
}

static void line_mesh_nearest(int M, const double* p, const double* dir, const double* verts, int T, const int32_t* tri,
                              int skip_incident, double* min_dist, double* hit_pt) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < M; ++i) {
    const double ox = p[3 * i], oy = p[3 * i + 1], oz = p[3 * i + 2];
    const double dx = dir[3 * i], dy = dir[3 * i + 1], dz = dir[3 * i + 2];
    double best = INFINITY, bx = ox, by = oy, bz = oz;
    for (int t = 0; t < T; ++t) {
      /* query i is vertex i of this mesh: its incident triangles meet the line only at p itself, which
       * the reference removes with .filter(f => f != p) (:67) -- skip them outright */
      if (skip_incident && (tri[3 * t] == i || tri[3 * t + 1] == i || tri[3 * t + 2] == i)) continue;
      const double* a = verts + 3 * tri[3 * t];
      const double* b = verts + 3 * tri[3 * t + 1];
      const double* c = verts + 3 * tri[3 * t + 2];
      const double e1x = b[0] - a[0], e1y = b[1] - a[1], e1z = b[2] - a[2];
      const double e2x = c[0] - a[0], e2y = c[1] - a[1], e2z = c[2] - a[2];
      const double px = dy * e2z - dz * e2y, py = dz * e2x - dx * e2z, pz = dx * e2y - dy * e2x;
      const double det = e1x * px + e1y * py + e1z * pz;
      if (det == 0.0) continue; /* parallel */
      const double inv = 1.0 / det;
      const double tx = ox - a[0], ty = oy - a[1], tz = oz - a[2];
      const double u = (tx * px + ty * py + tz * pz) * inv;
      if (u < 0.0 || u > 1.0) continue;
      const double qx = ty * e1z - tz * e1y, qy = tz * e1x - tx * e1z, qz = tx * e1y - ty * e1x;
      const double vv = (dx * qx + dy * qy + dz * qz) * inv;
      if (vv < 0.0 || u + vv > 1.0) continue;
      const double s = (e2x * qx + e2y * qy + e2z * qz) * inv;
      const double ix = ox + s * dx, iy = oy + s * dy, iz = oz + s * dz;
      if (ix == ox && iy == oy && iz == oz) continue; /* .filter(f => f != p)  :67 */
      const double d = sqrt(norm2_3(ox - ix, oy - iy, oz - iz));
      if (d < best) { best = d; bx = ix; by = iy; bz = iz; }
    }
    min_dist[i] = best;
    if (hit_pt) { hit_pt[3 * i] = bx; hit_pt[3 * i + 1] = by; hit_pt[3 * i + 2] = bz; }
  }
}
