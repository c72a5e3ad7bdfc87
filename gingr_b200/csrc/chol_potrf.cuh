// chol_potrf.cuh -- the in-CTA 64 x 64 factorisation of the data-flow Cholesky (chol_df.cu); device code only, so that
// tools/potrf_bench.cu can time it in isolation.  Expects TB, TP, DF_THREADS, DF_CLOCK, PF_MARK to be defined.
#pragma once

// ---------------------------------------------------------------------------------------------
// In-CTA factorisation of a 64 x 64 block, L AND L^-1 in one pass (all 256 threads), itself a small data-flow.
//
// What bounds it (tools/lat_bench.cu on a B200): ONE warp issues a DFMA/DMUL only every ~6 cycles (8 cycles dependent), a
// 64-bit shuffle costs 26, a shared-memory exchange 35, the branch-free rsqrt 49.  So (1) the trailing updates must be
// spread over all warps, (2) the dependent chain  pivot -> rsqrt -> scale -> exchange -> update  is walked as few times as
// possible and must be short in INSTRUCTIONS, and (3) nothing but that chain may sit between two pivots (measured and
// dropped: CTA barrier per step with every warp doing chain + update, 584 cycles per step; 4-column steps with a redundant
// 4 x 4 factorisation per thread, 1350; one row per thread with the pivots exchanged through a named barrier, 915).
// Layout: warp g holds the column PAIRS 4 g .. 4 g + 3 of the rows `lane` and `lane + 32`: 16 entries per thread.
// Step s eliminates the columns j = 2 s and j + 1 at once: with p = a_jj, q = a_{j+1,j}, t = a_{j+1,j+1} the second pivot is
// det / p, det = p t - q^2, so rsqrt(p) and rsqrt(det) are independent (one chain per TWO columns; det loses what
// t - (q rs)^2 loses).  The owner warp (g = s mod 8) gets (p, q, t) by shuffles, scales its two rows, stores two factors
// per row (G2) and raises ready[s] (release / acquire on a shared-memory word: no CTA barrier inside the factorisation).
// The other warps apply the rank-2 update to their own entries whenever they get to it; only the NEXT owner is in a
// hurry, so it updates just its pivot pair, runs its chain, and catches up on the rest afterwards.
// L^-1 rides along for free: a row that has been eliminated continues as the corresponding row of an identity appended
// below the block -- rows below the square part receive X = I L^-T under exactly the same column operations, and row m
// of the identity is zero before its own step -- so every thread always carries two ACTIVE rows, and at the end
//   G2[r][s] = (L[r][2s], L[r][2s+1])  for r > 2 s + 1,      G2[r][s] = (Z[2s][r], Z[2s+1][r])  for r <= 2 s + 1   (Z = L^-1)
// with L[2s+1][2s] in sSub[s] and the diagonal of L in sDiag.
// ---------------------------------------------------------------------------------------------
constexpr int G2P = 33;   // double2 pitch of a G2 row: 132 words == 4 (mod 32), conflict-free LDS.128 down a column

__device__ __forceinline__ double rsq_nr(double d) {   // MUFU.RSQ64H seed + one cubic step: rsqrt() without its slow-path branch
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(d));
  const double e = fma(-d * y0, y0, 1.0);
  const double h = fma(e, 0.375, 0.5);
  return fma(y0 * e, h, y0);
}
// step flags of the in-CTA factorisation: one mbarrier per step (count 1: the owner's lane 0 arrives, everybody else
// sleeps in try_wait -- a polling loop on a shared-memory word was measured to steal half the issue slots of the warp on
// the critical path).  Each factorisation completes exactly one phase of every barrier, so the parity alternates per call.
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(sa), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.release.cta.shared::cta.b64 st, [%0];\n}\n" ::"r"(sa) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n"
      " .reg .pred p;\n"
      "LAB_WAIT:\n"
      " mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%0], %1;\n"
      " @p bra LAB_DONE;\n"
      " bra LAB_WAIT;\n"
      "LAB_DONE:\n"
      "}\n" ::"r"(sa), "r"(parity)
      : "memory");
}

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

// sT: the block (lower triangle, pitch TP, identity-padded beyond the valid part).  On return G2 / sDiag / sSub hold L and
// Z = L^-1 as described above (read them with potrf_L / potrf_Z).  sBar: the 32 step barriers (initialised once per
// kernel), parity: number of factorisations this CTA has run before, mod 2.  Returns true if a pivot was not positive and
// finite.
// Zout (optional): global [64][64] that receives Z row by row WHILE the factorisation runs -- rows 2 s and 2 s + 1 are final
// as soon as step s is published (column s of G2), and the warps whose columns are all eliminated have nothing else to do:
// warp s mod (s / 4) copies them.  Rows 0..7 (steps 0..3: no warp is idle yet) are left to the caller.
__device__ __forceinline__ bool potrf64(const double* __restrict__ sT, double2* G2, double* __restrict__ sDiag,
                                        double* __restrict__ sSub, unsigned long long* sBar, unsigned parity, int tid,
                                        int tile_id, double* __restrict__ Zout = nullptr) {
  const int lane = tid & 31, g = tid >> 5;
  (void)tile_id;
  DF_CLOCK(tile_id, 0);
  double a0[8], a1[8];   // rows lane / lane + 32; local pair li = columns 2 (4 g + li) + e at index 2 li + e
#pragma unroll
  for (int li = 0; li < 4; ++li)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int col = 2 * (4 * g + li) + e;
      a0[2 * li + e] = (col <= lane) ? sT[lane * TP + col] : 0.0;
      a1[2 * li + e] = (col <= lane + 32) ? sT[(lane + 32) * TP + col] : 0.0;
    }
  bool bad = false;
  // rank-2 update of the local pair li with the factors (Fa, Fb of this thread's two rows) of step st; a row eliminated
  // at that step (za / zb) restarts from zero: it continues as a row of the identity
#define POTRF_UPDATE_PAIR(li, st, Fa, Fb, za, zb)                                                  \
  {                                                                                                \
    const int qc = 2 * (4 * g + (li));                                                             \
    const double2 c0 = G2[qc * G2P + (st)], c1 = G2[(qc + 1) * G2P + (st)];                        \
    a0[2 * (li)] = fma(-Fa.y, c0.y, fma(-Fa.x, c0.x, za ? 0.0 : a0[2 * (li)]));                    \
    a0[2 * (li) + 1] = fma(-Fa.y, c1.y, fma(-Fa.x, c1.x, za ? 0.0 : a0[2 * (li) + 1]));            \
    a1[2 * (li)] = fma(-Fb.y, c0.y, fma(-Fb.x, c0.x, zb ? 0.0 : a1[2 * (li)]));                    \
    a1[2 * (li) + 1] = fma(-Fb.y, c1.y, fma(-Fb.x, c1.x, zb ? 0.0 : a1[2 * (li) + 1]));            \
  }
#pragma unroll 1
  for (int go = 0; go < 8; ++go) {      // the warp that owns the pairs 4 go .. 4 go + 3
    if (g < go) break;                  // all columns of this warp are eliminated
#pragma unroll
    for (int li = 0; li < 4; ++li) {
      const int s = 4 * go + li;
      const int j = 2 * s, pl = j & 31;             // the pivot rows j, j + 1 are lanes pl, pl + 1 of the lower / upper rows
      const bool up = (s >= 16);
      const bool swA = (lane == pl), swB = (lane == pl + 1);
      const bool za = !up && (swA || swB), zb = up && (swA || swB);
      if (g == go) {
        // ---- the chain: pivot block -> two independent rsqrt -> scaled columns of both rows ----------------------
        PF_MARK(s, 0);
        const double v0 = up ? a1[2 * li] : a0[2 * li];
        const double v1 = up ? a1[2 * li + 1] : a0[2 * li + 1];
        const double p = __shfl_sync(0xffffffffu, v0, pl);
        const double q = __shfl_sync(0xffffffffu, v0, pl + 1);
        const double t = __shfl_sync(0xffffffffu, v1, pl + 1);
#ifdef POTRF_PROFILE
        if (p + q + t == 12345.678) bad = true;
        PF_MARK(s, 1);
#endif
        const double det = fma(p, t, -q * q);
        const double rs1 = rsq_nr(p), rsd = rsq_nr(det);
#ifdef POTRF_PROFILE
        if (rs1 + rsd == 12345.678) bad = true;
        PF_MARK(s, 2);
#endif
        const double spv = p * rs1;        // L[j][j]
        const double l21 = q * rs1;        // L[j+1][j]
        const double i22 = rsd * spv;      // 1 / L[j+1][j+1]
        // rows j and j + 1 turn into rows of the appended identity: (1, 0) and (0, 1) in the pivot columns
        const double xa = za ? (swA ? 1.0 : 0.0) : a0[2 * li], ya = za ? (swA ? 0.0 : 1.0) : a0[2 * li + 1];
        const double xb = zb ? (swA ? 1.0 : 0.0) : a1[2 * li], yb = zb ? (swA ? 0.0 : 1.0) : a1[2 * li + 1];
        double2 Fa, Fb;
        Fa.x = xa * rs1;
        Fb.x = xb * rs1;
        Fa.y = fma(-Fa.x, l21, ya) * i22;
        Fb.y = fma(-Fb.x, l21, yb) * i22;
#ifdef POTRF_PROFILE
        if (Fa.y + Fb.y == 12345.678) bad = true;
        PF_MARK(s, 3);
#endif
        G2[lane * G2P + s] = Fa;
        G2[(lane + 32) * G2P + s] = Fb;
        if (lane == 0) {
          sDiag[j] = spv;
          sDiag[j + 1] = fma(-l21, l21, t) * i22;
          sSub[s] = l21;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sBar[s]);
        PF_MARK(s, 4);
        if (!(p > 0.0) || !(p < INFINITY) || !(det > 0.0) || !(det < INFINITY)) bad = true;
        // the own later pairs (the next chain needs pair li + 1 first)
#pragma unroll
        for (int l2 = li + 1; l2 < 4; ++l2) POTRF_UPDATE_PAIR(l2, s, Fa, Fb, za, zb)
        PF_MARK(s, 5);
      } else {
        mbar_wait(&sBar[s], parity);
        if (g == go + 1) PF_MARK(s, 6);
        const double2 Fa = G2[lane * G2P + s], Fb = G2[(lane + 32) * G2P + s];
#pragma unroll
        for (int l2 = 0; l2 < 4; ++l2) POTRF_UPDATE_PAIR(l2, s, Fa, Fb, za, zb)
        if (g == go + 1) PF_MARK(s, 7);
      }
    }
  }
#undef POTRF_UPDATE_PAIR
  if (Zout != nullptr && g < 7) {
#pragma unroll 1
    for (int s = 4 * (g + 1); s < 32; ++s) {
      if (s % (s >> 2) != g) continue;
      mbar_wait(&sBar[s], parity);
      const int c = 2 * s;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k = lane + 32 * h;
        const double2 v = G2[k * G2P + s];
        Zout[c * 64 + k] = (k <= c) ? v.x : 0.0;
        Zout[(c + 1) * 64 + k] = (k <= c + 1) ? v.y : 0.0;
      }
    }
  }
  __syncthreads();
  DF_CLOCK(tile_id, 1);
  return __syncthreads_or(bad ? 1 : 0) != 0;
}

