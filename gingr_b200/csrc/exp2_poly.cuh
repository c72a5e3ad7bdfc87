// exp2_poly.cuh -- FP64 2^x for x <= 0 on the FP64 pipe, used by the E-step kernels.
//
// The Gaussian affinity exp(-d2 / (2 sigma2)) (CPD.scala:55-57) is evaluated as 2^(negk * d2) with
// negk = -log2(e) / (2 sigma2): the scale folds into the range-reduction FMA, so no extra multiply.
//   tmp = fma(d2, negk, SHIFT)      SHIFT = 1.5*2^52 + 2048  ->  low word of tmp = 2048 + n, n = rint(negk*d2)
//   r   = fma(d2, negk, -(tmp - SHIFT))      in [-0.5, 0.5], one rounding
//   2^r = degree-10 minimax polynomial (max relative error 2.1e-16; tools/gen_exp2_poly.py 10 0.5)
//   K'  = p * 2^(n + 64)  by exponent-field addition
// The value returned is K' = 2^64 * K.  The bias keeps every representable K (down to the smallest subnormal,
// 2^-1074) in the NORMAL range of K', so gradual underflow needs no special path: n in [-1085, 0] is the fast
// and only path, anything smaller is exactly 0 (K < 2^-1085 rounds to 0 in the reference as well).  Callers
// fold the 2^-64 into their column weights (a power of two: exact).  ncu showed each FP64 instruction costs
// two issue slots on B200 and every other instruction one, so the integer tail is kept to six instructions:
// two compares, two for the exponent, two selects.  Non-finite inputs are caught by a separate O(M+N)
// validation pass (a NaN would otherwise select the zero branch).
// 13 FP64 instructions per evaluation (3 range reduction + 10 Horner FMAs).
#pragma once
#include "exp2_tab.cuh"

namespace gingr {

constexpr double GAUSS_BIAS_SCALE = 0x1p+64;      // K' = K * 2^64
constexpr double GAUSS_BIAS_UNSCALE = 0x1p-64;

__device__ __forceinline__ double exp2_poly10(double r) {
  double p = 0x1.e3991ef300b90p-28;
  p = fma(p, r, 0x1.b6740fc4fd79dp-24);
  p = fma(p, r, 0x1.62c157ee08d1ep-20);
  p = fma(p, r, 0x1.ffcb55e82c907p-17);
  p = fma(p, r, 0x1.4309126056904p-13);
  p = fma(p, r, 0x1.5d87fe9cc5d7dp-10);
  p = fma(p, r, 0x1.3b2ab6fbde0f7p-7);
  p = fma(p, r, 0x1.c6b08d703d48ap-5);
  p = fma(p, r, 0x1.ebfbdff82c3b9p-3);
  p = fma(p, r, 0x1.62e42fefa3a17p-1);
  p = fma(p, r, 1.0);
  return p;
}

// 2^64 * 2^(negk * d2) for finite negk * d2 <= 0 (any magnitude).
__device__ __forceinline__ double gauss_exp2_biased(double d2, double negk) {
  const double SHIFT = 6755399441055744.0 + 2048.0;  // 1.5 * 2^52 + 2048
  const double tmp = fma(d2, negk, SHIFT);
  const int lo = __double2loint(tmp);   // 2048 + n
  const int hi = __double2hiint(tmp);   // 0x43380000 exactly while -2048 <= n < 2^32 - 2048
  const double nf = tmp - SHIFT;
  const double r = fma(d2, negk, -nf);
  const double p = exp2_poly10(r);
  // n in [-1085, 0]  <=>  lo in [963, 2048]
  const bool ok = (hi == 0x43380000) && ((unsigned)(lo - 963) <= 1085u);
  // exponent += n + 64 = lo - 1984
  const int phi = __double2hiint(p) + ((lo - 1984) << 20);
  return __hiloint2double(ok ? phi : 0, ok ? __double2loint(p) : 0);
}

// ---------------------------------------------------------------------------------------------------
// Table variant (the one the E-step sweeps use), N = GAUSS_TAB_N = 256 entries (the names below say "64" for the scaled
// quantities of the first, 64-entry version): 2^x = 2^n * 2^(k/N) * 2^(r/N), x*N = N n + k + r, |r| <= 1/2.
//   tmp = fma(d2, negk64, SHIFT64)        low word = (2048 + n) * N + k          negk64 = -N log2(e) / (2 sigma2)
//   r   = fma(d2, negk64, -(tmp - SHIFT64))
//   T   = table[k] with its exponent field advanced by n + 64 in ONE integer multiply-add (the table's hi words
//         are pre-adjusted, exp2_tab.cuh); the table lives in shared memory, replicated 16x so that the 16 lanes of
//         a half warp hit 16 distinct bank pairs whatever their k (LDS.64 is served per half warp)
//   K'  = T + (T r) q(r),  q = degree-3 Taylor tail of 2^(r/N) (truncation 3.8e-17 with N = 256)
// 9 FP64 instructions (3 range reduction, 4 Horner, 1 mul, 1 fma) instead of 13, and 7 other instructions instead
// of 12: the sweeps are bound by the FP64 pipe (2 cycles per warp instruction), so this is a direct 4/20 saving
// per pair in sweep A and 4/23 in sweep B.
constexpr int GAUSS_TAB_BYTES = GAUSS_TAB_N * 16 * 8;
constexpr double GAUSS_SCALE = (double)GAUSS_TAB_N;            // the scaled exponent is x * GAUSS_TAB_N
constexpr int GAUSS_TAB_SHIFT = 20 - GAUSS_TAB_BITS;            // low word -> exponent field of the table's hi word
constexpr int GAUSS_TAB_MASK = (GAUSS_TAB_N - 1) << 7;          // byte offset of row k (16 replicas x 8 bytes)
constexpr int GAUSS_LO_MIN = 963 * GAUSS_TAB_N;                 // n >= -1085
constexpr unsigned GAUSS_LO_SPAN = 1085u * GAUSS_TAB_N;

__device__ __forceinline__ void gauss_tab_fill(unsigned int* s_tab /*[N*16*2]*/, int tid, int nthreads) {
  for (int e = tid; e < GAUSS_TAB_N * 16; e += nthreads) {
    s_tab[2 * e] = GAUSS_EXP2_TAB[e >> 4][0];
    s_tab[2 * e + 1] = GAUSS_EXP2_TAB[e >> 4][1];
  }
}

// polynomial constants in constant memory: DFMA takes a c[bank][offset] operand directly, whereas 64-bit
// immediates are re-materialised with UMOV pairs in the loop (3 extra issue slots per evaluation)
__constant__ double GAUSS_C[8] = {GAUSS_EXP2_A1, GAUSS_EXP2_A2, GAUSS_EXP2_A3, GAUSS_EXP2_A4, 0.0,
                                  6755399441055744.0 + 2048.0 * GAUSS_TAB_N /* SHIFT = 1.5 * 2^52 + 2048 N */, 0.0, 0.0};

// Returns K' = 2^64 * 2^(negk64 * d2 / 64), or exactly 0 when K underflows (n < -1085) or the argument is out of
// range.  The zero is SELECTED (not a skipped accumulation): the reference evaluates 0 * (1 / 0) = NaN for a column
// whose denominators vanish (CPD.scala:71-74), and that NaN must reach P1 / PX.
// lane_off = (lane & 15) * 8; s_tab = the CTA's table (gauss_tab_fill).
// SAFE: the caller guarantees 0 <= -negk64 * d2 < 2^31 - 2^18 for every pair of the launch (bounding boxes, see
// estep_safe_range), so the high word of tmp cannot leave {0x43380000, 0x4337ffff} and ONE signed compare of the
// low word decides validity -- two issue slots less per evaluation in a loop that is issue bound.
template <bool SAFE>
__device__ __forceinline__ double gauss_exp2_tab(double d2, double negk64, const unsigned int* s_tab, int lane_off) {
  const double SHIFT = GAUSS_C[5];
  const double tmp = fma(d2, negk64, SHIFT);
  const int lo = __double2loint(tmp);  // (2048 + n) * 64 + k
  const int hi = __double2hiint(tmp);
  const double nf = tmp - SHIFT;
  const double r = fma(d2, negk64, -nf);
  const uint2 t = *reinterpret_cast<const uint2*>(reinterpret_cast<const char*>(s_tab) + (((lo << 7) & GAUSS_TAB_MASK) | lane_off));
  const double T = __hiloint2double((int)(((unsigned)lo << GAUSS_TAB_SHIFT) + t.y), (int)t.x);
  double p = GAUSS_C[3];
  p = fma(p, r, GAUSS_C[2]);
  p = fma(p, r, GAUSS_C[1]);
  p = fma(p, r, GAUSS_C[0]);
  const double s = T * r;
  const double v = fma(s, p, T);
  // n in [-1085, 0]  <=>  lo in [963 * 64, 2048 * 64]
  const bool ok = SAFE ? (lo >= GAUSS_LO_MIN) : ((hi == 0x43380000) && ((unsigned)(lo - GAUSS_LO_MIN) <= GAUSS_LO_SPAN));
  return __hiloint2double(ok ? __double2hiint(v) : 0, ok ? __double2loint(v) : 0);
}

// The same with the scaled argument u = d2 * negk64 already formed (expanded-distance path of estep.cu): the integer
// part comes from one add against the shift constant, the remainder r = u - n is exact.  Range-guaranteed callers only
// (0 <= -u < 2^31 - 2^18): one compare.
__device__ __forceinline__ double gauss_exp2_tab_u(double u, const unsigned int* s_tab, int lane_off) {
  const double SHIFT = GAUSS_C[5];
  const double tmp = u + SHIFT;
  const int lo = __double2loint(tmp);  // (2048 + n) * 64 + k
  const double nf = tmp - SHIFT;
  const double r = u - nf;
  const uint2 t = *reinterpret_cast<const uint2*>(reinterpret_cast<const char*>(s_tab) + (((lo << 7) & GAUSS_TAB_MASK) | lane_off));
  const double T = __hiloint2double((int)(((unsigned)lo << GAUSS_TAB_SHIFT) + t.y), (int)t.x);
  double p = GAUSS_C[3];
  p = fma(p, r, GAUSS_C[2]);
  p = fma(p, r, GAUSS_C[1]);
  p = fma(p, r, GAUSS_C[0]);
  const double s = T * r;
  const double v = fma(s, p, T);
  const bool ok = lo >= GAUSS_LO_MIN;
  return __hiloint2double(ok ? __double2hiint(v) : 0, ok ? __double2loint(v) : 0);
}

}  // namespace gingr
