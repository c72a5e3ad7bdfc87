// exp2_poly.cuh -- FP64 2^x for x <= 0 on the FP64 pipe, used by the E-step kernels.
//
// The Gaussian affinity exp(-d2 / (2 sigma2)) (CPD.scala:55-57) is evaluated as 2^(negk * d2) with
// negk = -log2(e) / (2 sigma2): the scale folds into the range reduction FMA, so no extra multiply.
//   tmp = fma(d2, negk, 1.5*2^52)   -> low word of tmp = n = rint(negk*d2)
//   r   = fma(d2, negk, -(tmp - 1.5*2^52))  in [-0.5, 0.5], exact single rounding
//   2^r = degree-10 minimax polynomial (max relative error 2.1e-16; tools/gen_exp2_poly.py 10 0.5)
//   result = p * 2^n by exponent-field addition; gradual underflow handled on a rare path.
// 13 FP64 instructions per evaluation (3 range reduction + 10 Horner FMAs).
#pragma once

namespace gingr {

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

// 2^(negk * d2) for negk * d2 <= 0 (any magnitude).  NaN propagates.
__device__ __forceinline__ double gauss_exp2(double d2, double negk) {
  const double SHIFT = 6755399441055744.0;  // 1.5 * 2^52
  const double tmp = fma(d2, negk, SHIFT);
  const int n = __double2loint(tmp);
  const int hi = __double2hiint(tmp);
  const double nf = tmp - SHIFT;
  const double r = fma(d2, negk, -nf);
  const double p = exp2_poly10(r);
  // fast path: -1021 <= n <= 0 and |negk*d2| < 2^32 (hi word of tmp is 0x4337FFFF or 0x43380000, or NaN)
  const bool in_range = (hi >= 0x4337FFFF) && ((unsigned)(n + 1021) <= 1021u);
  const int nn = in_range ? n : 0;
  double res = __hiloint2double(__double2hiint(p) + (nn << 20), __double2loint(p));
  if (!in_range) {
    res = 0.0;
    // gradual underflow band 2^-1080 .. 2^-1021: scale in two steps (rare)
    if (hi == 0x4337FFFF && n < -1021 && n >= -1080) {
      const double s = __hiloint2double(__double2hiint(p) + ((n + 128) << 20), __double2loint(p));
      res = s * 0x1p-128;
    }
  }
  return res;
}

}  // namespace gingr
