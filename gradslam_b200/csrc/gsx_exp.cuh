// exp(x) for the confidence weight alpha = clamp(exp(-|v|^2 / 2 sigma^2), 1e-7, 1.01) (gradslam/slam/fusionutils.py:69-72).
// The canonical value is the float32 rounding of the float64 exponential (oracle: torch.exp in float64, rounded once).
// CUDA's double-precision exp() costs ~80 instructions per pixel in K4; the arguments here are float32 values in
// [-17, 0] (below that the weight clamps to 1e-7), so a reduced-range evaluation is enough:
//     k = rint(x log2 e),  r = x - k ln2 (two-constant Cody-Waite, exact with FMA),  exp(r) by a degree-13 Taylor
//     polynomial in Horner form with FMA (truncation < 2^-57, rounding a few double ulps),  2^k by an exponent add.
// Rounding that double to float32 gives the same float as rounding the exact exponential unless the exact value sits
// within a few double ulps of a float32 rounding boundary; such results (bits 28..0 within +-kExpGuard of the midpoint
// pattern, odds 2^-22) are flagged and the caller falls back to the library exp().  tests/test_exp_host.py runs this
// header on the host for EVERY float32 argument in [-17, -2^-40] and checks that the unflagged results round to the
// same float32 as libm's exp.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>
#ifndef GSX_HD
#ifdef __CUDACC__
#define GSX_HD __host__ __device__
#else
#define GSX_HD
#endif
#endif

namespace gsx {

constexpr int64_t kExpGuard = 64;    // double ulps around the float32 rounding midpoint treated as undecided

GSX_HD inline double exp_bits_to_double(int64_t b) {
#ifdef __CUDA_ARCH__
  return __longlong_as_double(b);
#else
  double d;
  memcpy(&d, &b, 8);
  return d;
#endif
}
GSX_HD inline int64_t exp_double_to_bits(double d) {
#ifdef __CUDA_ARCH__
  return __double_as_longlong(d);
#else
  int64_t b;
  memcpy(&b, &d, 8);
  return b;
#endif
}

// x in [-17.5, 0].  Returns exp(x) to a few double ulps; *undecided is set when rounding the result to float32 could
// differ from rounding the exact value.
GSX_HD inline double exp_reduced(double x, bool *undecided) {
  const double kLog2e = 1.4426950408889634074, kLn2Hi = 6.93147180369123816490e-01, kLn2Lo = 1.90821492927058770002e-10;
  const double k = rint(x * kLog2e);
  double r = fma(-k, kLn2Hi, x);
  r = fma(-k, kLn2Lo, r);
  double p = 1.0 / 6227020800.0;            // 1/13!
  p = fma(p, r, 1.0 / 479001600.0);         // 1/12!
  p = fma(p, r, 1.0 / 39916800.0);          // 1/11!
  p = fma(p, r, 1.0 / 3628800.0);           // 1/10!
  p = fma(p, r, 1.0 / 362880.0);            // 1/9!
  p = fma(p, r, 1.0 / 40320.0);             // 1/8!
  p = fma(p, r, 1.0 / 5040.0);              // 1/7!
  p = fma(p, r, 1.0 / 720.0);               // 1/6!
  p = fma(p, r, 1.0 / 120.0);               // 1/5!
  p = fma(p, r, 1.0 / 24.0);                // 1/4!
  p = fma(p, r, 1.0 / 6.0);                 // 1/3!
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);                       // exp(r), in [0.70, 1.42]
  const int ki = (int)k;
  const int64_t bits = exp_double_to_bits(p) + (int64_t)((uint64_t)(int64_t)ki << 52);  // * 2^k, k in [-26, 0]: no underflow
  const int64_t low = bits & 0x1FFFFFFFll;                          // the 29 bits below a float32 mantissa
  const int64_t dist = low - 0x10000000ll;
  *undecided = (dist < kExpGuard) && (dist > -kExpGuard);
  return exp_bits_to_double(bits);
}

// float32(exp(float64(x))) for a float32 x; any x (NaN, +-inf, positive) is handled by the library path.
GSX_HD inline float exp_f32_via_f64(float x) {
  if (x >= -17.0f && x <= 0.0f) {
    bool undecided;
    const double e = exp_reduced((double)x, &undecided);
    if (!undecided) return (float)e;
  }
  return (float)exp((double)x);
}

}  // namespace gsx
