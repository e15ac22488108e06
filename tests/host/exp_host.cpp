// Exhaustive host check of gradslam_b200/csrc/gsx_exp.cuh (see tests/test_exp_host.py): every float32 argument in a
// range of bit patterns; counts arguments whose reduced-range result is flagged undecided and arguments whose
// UNFLAGGED result rounds to a different float32 than libm's double exp.
#include "../../gradslam_b200/csrc/gsx_exp.cuh"

extern "C" void exp_scan(uint32_t bits_begin, uint32_t bits_end, uint64_t *n_undecided, uint64_t *n_mismatch,
                         uint64_t *n_wrapper_mismatch, double *max_ulp_err) {
  uint64_t und = 0, bad = 0, wbad = 0;
  double worst = 0.0;
  for (uint32_t b = bits_begin; b < bits_end; ++b) {
    float x;
    memcpy(&x, &b, 4);
    bool flag;
    const double e = gsx::exp_reduced((double)x, &flag);
    const double ref = exp((double)x);
    const double err = fabs(e - ref) / (ref * 2.220446049250313e-16);
    if (err > worst) worst = err;
    if (flag) ++und;
    else if ((float)e != (float)ref) ++bad;
    if (gsx::exp_f32_via_f64(x) != (float)ref) ++wbad;
  }
  *n_undecided = und;
  *n_mismatch = bad;
  *n_wrapper_mismatch = wbad;
  *max_ulp_err = worst;
}

// ---- gsx_thresholds.h: `sqrtf(x) < t`  <=>  `x <= sqrt_lt_threshold(t)` ------------------------------------------------
#include "../../gradslam_b200/csrc/gsx_thresholds.h"

// scans `span` floats on either side of the threshold plus `n_random` bit patterns from a fixed LCG; returns mismatches
extern "C" uint64_t sqrt_threshold_scan(float t, int span, uint64_t n_random, float *threshold_out) {
  const float thr = gsx::sqrt_lt_threshold(t);
  *threshold_out = thr;
  uint64_t bad = 0;
  auto check = [&](float x) {
    if (x != x || x < 0.0f) return;
    if ((sqrtf(x) < t) != (x <= thr)) ++bad;
  };
  if (thr >= 0.0f) {
    float lo = thr, hi = thr;
    for (int i = 0; i < span; ++i) {
      check(lo);
      check(hi);
      lo = nextafterf(lo, -INFINITY);
      hi = nextafterf(hi, INFINITY);
    }
  }
  uint64_t state = 0x9E3779B97F4A7C15ull;
  for (uint64_t i = 0; i < n_random; ++i) {
    state = state * 6364136223846793005ull + 1442695040888963407ull;
    uint32_t b = (uint32_t)(state >> 33) & 0x7FFFFFFFu;
    float x;
    memcpy(&x, &b, 4);
    check(x);
  }
  check(0.0f);
  check(FLT_MAX);
  check(INFINITY);
  return bad;
}
