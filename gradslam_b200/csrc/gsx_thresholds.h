// Host-side helper: replaces the decision `sqrtf(x) < t` by `x <= sqrt_lt_threshold(t)`.
// The correctly rounded float32 square root is monotonic non-decreasing, so {x : sqrtf(x) < t} is a down-set of the
// floats and has a largest element; it is found here by stepping from float(t*t).  Used by the GSX_K2_FASTTEST variant
// of K2 (gsx_fusion.cu) and checked against sqrtf on the host by tests/test_exp_host.py::test_sqrt_threshold.
#pragma once
#include <cfloat>
#include <cmath>

namespace gsx {

// largest float x >= 0 with sqrtf(x) < t; -1 if there is none (t <= 0 or NaN)
inline float sqrt_lt_threshold(float t) {
  if (!(t > 0.0f)) return -1.0f;
  if (t == INFINITY) return FLT_MAX;
  float x = (float)((double)t * (double)t);
  if (x > FLT_MAX) x = FLT_MAX;
  while (x > 0.0f && !(sqrtf(x) < t)) x = nextafterf(x, -INFINITY);
  while (x < FLT_MAX && sqrtf(nextafterf(x, INFINITY)) < t) x = nextafterf(x, INFINITY);
  return (sqrtf(x) < t) ? x : -1.0f;
}

}  // namespace gsx
