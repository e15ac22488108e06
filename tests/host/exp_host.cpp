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
