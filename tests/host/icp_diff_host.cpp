// Host instantiation of the scalar-generic K7 templates (gradslam_b200/csrc/gsx_icp_diff.cuh) for the CPU test suite:
// forward values and dual-number Jacobians, so tests/test_host_logic.py can compare them with the oracle's autograd.
#include "../../gradslam_b200/csrc/gsx_icp_diff.cuh"

using namespace gsx;

extern "C" void host_solve(const float *in, float *out, float *jac) {  // jac[j * 22 + k] = d out[k] / d in[j]
  solve_step_t<float>(in, out);
  for (int j = 0; j < kSolveIn; ++j) {
    Dual din[kSolveIn], dout[kSolveOut];
    for (int i = 0; i < kSolveIn; ++i) din[i] = mk(in[i], i == j ? 1.0f : 0.0f);
    solve_step_t<Dual>(din, dout);
    for (int k = 0; k < kSolveOut; ++k) jac[j * kSolveOut + k] = dout[k].d;
  }
}

extern "C" void host_update(const float *in, int mode, float lambda_max, float B, float B2, float nu, float *out,
                            float *jac) {  // jac[j * 33 + k]
  const UpdateParams u{mode, 1.0f / lambda_max, lambda_max, B, B2, 1.0f / nu};
  update_step_t<float>(in, out, u);
  for (int j = 0; j < kUpdateIn; ++j) {
    Dual din[kUpdateIn], dout[kUpdateOut];
    for (int i = 0; i < kUpdateIn; ++i) din[i] = mk(in[i], i == j ? 1.0f : 0.0f);
    update_step_t<Dual>(din, dout, u);
    for (int k = 0; k < kUpdateOut; ++k) jac[j * kUpdateOut + k] = dout[k].d;
  }
}
