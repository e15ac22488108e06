// Scalar-generic arithmetic of the small ICP steps (K7): damped 6x6 solve, se3_exp, LM / gradLM update.
// Written once as templates; instantiated on float for the forward kernels and on dual numbers for the backward
// kernels (gsx_icp_diff.cu).  Host-callable too, so tests/test_host_logic.py can check the dual-number Jacobians
// against finite differences without a GPU.
#pragma once
#include <math.h>
#ifdef __CUDACC__
#define GSX_HD __host__ __device__
#else
#define GSX_HD
#endif

namespace gsx {

struct Dual {
  float v, d;
};
GSX_HD inline Dual mk(float v, float d = 0.0f) { return Dual{v, d}; }
GSX_HD inline Dual operator+(Dual a, Dual b) { return Dual{a.v + b.v, a.d + b.d}; }
GSX_HD inline Dual operator-(Dual a, Dual b) { return Dual{a.v - b.v, a.d - b.d}; }
GSX_HD inline Dual operator*(Dual a, Dual b) { return Dual{a.v * b.v, a.d * b.v + a.v * b.d}; }
GSX_HD inline Dual operator/(Dual a, Dual b) {
  const float q = a.v / b.v;
  return Dual{q, (a.d - q * b.d) / b.v};
}
GSX_HD inline Dual operator-(Dual a) { return Dual{-a.v, -a.d}; }
GSX_HD inline Dual operator+(float a, Dual b) { return Dual{a + b.v, b.d}; }
GSX_HD inline Dual operator+(Dual a, float b) { return Dual{a.v + b, a.d}; }
GSX_HD inline Dual operator-(float a, Dual b) { return Dual{a - b.v, -b.d}; }
GSX_HD inline Dual operator-(Dual a, float b) { return Dual{a.v - b, a.d}; }
GSX_HD inline Dual operator*(float a, Dual b) { return Dual{a * b.v, a * b.d}; }
GSX_HD inline Dual operator*(Dual a, float b) { return Dual{a.v * b, a.d * b}; }
GSX_HD inline Dual operator/(float a, Dual b) {
  const float q = a / b.v;
  return Dual{q, (-q * b.d) / b.v};
}
GSX_HD inline Dual operator/(Dual a, float b) { return Dual{a.v / b, a.d / b}; }

GSX_HD inline float val(float a) { return a; }
GSX_HD inline float val(Dual a) { return a.v; }
GSX_HD inline float t_sqrt(float a) { return sqrtf(a); }
GSX_HD inline Dual t_sqrt(Dual a) {
  const float r = sqrtf(a.v);
  return Dual{r, (r > 0.0f) ? a.d / (2.0f * r) : 0.0f};
}
GSX_HD inline float t_sin(float a) { return sinf(a); }
GSX_HD inline Dual t_sin(Dual a) { return Dual{sinf(a.v), cosf(a.v) * a.d}; }
GSX_HD inline float t_cos(float a) { return cosf(a); }
GSX_HD inline Dual t_cos(Dual a) { return Dual{cosf(a.v), -sinf(a.v) * a.d}; }
GSX_HD inline float t_exp(float a) { return expf(a); }
GSX_HD inline Dual t_exp(Dual a) {
  const float e = expf(a.v);
  return Dual{e, e * a.d};
}
GSX_HD inline float t_pow(float a, float p) { return powf(a, p); }
GSX_HD inline Dual t_pow(Dual a, float p) {
  const float r = powf(a.v, p);
  return Dual{r, p * (r / a.v) * a.d};
}
// clamp with the sub-gradient torch.clamp uses (1 inside [lo, hi], 0 outside)
GSX_HD inline float t_clamp(float a, float lo, float hi) { return fminf(fmaxf(a, lo), hi); }
GSX_HD inline Dual t_clamp(Dual a, float lo, float hi) {
  return Dual{fminf(fmaxf(a.v, lo), hi), (a.v >= lo && a.v <= hi) ? a.d : 0.0f};
}
template <class S>
GSX_HD inline S lit(float v);
template <>
GSX_HD inline float lit<float>(float v) { return v; }
template <>
GSX_HD inline Dual lit<Dual>(float v) { return Dual{v, 0.0f}; }

// ---- se3_exp (se3utils.py:77-115): xi = (v, omega) -> 4x4; for ||omega|| < 1e-6 both R and V are I + hat(omega) ------
template <class S>
GSX_HD void se3_exp_t(const S *xi, S *T) {
  const S vx = xi[0], vy = xi[1], vz = xi[2], wx = xi[3], wy = xi[4], wz = xi[5];
  const S zero = lit<S>(0.0f);
  const S Wm[9] = {zero, -wz, wy, wz, zero, -wx, -wy, wx, zero};
  const S theta = t_sqrt((wx * wx + wy * wy) + wz * wz);
  S R[9], V[9];
  if (val(theta) < 1e-6f) {
    for (int i = 0; i < 9; ++i) {
      const float I = (i % 4 == 0) ? 1.0f : 0.0f;
      R[i] = I + Wm[i];
      V[i] = I + Wm[i];
    }
  } else {
    const S s = t_sin(theta), c = t_cos(theta);
    S W2[9];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        S acc = Wm[i * 3 + 0] * Wm[0 * 3 + j];
        for (int k = 1; k < 3; ++k) acc = acc + Wm[i * 3 + k] * Wm[k * 3 + j];
        W2[i * 3 + j] = acc;
      }
    const S Ac = s / theta;
    const S Bc = (1.0f - c) / (theta * theta);
    const S Cc = (theta - s) / ((theta * theta) * theta);
    for (int i = 0; i < 9; ++i) {
      const float I = (i % 4 == 0) ? 1.0f : 0.0f;
      R[i] = (I + Ac * Wm[i]) + Bc * W2[i];
      V[i] = (I + Bc * Wm[i]) + Cc * W2[i];
    }
  }
  for (int i = 0; i < 3; ++i) {
    T[i * 4 + 0] = R[i * 3 + 0];
    T[i * 4 + 1] = R[i * 3 + 1];
    T[i * 4 + 2] = R[i * 3 + 2];
    T[i * 4 + 3] = (V[i * 3 + 0] * vx + V[i * 3 + 1] * vy) + V[i * 3 + 2] * vz;
  }
  T[12] = zero; T[13] = zero; T[14] = zero; T[15] = lit<S>(1.0f);
}

template <class S>
GSX_HD void mat4_mul_t(const S *A, const S *B, S *C) {
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      S acc = A[i * 4 + 0] * B[0 * 4 + j];
      for (int k = 1; k < 4; ++k) acc = acc + A[i * 4 + k] * B[k * 4 + j];
      C[i * 4 + j] = acc;
    }
}

// ---- K7a: damped 6x6 solve + se3_exp.  in = 28 sums (21 upper-triangular A^T A, 6 A^T b, r^T r) + damp ---------------
constexpr int kSolveIn = 29, kSolveOut = 22;  // out = xi (6) + dT (16)
template <class S>
GSX_HD void solve_step_t(const S *in, S *out) {
  S M[6][12];
  int k = 0;
  for (int i = 0; i < 6; ++i)
    for (int j = i; j < 6; ++j) {
      M[i][j] = in[k];
      M[j][i] = in[k];
      ++k;
    }
  const S damp = in[28];
  for (int i = 0; i < 6; ++i) {
    M[i][i] = M[i][i] + damp;
    for (int j = 0; j < 6; ++j) M[i][6 + j] = lit<S>((i == j) ? 1.0f : 0.0f);
  }
  // Gauss-Jordan inversion with partial pivoting (torch.inverse at icputils.py:86), then x = inv * A^T b
  for (int c = 0; c < 6; ++c) {
    int piv = c;
    float mx = fabsf(val(M[c][c]));
    for (int r = c + 1; r < 6; ++r)
      if (fabsf(val(M[r][c])) > mx) {
        mx = fabsf(val(M[r][c]));
        piv = r;
      }
    if (piv != c)
      for (int j = 0; j < 12; ++j) {
        const S t = M[c][j];
        M[c][j] = M[piv][j];
        M[piv][j] = t;
      }
    const S inv = 1.0f / M[c][c];
    for (int j = 0; j < 12; ++j) M[c][j] = M[c][j] * inv;
    for (int r = 0; r < 6; ++r) {
      if (r == c) continue;
      const S f = M[r][c];
      for (int j = 0; j < 12; ++j) M[r][j] = M[r][j] - f * M[c][j];
    }
  }
  for (int i = 0; i < 6; ++i) {
    S acc = lit<S>(0.0f);
    for (int j = 0; j < 6; ++j) acc = acc + M[i][6 + j] * in[21 + j];
    out[i] = acc;
  }
  se3_exp_t<S>(out, out + 6);
}

// ---- K7b: LM accept / reject or gradLM gates, applied step, pose accumulation ----------------------------------------
// in = xi (6), err, new_err, damp, T (16); out = new damp, applied step dT (16), new T = dT * T (16)
constexpr int kUpdateIn = 25, kUpdateOut = 33;
struct UpdateParams {
  int mode;  // 0 = LM (point_to_plane_ICP), 1 = gradLM (point_to_plane_gradICP)
  float lambda_min, lambda_max, B, B2, inv_nu;
};
template <class S>
GSX_HD void update_step_t(const S *in, S *out, const UpdateParams &u) {
  const S *xi = in, err = in[6], new_err = in[7], damp = in[8];
  const S *T = in + 9;
  S *dT = out + 1, *Tn = out + 17;
  if (u.mode == 0) {
    if (val(new_err) < val(err)) {  // trust region: accept the step
      se3_exp_t<S>(xi, dT);
      out[0] = damp / 2.0f;
      mat4_mul_t<S>(dT, T, Tn);
    } else {
      for (int i = 0; i < 16; ++i) {
        dT[i] = lit<S>((i % 5 == 0) ? 1.0f : 0.0f);
        Tn[i] = T[i];
      }
      out[0] = damp * 2.0f;
    }
  } else {
    const S diff = t_clamp(new_err - err, -70.0f, 70.0f);
    const S gate = u.lambda_min + (u.lambda_max - u.lambda_min) / (1.0f + t_exp(-(u.B * diff)));
    out[0] = damp * gate;
    const S sig = 1.0f / t_pow(1.0f + t_exp(-(u.B2 * diff)), u.inv_nu);
    S xs[6];
    for (int i = 0; i < 6; ++i) xs[i] = sig * xi[i];
    se3_exp_t<S>(xs, dT);
    mat4_mul_t<S>(dT, T, Tn);
  }
}

}  // namespace gsx
