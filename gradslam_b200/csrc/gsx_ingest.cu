// Dataset-native ingest: 8-bit colour + 16-bit depth (what ICL-NUIM / TUM / ScanNet store on disk) -> the float32
// tensors the hot path consumes, on the device.  Moves 5 bytes per pixel over PCIe instead of 16.
// Reference contract (gradslam/datasets/icl.py:467-513 and the same code in tum.py / scannet.py): colour is
// float(u8) (optionally / 255), depth is float32(float64(u16) / scaling_factor); both are reproduced bit for bit.
#include "gsx_common.cuh"
#include "../../include/gsx.h"

namespace gsx {

__global__ void __launch_bounds__(256) k_ingest_raw(const uint8_t *__restrict__ rgb, const uint16_t *__restrict__ depth,
                                                    int64_t n_pixels, double depth_div, int normalize, int vec_ok,
                                                    float *__restrict__ rgb_out, float *__restrict__ depth_out) {
  const int64_t quad = (int64_t)blockIdx.x * 256 + threadIdx.x;  // 4 pixels per thread
  const int64_t p0 = quad * 4;
  if (p0 >= n_pixels) return;
  const int n = (int)((n_pixels - p0) < 4 ? (n_pixels - p0) : 4);
  if (n == 4 && vec_ok) {
    // 12 colour bytes + 4 depth words in, 12 + 4 floats out, all 16-byte aligned when the buffers are
    const uint32_t *c = reinterpret_cast<const uint32_t *>(rgb + p0 * 3);
    const uint32_t w0 = __ldg(c), w1 = __ldg(c + 1), w2 = __ldg(c + 2);
    const uint2 dw = __ldg(reinterpret_cast<const uint2 *>(depth + p0));
    float f[12];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[i] = (float)((w0 >> (8 * i)) & 0xffu);
      f[4 + i] = (float)((w1 >> (8 * i)) & 0xffu);
      f[8 + i] = (float)((w2 >> (8 * i)) & 0xffu);
    }
    if (normalize)
#pragma unroll
      for (int i = 0; i < 12; ++i) f[i] = (float)((double)f[i] / 255.0);
    float4 *o = reinterpret_cast<float4 *>(rgb_out + p0 * 3);
    o[0] = make_float4(f[0], f[1], f[2], f[3]);
    o[1] = make_float4(f[4], f[5], f[6], f[7]);
    o[2] = make_float4(f[8], f[9], f[10], f[11]);
    const float d0 = (float)((double)(dw.x & 0xffffu) / depth_div), d1 = (float)((double)(dw.x >> 16) / depth_div);
    const float d2 = (float)((double)(dw.y & 0xffffu) / depth_div), d3 = (float)((double)(dw.y >> 16) / depth_div);
    *reinterpret_cast<float4 *>(depth_out + p0) = make_float4(d0, d1, d2, d3);
  } else {
    for (int i = 0; i < n; ++i) {
      for (int ch = 0; ch < 3; ++ch) {
        float v = (float)rgb[(p0 + i) * 3 + ch];
        if (normalize) v = (float)((double)v / 255.0);
        rgb_out[(p0 + i) * 3 + ch] = v;
      }
      depth_out[p0 + i] = (float)((double)depth[p0 + i] / depth_div);
    }
  }
}

// The rest of the loaders' output contract that is arithmetic rather than file parsing:
//   * intrinsics of resized frames: fx, cx *= w_ratio; fy, cy *= h_ratio in float32 (datasets/datautils.py:73-122);
//   * poses relative to the first frame of each sequence: T_s <- compose(inverse(T_0), T_s) (datasets/icl.py:515-533 =
//     geometryutils.relative_transformation with orthogonal_rotations=False: a GENERAL 4x4 inverse of T_0, then kornia's
//     compose_transformations on the rotation / translation blocks, bottom row forced to 0 0 0 1).
// One thread per matrix.  The inverse is Gauss-Jordan with partial pivoting in float32 (the reference's torch.inverse is
// LAPACK's LU: the results agree to float32 rounding, tests/test_gpu_ingest.py holds them to 1e-5).
__global__ void __launch_bounds__(64) k_scale_intrinsics(const float *__restrict__ K, int64_t n, int dim, float h_ratio,
                                                         float w_ratio, float *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * 64 + threadIdx.x;
  if (i >= n) return;
  const int sz = dim * dim;
  for (int e = 0; e < sz; ++e) out[i * sz + e] = K[i * sz + e];
  out[i * sz + 0] = K[i * sz + 0] * w_ratio;              // fx
  out[i * sz + dim + 1] = K[i * sz + dim + 1] * h_ratio;  // fy
  out[i * sz + 2] = K[i * sz + 2] * w_ratio;              // cx
  out[i * sz + dim + 2] = K[i * sz + dim + 2] * h_ratio;  // cy
}

__device__ inline bool invert4x4(const float *a, float *inv) {
  float m[4][8];
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) {
      m[r][c] = a[r * 4 + c];
      m[r][4 + c] = (r == c) ? 1.0f : 0.0f;
    }
  for (int col = 0; col < 4; ++col) {
    int piv = col;
    for (int r = col + 1; r < 4; ++r)
      if (fabsf(m[r][col]) > fabsf(m[piv][col])) piv = r;
    if (m[piv][col] == 0.0f) return false;
    if (piv != col)
      for (int c = 0; c < 8; ++c) {
        const float t = m[col][c];
        m[col][c] = m[piv][c];
        m[piv][c] = t;
      }
    const float d = 1.0f / m[col][col];
    for (int c = 0; c < 8; ++c) m[col][c] *= d;
    for (int r = 0; r < 4; ++r)
      if (r != col) {
        const float f = m[r][col];
        for (int c = 0; c < 8; ++c) m[r][c] -= f * m[col][c];
      }
  }
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) inv[r * 4 + c] = m[r][4 + c];
  return true;
}

__global__ void __launch_bounds__(64) k_poses_relative(const float *__restrict__ poses, int B, int L,
                                                       float *__restrict__ out, int32_t *singular) {
  const int i = blockIdx.x * 64 + threadIdx.x;
  if (i >= B * L) return;
  const int b = i / L;
  float t0[16], inv[16];
  for (int e = 0; e < 16; ++e) t0[e] = poses[(int64_t)b * L * 16 + e];
  if (!invert4x4(t0, inv)) {
    if (singular) *singular = 1;
    for (int e = 0; e < 16; ++e) out[(int64_t)i * 16 + e] = nanf("");
    return;
  }
  const float *t = poses + (int64_t)i * 16;
  float *o = out + (int64_t)i * 16;
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c)
      o[r * 4 + c] = (inv[r * 4 + 0] * t[0 * 4 + c] + inv[r * 4 + 1] * t[1 * 4 + c]) + inv[r * 4 + 2] * t[2 * 4 + c];
    o[r * 4 + 3] = ((inv[r * 4 + 0] * t[3] + inv[r * 4 + 1] * t[7]) + inv[r * 4 + 2] * t[11]) + inv[r * 4 + 3];
  }
  o[12] = o[13] = o[14] = 0.0f;
  o[15] = 1.0f;
}

}  // namespace gsx

extern "C" int gsx_ingest_calibration(const float *intrinsics, int64_t n_intrinsics, int intrinsics_dim, double h_ratio,
                                      double w_ratio, float *intrinsics_out, const float *poses, int B, int L,
                                      float *poses_out, int32_t *singular_flag, void *stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (intrinsics_out && n_intrinsics > 0) {
    GSX_CHECK_ARG(intrinsics && (intrinsics_dim == 3 || intrinsics_dim == 4), "gsx_ingest_calibration: bad intrinsics");
    gsx::k_scale_intrinsics<<<(unsigned)((n_intrinsics + 63) / 64), 64, 0, s>>>(intrinsics, n_intrinsics, intrinsics_dim,
                                                                               (float)h_ratio, (float)w_ratio,
                                                                               intrinsics_out);
  }
  if (poses_out && B > 0 && L > 0) {
    GSX_CHECK_ARG(poses && poses != poses_out, "gsx_ingest_calibration: poses must be given and must not alias the output");
    gsx::k_poses_relative<<<(unsigned)((B * L + 63) / 64), 64, 0, s>>>(poses, B, L, poses_out, singular_flag);
  }
  GSX_CHECK_LAUNCH("gsx_ingest_calibration");
  return 0;
}

extern "C" int gsx_ingest_raw(const uint8_t *rgb_u8, const uint16_t *depth_u16, int64_t n_pixels,
                              double depth_scaling_factor, int normalize_color, float *rgb_out, float *depth_out,
                              void *stream) {
  if (n_pixels == 0) return 0;
  GSX_CHECK_ARG(rgb_u8 && depth_u16 && rgb_out && depth_out, "gsx_ingest_raw: null pointer");
  GSX_CHECK_ARG(n_pixels > 0 && depth_scaling_factor != 0.0, "gsx_ingest_raw: bad arguments");
  // vector path needs 4-byte aligned colour, 8-byte aligned depth and 16-byte aligned outputs; otherwise scalar
  const int vec_ok = (((uintptr_t)rgb_u8 & 3) == 0) && (((uintptr_t)depth_u16 & 7) == 0) &&
                     ((((uintptr_t)rgb_out | (uintptr_t)depth_out) & 15) == 0);
  const int64_t quads = (n_pixels + 3) / 4;
  gsx::k_ingest_raw<<<(unsigned)((quads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      rgb_u8, depth_u16, n_pixels, depth_scaling_factor, normalize_color, vec_ok, rgb_out, depth_out);
  GSX_CHECK_LAUNCH("gsx_ingest_raw");
  return 0;
}
