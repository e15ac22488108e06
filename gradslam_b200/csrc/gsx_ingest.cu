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

}  // namespace gsx

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
