// libgsx: error reporting, version, and the whole-sequence drivers that chain the kernels without
// returning to Python between frames.
#include <cstdarg>
#include <cstdio>

#include "gsx_common.cuh"
#include "../../include/gsx.h"

namespace gsx {
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace gsx

extern "C" int gsx_version(void) { return GSX_VERSION; }
extern "C" const char *gsx_last_error(void) { return gsx::g_err; }

extern "C" int gsx_pointfusion_sequence_gt(float *map_points, float *map_normals, float *map_colors,
                                           float *map_ccounts, int32_t *counts, int64_t capacity,
                                           int64_t max_count0, const float *depth, const float *rgb,
                                           const float *intrinsics, const float *poses, int B, int L,
                                           int s_begin, int s_end, int H, int W, float dist_th, float dot_th,
                                           double sigma, float *scratch_maps,
                                           void *workspace, uint32_t epoch0, int32_t *overflow_flag,
                                           void *stream) {
  GSX_CHECK_ARG(B >= 0 && L >= 0 && H >= 2 && W >= 2, "gsx_pointfusion_sequence_gt: bad extents");
  GSX_CHECK_ARG(0 <= s_begin && s_begin <= s_end && s_end <= L, "gsx_pointfusion_sequence_gt: bad frame range");
  GSX_CHECK_ARG(counts && depth && rgb && intrinsics && poses, "gsx_pointfusion_sequence_gt: null pointer");
  (void)scratch_maps;
  const int64_t P = (int64_t)H * W;
  for (int s = s_begin; s < s_end; ++s) {
    int32_t *cin = counts + (int64_t)(s & 1) * B;
    int32_t *cout = counts + (int64_t)((s + 1) & 1) * B;
    int64_t max_count = max_count0 + (int64_t)(s - s_begin) * P;
    if (max_count > capacity) max_count = capacity;
    int rc = gsx_fusion_project_select(map_points, map_normals, map_ccounts, cin, capacity, max_count,
                                       poses + (int64_t)s * 16, (int64_t)L * 16, intrinsics, 16,
                                       depth + (int64_t)s * P, (int64_t)L * P, nullptr, nullptr, B, H, W, dist_th,
                                       dot_th, workspace, stream);
    if (rc) return rc;
    rc = gsx_fusion_merge_append(map_points, map_normals, map_colors, map_ccounts, cin, cout, capacity,
                                 depth + (int64_t)s * P, (int64_t)L * P, rgb + (int64_t)s * P * 3,
                                 (int64_t)L * P * 3, intrinsics, 16, poses + (int64_t)s * 16, (int64_t)L * 16,
                                 nullptr, nullptr, B, H, W, sigma, workspace, epoch0 + (uint32_t)(s - s_begin),
                                 overflow_flag, stream);
    if (rc) return rc;
  }
  return 0;
}
