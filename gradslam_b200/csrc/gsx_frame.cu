// K1: depth -> local/global vertex and normal maps.  One thread per pixel; per-image constants (inverse
// intrinsics, pose) are computed once per CTA into shared memory; outputs are staged through shared memory
// so every map is written with coalesced 16-byte stores.
// Reference op chain: gradslam/structures/rgbdimages.py:643-762 (see include/gsx.h).
#include "gsx_common.cuh"
#include "../../include/gsx.h"

namespace gsx {

constexpr int kTile = 256;  // pixels per CTA; 256*3 floats = 3072 B per map tile (16-byte aligned)

struct FrameArgs {
  const float *depth;
  int64_t depth_bstride;
  const float *K;
  int64_t K_bstride;
  const float *poses;  // may be null
  int64_t pose_bstride;
  int B, L, H, W;
  float *out[4];  // vertex, normal, gvertex, gnormal (any may be null)
};

__global__ void __launch_bounds__(kTile) k_backproject_normals(FrameArgs a) {
  __shared__ __align__(16) float stage[4][kTile * 3];
  __shared__ KInv s_k;
  __shared__ Rigid s_pose;
  const int img = blockIdx.y;  // b*L + l
  const int b = img / a.L, l = img - b * a.L;
  const int P = a.H * a.W;
  if (threadIdx.x == 0) s_k = load_kinv(a.K + b * a.K_bstride);
  if (threadIdx.x == 32 && a.poses) s_pose = load_rigid(a.poses + b * a.pose_bstride + (int64_t)l * 16);
  __syncthreads();
  const KInv k = s_k;
  const int tile0 = blockIdx.x * kTile;
  const int pix = tile0 + threadIdx.x;
  const bool want_local = a.out[0] || a.out[1], want_global = a.out[2] || a.out[3];
  if (pix < P) {
    const float *dimg = a.depth + b * a.depth_bstride + (int64_t)l * P;
    const int h = pix / a.W, w = pix - h * a.W;
    const FrameSample f = frame_sample<true>(dimg, k, a.poses ? &s_pose : nullptr, h, w, a.H, a.W);
    const int t3 = threadIdx.x * 3;
    if (want_local) {
      stage[0][t3] = f.v.x; stage[0][t3 + 1] = f.v.y; stage[0][t3 + 2] = f.v.z;
      stage[1][t3] = f.n.x; stage[1][t3 + 1] = f.n.y; stage[1][t3 + 2] = f.n.z;
    }
    if (want_global) {
      stage[2][t3] = f.gv.x; stage[2][t3 + 1] = f.gv.y; stage[2][t3 + 2] = f.gv.z;
      stage[3][t3] = f.gn.x; stage[3][t3 + 1] = f.gn.y; stage[3][t3 + 2] = f.gn.z;
    }
  }
  __syncthreads();
  const int remaining = P - tile0;
  const int nfloat = (remaining < kTile ? remaining : kTile) * 3;
  const int64_t obase = ((int64_t)img * P + tile0) * 3;
  // every image starts 16-byte aligned only if P*3 floats is a multiple of 4
  const bool vec_ok = (nfloat == kTile * 3) && ((obase & 3) == 0);
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    float *out = a.out[m];
    if (!out) continue;
    float *dst = out + obase;
    if (vec_ok) {
      if (threadIdx.x < kTile * 3 / 4)
        reinterpret_cast<float4 *>(dst)[threadIdx.x] = reinterpret_cast<const float4 *>(stage[m])[threadIdx.x];
    } else {
      for (int j = threadIdx.x; j < nfloat; j += kTile) dst[j] = stage[m][j];
    }
  }
}

int launch_backproject(const FrameArgs &a, cudaStream_t stream) {
  const int64_t P = (int64_t)a.H * a.W;
  if ((int64_t)a.B * a.L * P == 0) return 0;
  const int64_t bx = (P + kTile - 1) / kTile;
  k_backproject_normals<<<dim3((unsigned)bx, (unsigned)(a.B * a.L)), kTile, 0, stream>>>(a);
  GSX_CHECK_LAUNCH("gsx_backproject_normals_fwd");
  return 0;
}

}  // namespace gsx

extern "C" int gsx_backproject_normals_fwd(const float *depth, int64_t depth_bstride, const float *intrinsics,
                                           int64_t K_bstride, const float *poses, int64_t pose_bstride, int B,
                                           int L, int H, int W, float *vertex, float *normal, float *gvertex,
                                           float *gnormal, void *stream) {
  GSX_CHECK_ARG(depth && intrinsics, "gsx_backproject_normals_fwd: null depth/intrinsics");
  GSX_CHECK_ARG(B >= 0 && L >= 0 && H >= 2 && W >= 2, "gsx_backproject_normals_fwd: need H,W >= 2 (got %d x %d)", H, W);
  GSX_CHECK_ARG((int64_t)B * L <= 65535 && (int64_t)H * W < (1ll << 30), "gsx_backproject_normals_fwd: extents too large");
  float *outs[4] = {vertex, normal, gvertex, gnormal};
  for (float *o : outs)
    GSX_CHECK_ARG(((uintptr_t)o & 15) == 0, "gsx_backproject_normals_fwd: outputs must be 16-byte aligned");
  gsx::FrameArgs a{depth, depth_bstride, intrinsics, K_bstride, poses, pose_bstride, B, L, H, W,
                   {vertex, normal, gvertex, gnormal}};
  return gsx::launch_backproject(a, (cudaStream_t)stream);
}
