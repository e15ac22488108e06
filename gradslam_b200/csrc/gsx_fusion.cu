// Fused PointFusion map update for sm_100a.
//   k_project_select  (K2+K3)  one thread per map point: project into the live camera, frustum / distance /
//                              normal tests, then a 128-bit atomic arg-min per pixel on the key
//                              (1/ccount, ray distance, point index).
//   k_merge_append    (K4)     one thread per pixel: confidence-weighted merge of the selected map point,
//                              or stable append of unmatched valid pixels (single-pass decoupled look-back
//                              scan, row-major order per batch element).  No float atomics anywhere.
// Reference op chains: gradslam/slam/fusionutils.py:198-722 (see include/gsx.h).
#include "gsx_common.cuh"
#include "../../include/gsx.h"

namespace gsx {

constexpr int kBlock = 256;

// ---- workspace layout -----------------------------------------------------------------------------------
//   [0, B*P*16)                       U128 best[B][P]     complemented arg-min records (0 = empty)
//   then  uint64 tile_state[B][T]     (epoch<<34 | flag<<32 | value), T = ceil(P / kBlock)
//   then  uint32 ticket[B]            dynamic tile ids (monotonic; tile = ticket - (epoch-1)*T)
struct Workspace {
  U128 *best;
  unsigned long long *tile_state;
  unsigned int *ticket;
  int tiles;
};

__host__ __device__ inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

inline Workspace carve(void *ws, int B, int H, int W) {
  const int64_t P = (int64_t)H * W;
  Workspace w;
  w.tiles = (int)((P + kBlock - 1) / kBlock);
  char *p = (char *)ws;
  w.best = (U128 *)p;
  p += align_up(B * P * 16, 256);
  w.tile_state = (unsigned long long *)p;
  p += align_up((int64_t)B * w.tiles * 8, 256);
  w.ticket = (unsigned int *)p;
  return w;
}

inline int64_t workspace_bytes(int B, int H, int W) {
  const int64_t P = (int64_t)H * W;
  const int64_t tiles = (P + kBlock - 1) / kBlock;
  return align_up(B * P * 16, 256) + align_up(B * tiles * 8, 256) + align_up((int64_t)B * 4, 256);
}

// ---- K2 + K3 ------------------------------------------------------------------------------------------
struct ProjectArgs {
  const float *pts, *nrm, *cc;
  const int32_t *counts;
  int64_t cap;
  const float *poses;
  int64_t pose_bstride;
  const float *K;
  int64_t K_bstride;
  const float *gv, *gn;  // (B,H,W,3)
  int B, H, W;
  float dist_th, dot_th, u_hi, v_hi;  // u_hi = float(W - 0.999), v_hi = float(H - 0.999)
  U128 *best;
};

__global__ void __launch_bounds__(kBlock) k_project_select(ProjectArgs a) {
  const int b = blockIdx.y;
  const int count = a.counts[b];
  const int64_t stride = (int64_t)gridDim.x * kBlock;
  int64_t n = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (n >= count) return;
  const Rigid Tinv = rigid_inverse(load_rigid(a.poses + b * a.pose_bstride));
  const float *K = a.K + b * a.K_bstride;
  float k[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) k[i] = __ldg(K + i);
  const int64_t P = (int64_t)a.H * a.W;
  const float *pts = a.pts + (int64_t)b * a.cap * 3;
  const float *nrm = a.nrm + (int64_t)b * a.cap * 3;
  const float *cc = a.cc + (int64_t)b * a.cap;
  const float *gv = a.gv + (int64_t)b * P * 3;
  const float *gn = a.gn + (int64_t)b * P * 3;
  U128 *best = a.best + (int64_t)b * P;
  for (; n < count; n += stride) {
    const float px = __ldg(pts + n * 3), py = __ldg(pts + n * 3 + 1), pz = __ldg(pts + n * 3 + 2);
    // world -> camera (pointclouds.py:526-573), then pinhole projection with the 4x4 K on the homogeneous
    // point (projutils.py:92-238): z == 0 divides by 1.
    const float3 q = rigid_apply(Tinv, px, py, pz);
    const float hx = ((k[0] * q.x + k[1] * q.y) + k[2] * q.z) + k[3];
    const float hy = ((k[4] * q.x + k[5] * q.y) + k[6] * q.z) + k[7];
    const float hz = ((k[8] * q.x + k[9] * q.y) + k[10] * q.z) + k[11];
    const float den = (hz != 0.0f) ? hz : 1.0f;
    const float u = hx / den, v = hy / den;
    // fusionutils.py:259-266
    const bool in_frame = (u > -1e-3f) && (u < a.u_hi) && (v > -1e-3f) && (v < a.v_hi) && (q.z > 0.0f);
    if (!in_frame) continue;
    // round-half-even like torch.round, then clamp (fusionutils.py:267-274)
    int w = (int)rintf(u), h = (int)rintf(v);
    w = min(max(w, 0), a.W - 1);
    h = min(max(h, 0), a.H - 1);
    const int64_t pix = (int64_t)h * a.W + w;
    const float fx = __ldg(gv + pix * 3), fy = __ldg(gv + pix * 3 + 1), fz = __ldg(gv + pix * 3 + 2);
    // are_points_close (fusionutils.py:130): ||frame - map|| < dist_th
    const float dx = fx - px, dy = fy - py, dz = fz - pz;
    const float d2 = (dx * dx + dy * dy) + dz * dz;
    if (!(sqrtf(d2) < a.dist_th)) continue;
    // are_normals_similar (fusionutils.py:187-195): n_frame . n_map > dot_th
    const float nx = __ldg(gn + pix * 3), ny = __ldg(gn + pix * 3 + 1), nz = __ldg(gn + pix * 3 + 2);
    const float mx = __ldg(nrm + n * 3), my = __ldg(nrm + n * 3 + 1), mz = __ldg(nrm + n * 3 + 2);
    const float dot = (nx * mx + ny * my) + nz * mz;
    if (!(dot > a.dot_th)) continue;
    // sort key of find_best_unique_correspondences (fusionutils.py:491-517): 1/(cc+1e-20), then the squared
    // distance (map - frame)^2 (same value as d2: the squares are sign-independent), then n.
    const float inv_cc = 1.0f / (__ldg(cc + n) + 1e-20f);
    // positive floats order like their bit patterns; flip negatives so the order stays total.
    unsigned int kb = __float_as_uint(inv_cc);
    kb = (kb & 0x80000000u) ? ~kb : (kb | 0x80000000u);
    const unsigned int rb = __float_as_uint(d2) | 0x80000000u;  // d2 >= 0
    const unsigned long long hi = ((unsigned long long)kb << 32) | rb;
    atomic_min_key128(best + pix, hi, (unsigned long long)n);
  }
}

// ---- K4 -------------------------------------------------------------------------------------------------
struct MergeArgs {
  float *pts, *nrm, *col, *cc;
  const int32_t *counts_in;
  int32_t *counts_out;
  int64_t cap;
  const float *depth;
  int64_t depth_bstride;
  const float *rgb;
  int64_t rgb_bstride;
  const float *K;
  int64_t K_bstride;
  const float *gv, *gn;
  int B, H, W;
  float two_sigma_sq;
  Workspace ws;
  unsigned int epoch;
  int32_t *overflow;
};

constexpr unsigned long long kFlagAgg = 1ull, kFlagPrefix = 2ull;

__device__ __forceinline__ unsigned long long pack_state(unsigned int epoch, unsigned long long flag, unsigned int value) {
  return ((unsigned long long)epoch << 34) | (flag << 32) | value;
}

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(kBlock) k_merge_append(MergeArgs a) {
  __shared__ int s_tile;
  __shared__ int s_warp_sums[kBlock / 32];
  __shared__ int s_excl;
  const int b = blockIdx.y;
  const int T = a.ws.tiles;
  if (threadIdx.x == 0) {
    // dynamic tile id: tiles start in ticket order, so every predecessor of a running tile is running or done
    const unsigned int t = atomicAdd(a.ws.ticket + b, 1u);
    s_tile = (int)(t - (a.epoch - 1u) * (unsigned int)T);
  }
  __syncthreads();
  const int tile = s_tile;
  const int64_t P = (int64_t)a.H * a.W;
  const int64_t pix = (int64_t)tile * kBlock + threadIdx.x;
  const bool in_img = pix < P;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  U128 rec{0ull, 0ull};
  float d = 0.0f;
  if (in_img) {
    U128 *slot = a.ws.best + (int64_t)b * P + pix;
    rec = *slot;
    if (rec.lo | rec.hi) *slot = U128{0ull, 0ull};  // leave the workspace clean for the next frame
    d = __ldg(a.depth + b * a.depth_bstride + pix);
  }
  const bool matched = (rec.lo | rec.hi) != 0ull;
  const bool valid = d > 0.0f;
  const bool is_new = in_img && valid && !matched;

  // block-wide exclusive scan of is_new
  const unsigned int ballot = __ballot_sync(0xffffffffu, is_new);
  const int warp_excl = __popc(ballot & ((1u << lane) - 1u));
  if (lane == 0) s_warp_sums[warp] = __popc(ballot);
  __syncthreads();
  int block_excl = 0, block_total = 0;
#pragma unroll
  for (int i = 0; i < kBlock / 32; ++i) {
    const int s = s_warp_sums[i];
    if (i < warp) block_excl += s;
    block_total += s;
  }
  unsigned long long *state = a.ws.tile_state + (int64_t)b * T;
  if (threadIdx.x == 0 && tile + 1 < T) st_release_u64(state + tile, pack_state(a.epoch, kFlagAgg, (unsigned)block_total));

  // per-pixel frame sample: alpha from the LOCAL vertex (fusionutils.py:657, 69-72)
  float alpha = 0.0f;
  float3 fp, fn, fc;
  if (in_img && (matched || is_new)) {
    const int h = (int)(pix / a.W), w = (int)(pix - (int64_t)h * a.W);
    const KInv k = load_kinv(a.K + b * a.K_bstride);
    const float3 v = backproject(k, (float)w, (float)h, d);
    const float s = (v.x * v.x + v.y * v.y) + v.z * v.z;
    alpha = fminf(fmaxf(expf((-s) / a.two_sigma_sq), 1e-7f), 1.01f);
    const float *gv = a.gv + ((int64_t)b * P + pix) * 3;
    const float *gn = a.gn + ((int64_t)b * P + pix) * 3;
    const float *c = a.rgb + b * a.rgb_bstride + pix * 3;
    fp = make_float3(__ldg(gv), __ldg(gv + 1), __ldg(gv + 2));
    fn = make_float3(__ldg(gn), __ldg(gn + 1), __ldg(gn + 2));
    fc = make_float3(__ldg(c), __ldg(c + 1), __ldg(c + 2));
  }
  float *pts = a.pts + (int64_t)b * a.cap * 3;
  float *nrm = a.nrm + (int64_t)b * a.cap * 3;
  float *col = a.col + (int64_t)b * a.cap * 3;
  float *cc = a.cc ? a.cc + (int64_t)b * a.cap : nullptr;

  if (matched && cc) {
    // confidence-weighted running mean (fusionutils.py:678-699); exactly one pixel owns this map row
    const int64_t n = (int64_t)(~rec.lo);
    const float c0 = cc[n];
    const float tot = c0 + alpha;
    const float inv = 1.0f / ((tot == 0.0f) ? 1.0f : tot);
    pts[n * 3 + 0] = ((c0 * pts[n * 3 + 0]) + (alpha * fp.x)) * inv;
    pts[n * 3 + 1] = ((c0 * pts[n * 3 + 1]) + (alpha * fp.y)) * inv;
    pts[n * 3 + 2] = ((c0 * pts[n * 3 + 2]) + (alpha * fp.z)) * inv;
    nrm[n * 3 + 0] = ((c0 * nrm[n * 3 + 0]) + (alpha * fn.x)) * inv;
    nrm[n * 3 + 1] = ((c0 * nrm[n * 3 + 1]) + (alpha * fn.y)) * inv;
    nrm[n * 3 + 2] = ((c0 * nrm[n * 3 + 2]) + (alpha * fn.z)) * inv;
    col[n * 3 + 0] = ((c0 * col[n * 3 + 0]) + (alpha * fc.x)) * inv;
    col[n * 3 + 1] = ((c0 * col[n * 3 + 1]) + (alpha * fc.y)) * inv;
    col[n * 3 + 2] = ((c0 * col[n * 3 + 2]) + (alpha * fc.z)) * inv;
    cc[n] = tot;
  }

  // decoupled look-back: exclusive prefix of new-point counts over preceding tiles of this element
  if (threadIdx.x == 0) {
    unsigned int excl = 0;
    for (int j = tile - 1; j >= 0; --j) {
      unsigned long long s;
      do {
        s = ld_acquire_u64(state + j);
      } while ((unsigned int)(s >> 34) != a.epoch);
      excl += (unsigned int)s;
      if (((s >> 32) & 3ull) == kFlagPrefix) break;
    }
    if (tile + 1 < T) st_release_u64(state + tile, pack_state(a.epoch, kFlagPrefix, excl + (unsigned)block_total));
    s_excl = (int)excl;
  }
  __syncthreads();
  const int64_t base = (int64_t)a.counts_in[b] + s_excl;
  if (is_new) {
    // append in row-major pixel order (fusionutils.py:702-720; pointclouds.py:1203-1235)
    const int64_t n = base + block_excl + warp_excl;
    if (n < a.cap) {
      pts[n * 3 + 0] = fp.x; pts[n * 3 + 1] = fp.y; pts[n * 3 + 2] = fp.z;
      nrm[n * 3 + 0] = fn.x; nrm[n * 3 + 1] = fn.y; nrm[n * 3 + 2] = fn.z;
      col[n * 3 + 0] = fc.x; col[n * 3 + 1] = fc.y; col[n * 3 + 2] = fc.z;
      if (cc) cc[n] = alpha;
    } else {
      *a.overflow = 1;
    }
  }
  if (tile == T - 1 && threadIdx.x == 0) {
    const int64_t total = base + block_total;
    a.counts_out[b] = (int32_t)(total < a.cap ? total : a.cap);
  }
}

int launch_project_select(const ProjectArgs &a, int64_t max_count, cudaStream_t stream) {
  if (a.B == 0 || max_count <= 0) return 0;
  int64_t bx = (max_count + kBlock - 1) / kBlock;
  const int64_t cap_blocks = (int64_t)kNumSMs * 8;  // grid-stride beyond 8 CTAs per SM
  if (bx * a.B > cap_blocks) bx = (cap_blocks + a.B - 1) / a.B;
  if (bx < 1) bx = 1;
  k_project_select<<<dim3((unsigned)bx, (unsigned)a.B), kBlock, 0, stream>>>(a);
  GSX_CHECK_LAUNCH("gsx_fusion_project_select");
  return 0;
}

int launch_merge_append(const MergeArgs &a, cudaStream_t stream) {
  if (a.B == 0) return 0;
  k_merge_append<<<dim3((unsigned)a.ws.tiles, (unsigned)a.B), kBlock, 0, stream>>>(a);
  GSX_CHECK_LAUNCH("gsx_fusion_merge_append");
  return 0;
}

}  // namespace gsx

using namespace gsx;

extern "C" int64_t gsx_fusion_workspace_bytes(int B, int H, int W) {
  if (B < 0 || H < 0 || W < 0) return -1;
  return workspace_bytes(B, H, W);
}

extern "C" int gsx_fusion_project_select(const float *map_points, const float *map_normals,
                                         const float *map_ccounts, const int32_t *counts, int64_t capacity,
                                         int64_t max_count, const float *poses, int64_t pose_bstride,
                                         const float *intrinsics, int64_t K_bstride, const float *gvertex,
                                         const float *gnormal, int B, int H, int W, float dist_th, float dot_th,
                                         void *workspace, void *stream) {
  GSX_CHECK_ARG(B >= 0 && H >= 2 && W >= 2, "gsx_fusion_project_select: bad extents B=%d H=%d W=%d", B, H, W);
  if (max_count <= 0 || B == 0) return 0;
  GSX_CHECK_ARG(map_points && map_normals && map_ccounts && counts, "gsx_fusion_project_select: null map pointer");
  GSX_CHECK_ARG(poses && intrinsics && gvertex && gnormal && workspace, "gsx_fusion_project_select: null frame pointer");
  GSX_CHECK_ARG(max_count <= capacity, "gsx_fusion_project_select: max_count %lld > capacity %lld",
                (long long)max_count, (long long)capacity);
  const Workspace ws = carve(workspace, B, H, W);
  ProjectArgs a{map_points, map_normals, map_ccounts, counts, capacity, poses, pose_bstride, intrinsics, K_bstride,
                gvertex, gnormal, B, H, W, dist_th, dot_th, (float)(W - 0.999), (float)(H - 0.999), ws.best};
  return launch_project_select(a, max_count, (cudaStream_t)stream);
}

extern "C" int gsx_fusion_merge_append(float *map_points, float *map_normals, float *map_colors,
                                       float *map_ccounts, const int32_t *counts_in, int32_t *counts_out,
                                       int64_t capacity, const float *depth, int64_t depth_bstride,
                                       const float *rgb, int64_t rgb_bstride, const float *intrinsics,
                                       int64_t K_bstride, const float *gvertex, const float *gnormal, int B,
                                       int H, int W, double sigma, void *workspace, uint32_t epoch,
                                       int32_t *overflow_flag, void *stream) {
  GSX_CHECK_ARG(B >= 0 && H >= 2 && W >= 2, "gsx_fusion_merge_append: bad extents B=%d H=%d W=%d", B, H, W);
  if (B == 0) return 0;
  GSX_CHECK_ARG(map_points && map_normals && map_colors && counts_in && counts_out,
                "gsx_fusion_merge_append: null map pointer");  // map_ccounts may be NULL (aggregation-only maps)
  GSX_CHECK_ARG(counts_in != counts_out, "gsx_fusion_merge_append: counts_in and counts_out must not alias");
  GSX_CHECK_ARG(depth && rgb && intrinsics && gvertex && gnormal && workspace && overflow_flag,
                "gsx_fusion_merge_append: null frame pointer");
  GSX_CHECK_ARG(epoch >= 1 && epoch < (1u << 30), "gsx_fusion_merge_append: epoch out of range");
  const Workspace ws = carve(workspace, B, H, W);
  MergeArgs a{map_points, map_normals, map_colors, map_ccounts, counts_in, counts_out, capacity, depth, depth_bstride,
              rgb, rgb_bstride, intrinsics, K_bstride, gvertex, gnormal, B, H, W,
              (float)(2.0 * (sigma * sigma)), ws, epoch, overflow_flag};
  return launch_merge_append(a, (cudaStream_t)stream);
}
