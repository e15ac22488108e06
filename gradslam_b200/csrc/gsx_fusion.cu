// Fused PointFusion map update for sm_100a.
//   k_frame_records   (K1r)    one thread per live pixel: world-frame vertex / normal, confidence weight and depth of the
//                              pixel as ONE 32-byte record (the whole op chain of rgbdimages.py:643-762 and
//                              fusionutils.py:16-73, evaluated once per pixel); also re-arms the per-frame workspace.
//   k_project_select  (K2+K3)  one thread per map row: project into the live camera, frustum test, ONE 32-byte gather of
//                              the frame record under the projection, distance / normal tests, then a 128-bit atomic
//                              arg-min per pixel on the key (1/ccount, ray distance, row index).
//   k_merge_append    (K4)     one thread per pixel: confidence-weighted merge of the selected map row, or stable append of
//                              unmatched valid pixels (single-pass decoupled look-back scan, row-major order per batch
//                              element).  No float atomics anywhere.
// Map rows are sector-packed: geometry rows (px,py,pz,nx,ny,nz,ccount,0) of exactly one 32-byte
// sector, colour rows (r,g,b,0) of 16 bytes; every row access is a 128-bit load / store.
// Reference op chains: gradslam/slam/fusionutils.py:198-722 (see include/gsx.h).
#include <cuda.h>  // CUtensorMap (the encoder is fetched with cudaGetDriverEntryPoint: no link against libcuda)

#include "gsx_common.cuh"
#include "gsx_exp.cuh"
#include "gsx_thresholds.h"
#include "../../include/gsx.h"

namespace gsx {

constexpr int kBlock = 256;
#ifndef GSX_KPIX
#define GSX_KPIX 2
#endif
constexpr int kMB = 256;                  // threads per CTA of the merge/append kernel
constexpr int kPix = GSX_KPIX;            // pixels per thread
constexpr int kTilePix = kMB * kPix;      // pixels per merge tile
constexpr int kGeoW = 8, kColW = 4, kRecW = 8;  // floats per geometry row / colour row / frame record

// ---- workspace layout -----------------------------------------------------------------------------------
//   float  frec[B][P][8]       frame records (gvx,gvy,gvz,gnx,gny,gnz,alpha,depth)             written by K1r
//   U128   best[B][P]          complemented arg-min records (0 = empty)                         cleared by K1r
//   uint64 tile_state[B][T]    (flag<<32 | value) of the look-back scan, T = ceil(P / kTilePix) cleared by K1r
//   uint32 ticket[B]           dynamic tile ids of K4                                           cleared by K1r
//   uint64 stats[B][2]         running totals: {active map rows (in frustum), merged rows}      caller zeroes once
// Nothing in here has to survive from one frame to the next (the stats are bookkeeping only): every frame's K1r
// re-arms what K2 / K4 of that frame consume, so a failed or abandoned call cannot poison a later one.
struct Workspace {
  float *frec;
  U128 *best;
  unsigned long long *tile_state;
  unsigned int *ticket;
  unsigned long long *stats;
  int tiles;
};

__host__ __device__ inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

inline Workspace carve(void *ws, int B, int H, int W) {
  const int64_t P = (int64_t)H * W;
  Workspace w;
  w.tiles = (int)((P + kTilePix - 1) / kTilePix);
  char *p = (char *)ws;
  w.frec = (float *)p;
  p += align_up(B * P * 32, 256);
  w.best = (U128 *)p;
  p += align_up(B * P * 16, 256);
  w.tile_state = (unsigned long long *)p;
  p += align_up((int64_t)B * w.tiles * 8, 256);
  w.ticket = (unsigned int *)p;
  p += align_up((int64_t)B * 4, 256);
  w.stats = (unsigned long long *)p;
  return w;
}

inline int64_t stats_offset(int B, int H, int W) {
  const int64_t P = (int64_t)H * W;
  const int64_t tiles = (P + kTilePix - 1) / kTilePix;
  return align_up(B * P * 32, 256) + align_up(B * P * 16, 256) + align_up(B * tiles * 8, 256) +
         align_up((int64_t)B * 4, 256);
}

inline int64_t workspace_bytes(int B, int H, int W) { return stats_offset(B, H, W) + align_up((int64_t)B * 16, 256); }

// alpha = clamp(exp(-|v|^2 / 2 sigma^2), 1e-7, 1.01) (fusionutils.py:69-72).  The exponential is evaluated in double and
// rounded once: that is the correctly rounded float32 exp (up to 2^-29 odds), so the CUDA path and the CPU oracle agree
// bit for bit and no later threshold / arg-min decision can flip because of a 1-ulp difference in a confidence weight.
__device__ __forceinline__ float confidence_exp(float sq_norm, float two_sigma_sq) {
  const float x = (-sq_norm) / two_sigma_sq;
  if (!(x >= -17.0f)) return 0.0f;  // exp(x) < 4.2e-8: clamps to 1e-7 below (also NaN, like fmaxf(NaN, 1e-7f))
  return exp_f32_via_f64(x);
}
__device__ __forceinline__ float confidence_alpha(float sq_norm, float two_sigma_sq) {
  return fminf(fmaxf(confidence_exp(sq_norm, two_sigma_sq), 1e-7f), 1.01f);
}

// ---- K1r ------------------------------------------------------------------------------------------------
struct FrameRecArgs {
  const float *depth;  // (B,H,W) live depth
  int64_t depth_bstride;
  const float *K;
  int64_t K_bstride;
  const float *poses;  // camera-to-world, or null: world frame == camera frame
  int64_t pose_bstride;
  const float *gv, *gn, *vloc;  // kFromMaps: materialised (B,H,W,3) world vertex / world normal / camera vertex maps
  int B, H, W;
  float two_sigma_sq;
  Workspace ws;
};

constexpr int kRecTW = 32, kRecTH = 8;  // pixel tile of one K1r CTA

template <bool kFromMaps>
__global__ void __launch_bounds__(kRecTW *kRecTH) k_frame_records(FrameRecArgs a) {
  __shared__ Rigid s_pose;
  __shared__ KInv s_k;
  const int b = blockIdx.z;
  const int tid = threadIdx.y * kRecTW + threadIdx.x;
  if (!kFromMaps) {
    if (tid == 0) s_k = load_kinv(a.K + b * a.K_bstride);
    if (tid == 32 && a.poses) s_pose = load_rigid(a.poses + b * a.pose_bstride);
  }
  // re-arm the scan state of this element for the frame's K4
  const int lin = (blockIdx.y * gridDim.x + blockIdx.x) * (kRecTW * kRecTH) + tid;
  if (lin < a.ws.tiles) a.ws.tile_state[(int64_t)b * a.ws.tiles + lin] = 0ull;
  if (lin == 0) a.ws.ticket[b] = 0u;
  if (!kFromMaps) __syncthreads();
  const int w = blockIdx.x * kRecTW + threadIdx.x, h = blockIdx.y * kRecTH + threadIdx.y;
  if (w >= a.W || h >= a.H) return;
  const int P = a.H * a.W;
  const int pix = h * a.W + w;
  const float *dimg = a.depth + b * a.depth_bstride;
  float4 r0, r1;
  if (kFromMaps) {
    const int64_t o = ((int64_t)b * P + pix) * 3;
    const float vx = __ldg(a.vloc + o), vy = __ldg(a.vloc + o + 1), vz = __ldg(a.vloc + o + 2);
    r0 = make_float4(__ldg(a.gv + o), __ldg(a.gv + o + 1), __ldg(a.gv + o + 2), __ldg(a.gn + o));
    r1 = make_float4(__ldg(a.gn + o + 1), __ldg(a.gn + o + 2),
                     confidence_alpha((vx * vx + vy * vy) + vz * vz, a.two_sigma_sq), __ldg(dimg + pix));
  } else {
    const FrameSample f = frame_sample<true>(dimg, s_k, a.poses ? &s_pose : nullptr, h, w, a.H, a.W);
    // alpha from the LOCAL vertex (fusionutils.py:657, 69-72)
    const float s = (f.v.x * f.v.x + f.v.y * f.v.y) + f.v.z * f.v.z;
    r0 = make_float4(f.gv.x, f.gv.y, f.gv.z, f.gn.x);
    r1 = make_float4(f.gn.y, f.gn.z, confidence_alpha(s, a.two_sigma_sq), f.d);
  }
  float4 *rec = reinterpret_cast<float4 *>(a.ws.frec + ((int64_t)b * P + pix) * kRecW);
  rec[0] = r0;
  rec[1] = r1;
  a.ws.best[(int64_t)b * P + pix] = U128{0ull, 0ull};
}

// The same kernel with the depth tile staged by the TMA unit: the frame is a regular grid, so the (32 + halo) x (8 + halo)
// depth tile a CTA needs (each pixel reads its right and lower neighbour) is ONE 3-D tensor-map box {36, 9, 1} of the
// (W, H, element) depth tensor, copied into shared memory by cp.async.bulk.tensor (SASS: UTMALDG) and signalled on an
// mbarrier, while the CTA fetches its constants.  Elements of the box beyond the image are zero-filled by the unit; they
// are never used (edge pixels take the difference of the previous column / row, which lies inside the box).  The arithmetic
// is frame_sample_from(): bit-identical to frame_sample<true>() of the plain kernel.
constexpr int kRecBoxW = kRecTW + 4, kRecBoxH = kRecTH + 1;  // box width: 36 floats = 144 bytes (a multiple of 16)

__device__ __forceinline__ unsigned int smem_u32(const void *p) { return (unsigned int)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(kRecTW *kRecTH) k_frame_records_tma(const __grid_constant__ CUtensorMap tmap,
                                                                       FrameRecArgs a) {
  __shared__ __align__(128) float s_d[kRecBoxH][kRecBoxW];
  __shared__ __align__(8) unsigned long long s_bar;
  __shared__ Rigid s_pose;
  __shared__ KInv s_k;
  const int b = blockIdx.z;
  const int tid = threadIdx.y * kRecTW + threadIdx.x;
  const int w0 = blockIdx.x * kRecTW, h0 = blockIdx.y * kRecTH;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&s_bar)),
                 "r"((unsigned int)(kRecBoxH * kRecBoxW * 4))
                 : "memory");
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(&s_d[0][0])),
        "l"(&tmap), "r"(smem_u32(&s_bar)), "r"(w0), "r"(h0), "r"(b)
        : "memory");
  }
  // constants and the scan state of the frame's K4 while the tile is in flight
  if (tid == 32) s_k = load_kinv(a.K + b * a.K_bstride);
  if (tid == 64 && a.poses) s_pose = load_rigid(a.poses + b * a.pose_bstride);
  const int lin = (blockIdx.y * gridDim.x + blockIdx.x) * (kRecTW * kRecTH) + tid;
  if (lin < a.ws.tiles) a.ws.tile_state[(int64_t)b * a.ws.tiles + lin] = 0ull;
  if (lin == 0) a.ws.ticket[b] = 0u;
  __syncthreads();
  {  // wait for the tile (phase 0 of the barrier)
    unsigned int done = 0;
    while (!done)
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done)
          : "r"(smem_u32(&s_bar))
          : "memory");
  }
  const int w = w0 + threadIdx.x, h = h0 + threadIdx.y;
  if (w >= a.W || h >= a.H) return;
  const int P = a.H * a.W;
  const int pix = h * a.W + w;
  const int wa = (w < a.W - 1) ? w : w - 1, ha = (h < a.H - 1) ? h : h - 1;
  DepthStencil t;
  t.c = s_d[threadIdx.y][threadIdx.x];
  t.l = s_d[threadIdx.y][wa - w0];
  t.r = s_d[threadIdx.y][wa - w0 + 1];
  t.u = s_d[ha - h0][threadIdx.x];
  t.d = s_d[ha - h0 + 1][threadIdx.x];
  const FrameSample f = frame_sample_from(t, s_k, a.poses ? &s_pose : nullptr, h, w, a.H, a.W);
  // alpha from the LOCAL vertex (fusionutils.py:657, 69-72)
  const float s = (f.v.x * f.v.x + f.v.y * f.v.y) + f.v.z * f.v.z;
  float4 *rec = reinterpret_cast<float4 *>(a.ws.frec + ((int64_t)b * P + pix) * kRecW);
  rec[0] = make_float4(f.gv.x, f.gv.y, f.gv.z, f.gn.x);
  rec[1] = make_float4(f.gn.y, f.gn.z, confidence_alpha(s, a.two_sigma_sq), f.d);
  a.ws.best[(int64_t)b * P + pix] = U128{0ull, 0ull};
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn tensor_map_encoder() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
    cudaGetLastError();  // (a failed query must not poison the next launch check)
  }
  return fn;
}

// depth (nb, H, W) with element stride depth_bstride as a 3-D tensor map; false if the layout does not qualify
static bool depth_tensor_map(const FrameRecArgs &a, CUtensorMap *tm) {
  if (getenv("GSX_NO_TMA")) return false;
  const EncodeTiledFn enc = tensor_map_encoder();
  if (!enc) return false;
  // strides must be multiples of 16 bytes, the base 16-byte aligned; a last column / row that starts a tile of its own
  // would need a halo on the other side
  if (a.W % 4 || a.depth_bstride % 4 || (reinterpret_cast<uintptr_t>(a.depth) & 15)) return false;
  if ((a.W - 1) % kRecTW == 0 || (a.H - 1) % kRecTH == 0 || a.W < 2 || a.H < 2) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.B};
  const cuuint64_t strides[2] = {(cuuint64_t)a.W * 4, (cuuint64_t)a.depth_bstride * 4};
  const cuuint32_t box[3] = {(cuuint32_t)kRecBoxW, (cuuint32_t)kRecBoxH, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(a.depth), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int launch_frame_records(const FrameRecArgs &a, cudaStream_t stream) {
  if (a.B == 0) return 0;
  const dim3 grid((unsigned)((a.W + kRecTW - 1) / kRecTW), (unsigned)((a.H + kRecTH - 1) / kRecTH), (unsigned)a.B);
  const dim3 block(kRecTW, kRecTH);
  CUtensorMap tm;
  if (!a.gv && depth_tensor_map(a, &tm)) {
    k_frame_records_tma<<<grid, block, 0, stream>>>(tm, a);
    GSX_CHECK_LAUNCH("gsx_fusion_frame_records(tma)");
    return 0;
  }
  if (a.gv)
    k_frame_records<true><<<grid, block, 0, stream>>>(a);
  else
    k_frame_records<false><<<grid, block, 0, stream>>>(a);
  GSX_CHECK_LAUNCH("gsx_fusion_frame_records");
  return 0;
}

// ---- K2 + K3 ------------------------------------------------------------------------------------------
struct ProjectArgs {
  const float *geo;  // (B,cap,8) geometry rows
  const int32_t *counts;
  int64_t cap;
  const float *poses;
  int64_t pose_bstride;
  const float *K;
  int64_t K_bstride;
  int B, H, W;
  float dot_th, u_hi, v_hi;  // u_hi = float(W - 0.999), v_hi = float(H - 0.999)
  float d2_max;              // largest float x with sqrtf(x) < dist_th (-1 if none): sqrtf(d2) < dist_th <=> d2 <= d2_max
  const float *frec;
  U128 *best;
  unsigned long long *stats;
};

#ifndef GSX_K2_MINB
#define GSX_K2_MINB 5
#endif

struct MapRow {  // one geometry row
  float4 a, b;   // a = (px,py,pz,nx)  b = (ny,nz,cc,0)
};
__device__ __forceinline__ MapRow load_map_row(const float *geo, int64_t n) {
  MapRow m;
  m.a = __ldg(reinterpret_cast<const float4 *>(geo + n * kGeoW));
  m.b = __ldg(reinterpret_cast<const float4 *>(geo + n * kGeoW + 4));
  return m;
}

// The kernel is bound by (threads in flight) / (length of the dependent memory chain) and by instruction issue, not by
// bytes, so the chain is kept short and the per-row work small:
//   * the map row of the NEXT grid-stride iteration is fetched while the current row is processed (software pipelining);
//   * everything the tests need from the frame sits in ONE 32-byte record under the projection (K1r): one gather;
//   * sqrtf(d2) < dist_th is decided as d2 <= d2_max (exact: the correctly rounded square root is monotonic; the
//     threshold is found on the host, gsx_thresholds.h);
//   * the result of the 128-bit CAS is only looked at one iteration later.
__global__ void __launch_bounds__(kBlock, GSX_K2_MINB) k_project_select(ProjectArgs a) {
  __shared__ Rigid s_tinv;
  __shared__ float s_k[12];
  __shared__ unsigned int s_act[kBlock / 32];
  const int b = blockIdx.y;
  const int count = a.counts[b];
  if ((int64_t)blockIdx.x * kBlock >= count) return;
  if (threadIdx.x == 0) s_tinv = rigid_inverse(load_rigid(a.poses + b * a.pose_bstride));
  if (threadIdx.x >= 32 && threadIdx.x < 44) s_k[threadIdx.x - 32] = __ldg(a.K + b * a.K_bstride + (threadIdx.x - 32));
  __syncthreads();
  const int P = a.H * a.W;
  const float *geo = a.geo + (int64_t)b * a.cap * kGeoW;
  const float *frec = a.frec + (int64_t)b * P * kRecW;
  U128 *best = a.best + (int64_t)b * P;
  unsigned int n_active = 0;
  const int64_t stride = (int64_t)gridDim.x * kBlock;
  int64_t n = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  U128 mine{0ull, 0ull}, old{0ull, 0ull};
  int pend_pix = -1;
  MapRow cur = load_map_row(geo, n < count ? n : 0);
  for (; n < count; n += stride) {
    const MapRow m = cur;
    const int64_t nn = n + stride;
    if (nn < count) cur = load_map_row(geo, nn);  // in flight while this row is processed
    // world -> camera (pointclouds.py:526-573), then pinhole projection with the 4x4 K on the homogeneous
    // point (projutils.py:92-238): z == 0 divides by 1.
    const float3 q = rigid_apply(s_tinv, m.a.x, m.a.y, m.a.z);
    const float hx = ((s_k[0] * q.x + s_k[1] * q.y) + s_k[2] * q.z) + s_k[3];
    const float hy = ((s_k[4] * q.x + s_k[5] * q.y) + s_k[6] * q.z) + s_k[7];
    const float hz = ((s_k[8] * q.x + s_k[9] * q.y) + s_k[10] * q.z) + s_k[11];
    const float den = (hz != 0.0f) ? hz : 1.0f;
    const float u = hx / den, v = hy / den;
    // fusionutils.py:259-266
    bool live = (u > -1e-3f) && (u < a.u_hi) && (v > -1e-3f) && (v < a.v_hi) && (q.z > 0.0f);
    if (live) {
      ++n_active;
      // round-half-even like torch.round, then clamp (fusionutils.py:267-274)
      int w = (int)rintf(u), h = (int)rintf(v);
      w = min(max(w, 0), a.W - 1);
      h = min(max(h, 0), a.H - 1);
      const int pix = h * a.W + w;
      const float4 f0 = __ldg(reinterpret_cast<const float4 *>(frec + (int64_t)pix * kRecW));
      const float2 f1 = __ldg(reinterpret_cast<const float2 *>(frec + (int64_t)pix * kRecW + 4));
      // are_points_close (fusionutils.py:130): ||frame - map|| < dist_th
      const float dx = f0.x - m.a.x, dy = f0.y - m.a.y, dz = f0.z - m.a.z;
      const float d2 = (dx * dx + dy * dy) + dz * dz;
      // are_normals_similar (fusionutils.py:187-195): n_frame . n_map > dot_th
      const float dot = (f0.w * m.a.w + f1.x * m.b.x) + f1.y * m.b.y;
      live = (d2 <= a.d2_max) && (dot > a.dot_th);
      if (pend_pix >= 0) {  // settle the previous candidate's CAS before re-using the slot
        atomic_max_rec128_finish(best + pend_pix, mine, old);
        pend_pix = -1;
      }
      if (live) {
        // sort key of find_best_unique_correspondences (fusionutils.py:491-517): 1/(cc+1e-20), then the squared
        // distance (map - frame)^2 (== d2: squares are sign-independent), then n.
        const float inv_cc = 1.0f / (m.b.z + 1e-20f);
        // positive floats order like their bit patterns; flip negatives so the order stays total.
        unsigned int kb = __float_as_uint(inv_cc);
        kb = (kb & 0x80000000u) ? ~kb : (kb | 0x80000000u);
        const unsigned int rb = __float_as_uint(d2) | 0x80000000u;  // d2 >= 0
        const unsigned long long hi = ((unsigned long long)kb << 32) | rb;
        mine = U128{~(unsigned long long)n, ~hi};
        old = cas128(best + pix, U128{0ull, 0ull}, mine);  // optimistic: most pixels see a single candidate
        pend_pix = pix;
      }
    }
  }
  if (pend_pix >= 0) atomic_max_rec128_finish(best + pend_pix, mine, old);
  // bookkeeping for the roofline's algorithmic-byte count: ONE atomic per CTA (every warp of the grid adding to the same
  // address serialises in the L2: 9472 same-address atomics cost ~20 us when the whole grid works on one map)
  n_active = __reduce_add_sync(0xffffffffu, n_active);
  if ((threadIdx.x & 31) == 0) s_act[threadIdx.x >> 5] = n_active;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int t = 0;
#pragma unroll
    for (int i = 0; i < kBlock / 32; ++i) t += s_act[i];
    if (t) atomicAdd(a.stats + 2 * b, (unsigned long long)t);
  }
}

#ifndef GSX_K2_CTAS_PER_SM
#define GSX_K2_CTAS_PER_SM 8
#endif

int launch_project_select(const ProjectArgs &a, int64_t max_count, cudaStream_t stream) {
  if (a.B == 0 || max_count <= 0) return 0;
  int64_t bx = (max_count + kBlock - 1) / kBlock;
  const int64_t cap_blocks = (int64_t)kNumSMs * GSX_K2_CTAS_PER_SM;  // grid-stride beyond this many CTAs per SM
  if (bx * a.B > cap_blocks) bx = (cap_blocks + a.B - 1) / a.B;
  if (bx < 1) bx = 1;
  k_project_select<<<dim3((unsigned)bx, (unsigned)a.B), kBlock, 0, stream>>>(a);
  GSX_CHECK_LAUNCH("gsx_fusion_project_select");
  return 0;
}

// ---- K4 -------------------------------------------------------------------------------------------------
struct MergeArgs {
  float *geo, *col;  // (B,cap,8), (B,cap,4)
  int with_cc;       // 0: maps without confidence counts (ICPSLAM aggregation): nothing merges, slot 6 stays 0
  const int32_t *counts_in;
  int32_t *counts_out;
  int64_t cap;
  const float *rgb;  // (B,H,W,3) live colours
  int64_t rgb_bstride;
  int B, H, W;
  Workspace ws;
  int32_t *overflow;
  int32_t *assoc;  // optional (B,P): +row+1 appended at `row`, -(row+1) merged into `row`, 0 untouched
};

constexpr unsigned long long kFlagAgg = 1ull, kFlagPrefix = 2ull;

__device__ __forceinline__ unsigned long long pack_state(unsigned long long flag, unsigned int value) {
  return (flag << 32) | value;
}
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// exclusive prefix of the new-point counts of all preceding tiles (decoupled look-back, one warp, 32
// predecessors per step)
__device__ __forceinline__ unsigned int lookback_warp(const unsigned long long *state, int tile, int lane) {
  unsigned int excl = 0;
  for (int base = tile - 1; base >= 0; base -= 32) {
    const int j = base - lane;
    unsigned long long s = 0ull;
    if (j >= 0) {
      do {
        s = ld_acquire_u64(state + j);
      } while ((s >> 32) == 0ull);
    }
    const bool is_prefix = (j >= 0) && ((s >> 32) == kFlagPrefix);
    const unsigned int pm = __ballot_sync(0xffffffffu, is_prefix);
    const int first = pm ? (__ffs(pm) - 1) : 32;  // nearest predecessor that already knows its inclusive prefix
    const unsigned int v = (j >= 0 && lane <= first) ? (unsigned int)s : 0u;
    excl += __reduce_add_sync(0xffffffffu, v);
    if (pm) break;
  }
  return excl;
}

#ifndef GSX_K4_MINB
#define GSX_K4_MINB 4
#endif

template <bool kAssoc>
__global__ void __launch_bounds__(kMB, GSX_K4_MINB) k_merge_append(MergeArgs a) {
  __shared__ int s_tile;
  __shared__ int s_warp_sums[kPix][kMB / 32];
  __shared__ int s_matched[kMB / 32];
  __shared__ int s_excl;
  __shared__ __align__(16) float s_rgb[kTilePix * 3];
  // batch element varies fastest in the grid: CTAs resident at the same time belong to different elements, so
  // each element's look-back chain only sees ~1/B of the in-flight tiles
  const int b = blockIdx.x % a.B;
  const int T = a.ws.tiles;
  if (threadIdx.x == 0) s_tile = (int)atomicAdd(a.ws.ticket + b, 1u);  // tiles start in ticket order
  const int count_in = a.counts_in[b];  // loaded early: its latency hides behind everything below
  __syncthreads();
  const int tile = s_tile;
  const int P = a.H * a.W;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int pix0 = tile * kTilePix;
  const U128 *best = a.ws.best + (int64_t)b * P;
  const float *frec = a.ws.frec + (int64_t)b * P * kRecW;
  const float *rgb = a.rgb + b * a.rgb_bstride + (int64_t)pix0 * 3;

  // live colours of the tile: coalesced 128-bit loads into shared memory (stride-3 reads are conflict free)
  const int tile_px = min(kTilePix, P - pix0);
  if ((reinterpret_cast<uintptr_t>(rgb) & 15) == 0) {
    const int n4 = (tile_px * 3) >> 2;
    for (int i = threadIdx.x; i < n4; i += kMB)
      reinterpret_cast<float4 *>(s_rgb)[i] = __ldg(reinterpret_cast<const float4 *>(rgb) + i);
    for (int i = (n4 << 2) + threadIdx.x; i < tile_px * 3; i += kMB) s_rgb[i] = __ldg(rgb + i);
  } else {
    for (int i = threadIdx.x; i < tile_px * 3; i += kMB) s_rgb[i] = __ldg(rgb + i);
  }

  int pix[kPix];
  unsigned long long rec_lo[kPix];
  float4 f0[kPix], f1[kPix];
  bool matched[kPix], is_new[kPix];
  int warp_excl[kPix];
#pragma unroll
  for (int j = 0; j < kPix; ++j) {
    pix[j] = pix0 + j * kMB + threadIdx.x;
    U128 rec{0ull, 0ull};
    f0[j] = f1[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (pix[j] < P) {
      rec = best[pix[j]];
      f0[j] = __ldg(reinterpret_cast<const float4 *>(frec + (int64_t)pix[j] * kRecW));
      f1[j] = __ldg(reinterpret_cast<const float4 *>(frec + (int64_t)pix[j] * kRecW + 4));
    }
    matched[j] = a.with_cc && ((rec.lo | rec.hi) != 0ull);
    rec_lo[j] = rec.lo;
  }
  int n_matched = 0;
#pragma unroll
  for (int j = 0; j < kPix; ++j) {
    is_new[j] = (pix[j] < P) && (f1[j].w > 0.0f) && !matched[j];
    n_matched += matched[j] ? 1 : 0;
    // row-major order inside the tile: chunk j (256 consecutive pixels), then warp, then lane
    const unsigned int ballot = __ballot_sync(0xffffffffu, is_new[j]);
    warp_excl[j] = __popc(ballot & ((1u << lane) - 1u));
    if (lane == 0) s_warp_sums[j][warp] = __popc(ballot);
  }
  n_matched = __reduce_add_sync(0xffffffffu, n_matched);
  if (lane == 0) s_matched[warp] = n_matched;
  __syncthreads();
  if (threadIdx.x == 0) {  // bookkeeping (merged rows of this element): one atomic per CTA
    int t = 0;
#pragma unroll
    for (int i = 0; i < kMB / 32; ++i) t += s_matched[i];
    if (t) atomicAdd(a.ws.stats + 2 * b + 1, (unsigned long long)t);
  }
  int block_total = 0;
  int block_excl[kPix];
#pragma unroll
  for (int j = 0; j < kPix; ++j) {
    block_excl[j] = block_total;
#pragma unroll
    for (int i = 0; i < kMB / 32; ++i) {
      const int c = s_warp_sums[j][i];
      if (i < warp) block_excl[j] += c;
      block_total += c;
    }
  }
  unsigned long long *state = a.ws.tile_state + (int64_t)b * T;
  if (threadIdx.x == 0 && tile + 1 < T) st_release_u64(state + tile, pack_state(kFlagAgg, (unsigned)block_total));

  float *geo = a.geo + (int64_t)b * a.cap * kGeoW;
  float *col = a.col + (int64_t)b * a.cap * kColW;

  // matched map rows: issue all loads first, then the arithmetic and the stores
  float4 g0[kPix], g1[kPix], c4[kPix];
#pragma unroll
  for (int j = 0; j < kPix; ++j) {
    if (matched[j]) {
      const int64_t n = (int64_t)(~rec_lo[j]);
      g0[j] = *reinterpret_cast<const float4 *>(geo + n * kGeoW);
      g1[j] = *reinterpret_cast<const float4 *>(geo + n * kGeoW + 4);
      c4[j] = *reinterpret_cast<const float4 *>(col + n * kColW);
    }
  }
#pragma unroll
  for (int j = 0; j < kPix; ++j) {
    if (matched[j]) {
      // confidence-weighted running mean (fusionutils.py:678-699); exactly one pixel owns this map row
      const int64_t n = (int64_t)(~rec_lo[j]);
      const float alpha = f1[j].z;
      const float c0 = g1[j].z;
      const float tot = c0 + alpha;
      const float inv = 1.0f / ((tot == 0.0f) ? 1.0f : tot);
      const float *fc = s_rgb + (j * kMB + (int)threadIdx.x) * 3;
      float4 o0, o1, oc;
      o0.x = ((c0 * g0[j].x) + (alpha * f0[j].x)) * inv;
      o0.y = ((c0 * g0[j].y) + (alpha * f0[j].y)) * inv;
      o0.z = ((c0 * g0[j].z) + (alpha * f0[j].z)) * inv;
      o0.w = ((c0 * g0[j].w) + (alpha * f0[j].w)) * inv;
      o1.x = ((c0 * g1[j].x) + (alpha * f1[j].x)) * inv;
      o1.y = ((c0 * g1[j].y) + (alpha * f1[j].y)) * inv;
      o1.z = tot;
      o1.w = 0.0f;
      oc.x = ((c0 * c4[j].x) + (alpha * fc[0])) * inv;
      oc.y = ((c0 * c4[j].y) + (alpha * fc[1])) * inv;
      oc.z = ((c0 * c4[j].z) + (alpha * fc[2])) * inv;
      oc.w = 0.0f;
      *reinterpret_cast<float4 *>(geo + n * kGeoW) = o0;
      *reinterpret_cast<float4 *>(geo + n * kGeoW + 4) = o1;
      *reinterpret_cast<float4 *>(col + n * kColW) = oc;
      if (kAssoc) a.assoc[(int64_t)b * P + pix[j]] = -(int32_t)(n + 1);
    }
  }

  // decoupled look-back (warp 0): exclusive prefix of new-point counts over preceding tiles of this element
  if (warp == 0) {
    const unsigned int excl = lookback_warp(state, tile, lane);
    if (lane == 0) {
      if (tile + 1 < T) st_release_u64(state + tile, pack_state(kFlagPrefix, excl + (unsigned)block_total));
      s_excl = (int)excl;
    }
  }
  __syncthreads();
  const int64_t base = (int64_t)count_in + s_excl;
#pragma unroll
  for (int j = 0; j < kPix; ++j) {
    if (is_new[j]) {
      // append in row-major pixel order (fusionutils.py:702-720; pointclouds.py:1203-1235)
      const int64_t n = base + block_excl[j] + warp_excl[j];
      if (n < a.cap) {
        const float *fc = s_rgb + (j * kMB + (int)threadIdx.x) * 3;
        *reinterpret_cast<float4 *>(geo + n * kGeoW) = f0[j];
        *reinterpret_cast<float4 *>(geo + n * kGeoW + 4) =
            make_float4(f1[j].x, f1[j].y, a.with_cc ? f1[j].z : 0.0f, 0.0f);
        *reinterpret_cast<float4 *>(col + n * kColW) = make_float4(fc[0], fc[1], fc[2], 0.0f);
        if (kAssoc) a.assoc[(int64_t)b * P + pix[j]] = (int32_t)(n + 1);
      } else {
        *a.overflow = 1;
      }
    }
  }
  if (tile == T - 1 && threadIdx.x == 0) {
    const int64_t total = base + block_total;
    a.counts_out[b] = (int32_t)(total < a.cap ? total : a.cap);
  }
}

int launch_merge_append(const MergeArgs &a, cudaStream_t stream) {
  if (a.B == 0) return 0;
  const dim3 grid((unsigned)(a.ws.tiles * a.B));
  if (a.assoc)
    k_merge_append<true><<<grid, kMB, 0, stream>>>(a);  // differentiable forward
  else
    k_merge_append<false><<<grid, kMB, 0, stream>>>(a);
  GSX_CHECK_LAUNCH("gsx_fusion_merge_append");
  return 0;
}

// ---- K4 backward ------------------------------------------------------------------------------------------
// The differentiable forward (k_merge_append<true>) leaves, per pixel, where its sample went: merged into map row n
// (assoc = -(n+1)), appended as row n (assoc = n+1) or dropped (0).  With the pre-merge map and the frame values the
// backward is a pure per-pixel / per-row map - no atomics, no scan:
//   merged   out = (c*m + a*f) * inv,  inv = 1/(c+a)   (fusionutils.py:678-699)
//            d m = g*c*inv      d f = g*a*inv      d c += g*(m*inv - num*inv^2)      d a += g*(f*inv - num*inv^2)
//            cc_out = c + a  =>  d c += g_cc,  d a += g_cc
//   appended out = f, cc_out = a   =>  d f = g,  d a = g_cc
//   alpha    a = clamp(exp(-|v|^2 / 2 sigma^2), 1e-7, 1.01) (fusionutils.py:69-72)  =>  d v = d a * e * (-2 v / 2 sigma^2)
//            inside the clamp range, 0 outside.
// Map tensors and their gradients use the packed row layout (geometry rows of 8 floats, colour rows of 4; the padding
// slots carry zero gradient).
struct MergeBwdArgs {
  const int32_t *assoc;
  const int32_t *counts_in;
  const float *geo, *col;  // pre-merge map (B, cap_in, 8 / 4)
  int with_cc;
  int64_t cap_in;
  const float *g_geo, *g_col;  // upstream gradients (B, cap_out, 8 / 4); null = zero
  int64_t cap_out;
  const float *gv, *gn, *rgb, *vloc;  // frame values (B, P, 3)
  float *d_geo, *d_col;               // (B, cap_in, 8 / 4)
  float *d_gv, *d_gn, *d_rgb, *d_vloc;  // (B, P, 3)
  int B, P;
  float two_sigma_sq;
};

// rows the frame did not touch pass their gradient through; padding rows get zero
__global__ void __launch_bounds__(256) k_merge_bwd_rows(MergeBwdArgs a) {
  const int b = blockIdx.y;
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= a.cap_in) return;
  const bool live = n < a.counts_in[b] && n < a.cap_out;
  const int64_t ri = (int64_t)b * a.cap_in + n, ro = (int64_t)b * a.cap_out + n;
  float4 z = make_float4(0.f, 0.f, 0.f, 0.f), g0 = z, g1 = z, gc = z;
  if (live && a.g_geo) {
    g0 = *reinterpret_cast<const float4 *>(a.g_geo + ro * kGeoW);
    g1 = *reinterpret_cast<const float4 *>(a.g_geo + ro * kGeoW + 4);
    g1.w = 0.0f;
    if (!a.with_cc) g1.z = 0.0f;
  }
  if (live && a.g_col) {
    gc = *reinterpret_cast<const float4 *>(a.g_col + ro * kColW);
    gc.w = 0.0f;
  }
  *reinterpret_cast<float4 *>(a.d_geo + ri * kGeoW) = g0;
  *reinterpret_cast<float4 *>(a.d_geo + ri * kGeoW + 4) = g1;
  *reinterpret_cast<float4 *>(a.d_col + ri * kColW) = gc;
}

__global__ void __launch_bounds__(256) k_merge_bwd_pixels(MergeBwdArgs a) {
  const int b = blockIdx.y;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= a.P) return;
  const int64_t fi = ((int64_t)b * a.P + pix) * 3;
  const int32_t as = a.assoc[(int64_t)b * a.P + pix];
  float dgv[3] = {0.f, 0.f, 0.f}, dgn[3] = {0.f, 0.f, 0.f}, dc[3] = {0.f, 0.f, 0.f}, dv[3] = {0.f, 0.f, 0.f};
  if (as != 0) {
    const int64_t n = (as > 0) ? (int64_t)as - 1 : -(int64_t)as - 1;
    const int64_t ro = (int64_t)b * a.cap_out + n;
    float g[9];
#pragma unroll
    for (int q = 0; q < 6; ++q) g[q] = a.g_geo ? a.g_geo[ro * kGeoW + q] : 0.0f;
#pragma unroll
    for (int q = 0; q < 3; ++q) g[6 + q] = a.g_col ? a.g_col[ro * kColW + q] : 0.0f;
    const float gcc = (a.g_geo && a.with_cc) ? a.g_geo[ro * kGeoW + 6] : 0.0f;
    const float vx = a.vloc[fi], vy = a.vloc[fi + 1], vz = a.vloc[fi + 2];
    const float sq = (vx * vx + vy * vy) + vz * vz;
    const float e = (float)exp((double)((-sq) / a.two_sigma_sq));  // (cold path: library exp)
    float d_alpha = gcc;
    if (as > 0) {
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        dgv[q] = g[q];
        dgn[q] = g[3 + q];
        dc[q] = g[6 + q];
      }
    } else {
      const int64_t ri = (int64_t)b * a.cap_in + n;
      const float alpha = fminf(fmaxf(e, 1e-7f), 1.01f);
      const float c0 = a.geo[ri * kGeoW + 6];
      const float tot = c0 + alpha;
      const bool degenerate = tot == 0.0f;
      const float inv = 1.0f / (degenerate ? 1.0f : tot);
      const float dinv = degenerate ? 0.0f : -(inv * inv);
      float f[9], m[9];
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        f[q] = a.gv[fi + q];
        f[3 + q] = a.gn[fi + q];
        f[6 + q] = a.rgb[fi + q];
        m[q] = a.geo[ri * kGeoW + q];
        m[3 + q] = a.geo[ri * kGeoW + 3 + q];
        m[6 + q] = a.col[ri * kColW + q];
      }
      float d_c0 = gcc;
      float dm[9], df[9];
#pragma unroll
      for (int q = 0; q < 9; ++q) {
        const float num = c0 * m[q] + alpha * f[q];
        dm[q] = g[q] * c0 * inv;
        df[q] = g[q] * alpha * inv;
        d_c0 += g[q] * (m[q] * inv + num * dinv);
        d_alpha += g[q] * (f[q] * inv + num * dinv);
      }
#pragma unroll
      for (int q = 0; q < 6; ++q) a.d_geo[ri * kGeoW + q] = dm[q];
      a.d_geo[ri * kGeoW + 6] = d_c0;
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        a.d_col[ri * kColW + q] = dm[6 + q];
        dgv[q] = df[q];
        dgn[q] = df[3 + q];
        dc[q] = df[6 + q];
      }
    }
    if (e >= 1e-7f && e <= 1.01f) {
      const float s = d_alpha * e * (-2.0f / a.two_sigma_sq);
      dv[0] = s * vx;
      dv[1] = s * vy;
      dv[2] = s * vz;
    }
  }
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    a.d_gv[fi + q] = dgv[q];
    a.d_gn[fi + q] = dgn[q];
    a.d_rgb[fi + q] = dc[q];
    a.d_vloc[fi + q] = dv[q];
  }
}

// The two halves of one frame for the batch elements [b0, b0 + nb) of a B_total-element problem, on `st`: the frame
// records (K1r), and the map update that consumes them (K2 + K4).  All pointers are the FULL-batch base pointers; batch
// elements are independent, so disjoint groups may run concurrently on different streams, and the records of the next
// frame may be computed (into another workspace) while this frame's update runs (gsx_pointfusion_sequence_gt).
static Workspace group_workspace(void *workspace, int B_total, int b0, int H, int W) {
  const int64_t P = (int64_t)H * W;
  Workspace ws = carve(workspace, B_total, H, W);
  ws.frec += (int64_t)b0 * P * kRecW;
  ws.best += (int64_t)b0 * P;
  ws.tile_state += (int64_t)b0 * ws.tiles;
  ws.ticket += b0;
  ws.stats += 2 * b0;
  return ws;
}

int fusion_records_group(const float *poses, int64_t pose_bs, const float *K, int64_t K_bs, const float *depth,
                         int64_t d_bs, int B_total, int b0, int nb, int H, int W, double sigma, void *workspace,
                         cudaStream_t st) {
  FrameRecArgs fa{depth + (int64_t)b0 * d_bs, d_bs, K + (int64_t)b0 * K_bs, K_bs, poses + (int64_t)b0 * pose_bs, pose_bs,
                  nullptr, nullptr, nullptr, nb, H, W, (float)(2.0 * (sigma * sigma)),
                  group_workspace(workspace, B_total, b0, H, W)};
  return launch_frame_records(fa, st);
}

int fusion_update_group(float *geo, float *col, const int32_t *cin, int32_t *cout, int64_t cap, int64_t max_count,
                        const float *poses, int64_t pose_bs, const float *K, int64_t K_bs, const float *rgb,
                        int64_t rgb_bs, int B_total, int b0, int nb, int H, int W, float dist_th, float dot_th,
                        void *workspace, int32_t *overflow, cudaStream_t st) {
  const Workspace ws = group_workspace(workspace, B_total, b0, H, W);
  float *ggeo = geo + (int64_t)b0 * cap * kGeoW, *gcol = col + (int64_t)b0 * cap * kColW;
  if (max_count > 0) {
    ProjectArgs pa{ggeo, cin + b0, cap, poses + (int64_t)b0 * pose_bs, pose_bs, K + (int64_t)b0 * K_bs, K_bs, nb, H, W,
                   dot_th, (float)(W - 0.999), (float)(H - 0.999), sqrt_lt_threshold(dist_th), ws.frec, ws.best, ws.stats};
    const int rc = launch_project_select(pa, max_count, st);
    if (rc) return rc;
  }
  MergeArgs ma{ggeo, gcol, 1, cin + b0, cout + b0, cap, rgb + (int64_t)b0 * rgb_bs, rgb_bs, nb, H, W, ws, overflow, nullptr};
  return launch_merge_append(ma, st);
}

int64_t fusion_workspace_bytes(int B, int H, int W) { return workspace_bytes(B, H, W); }

}  // namespace gsx

using namespace gsx;

extern "C" int64_t gsx_fusion_workspace_bytes(int B, int H, int W) {
  if (B < 0 || H < 0 || W < 0) return -1;
  return workspace_bytes(B, H, W);
}

extern "C" int64_t gsx_fusion_workspace_stats_offset(int B, int H, int W) {
  if (B < 0 || H < 0 || W < 0) return -1;
  return stats_offset(B, H, W);
}

static bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" int gsx_fusion_frame_records(const float *depth, int64_t depth_bstride, const float *intrinsics,
                                        int64_t K_bstride, const float *poses, int64_t pose_bstride,
                                        const float *gvertex, const float *gnormal, const float *vertex, int B, int H,
                                        int W, double sigma, void *workspace, void *stream) {
  GSX_CHECK_ARG(B >= 0 && H >= 2 && W >= 2, "gsx_fusion_frame_records: bad extents B=%d H=%d W=%d", B, H, W);
  if (B == 0) return 0;
  GSX_CHECK_ARG(depth && workspace, "gsx_fusion_frame_records: null pointer");
  GSX_CHECK_ARG((gvertex && gnormal && vertex) || (!gvertex && !gnormal && !vertex && intrinsics),
                "gsx_fusion_frame_records: pass the three frame maps, or none of them together with the intrinsics");
  GSX_CHECK_ARG(aligned16(workspace), "gsx_fusion_frame_records: the workspace must be 16-byte aligned");
  FrameRecArgs a{depth, depth_bstride, intrinsics, K_bstride, poses, pose_bstride, gvertex, gnormal, vertex, B, H, W,
                 (float)(2.0 * (sigma * sigma)), carve(workspace, B, H, W)};
  return launch_frame_records(a, (cudaStream_t)stream);
}

extern "C" int gsx_fusion_project_select(const float *map_geometry, const int32_t *counts, int64_t capacity,
                                         int64_t max_count, const float *poses, int64_t pose_bstride,
                                         const float *intrinsics, int64_t K_bstride, int B, int H, int W,
                                         float dist_th, float dot_th, void *workspace, void *stream) {
  GSX_CHECK_ARG(B >= 0 && H >= 2 && W >= 2, "gsx_fusion_project_select: bad extents B=%d H=%d W=%d", B, H, W);
  if (max_count <= 0 || B == 0) return 0;
  GSX_CHECK_ARG(map_geometry && counts, "gsx_fusion_project_select: null map pointer");
  GSX_CHECK_ARG(poses && intrinsics && workspace, "gsx_fusion_project_select: null frame pointer");
  GSX_CHECK_ARG(aligned16(map_geometry) && aligned16(workspace),
                "gsx_fusion_project_select: geometry rows and workspace must be 16-byte aligned");
  GSX_CHECK_ARG(max_count <= capacity, "gsx_fusion_project_select: max_count %lld > capacity %lld",
                (long long)max_count, (long long)capacity);
  const Workspace ws = carve(workspace, B, H, W);
  ProjectArgs a{map_geometry, counts, capacity, poses, pose_bstride, intrinsics, K_bstride, B, H, W, dot_th,
                (float)(W - 0.999), (float)(H - 0.999), sqrt_lt_threshold(dist_th), ws.frec, ws.best, ws.stats};
  return launch_project_select(a, max_count, (cudaStream_t)stream);
}

extern "C" int gsx_fusion_merge_append(float *map_geometry, float *map_colors, int with_ccounts,
                                       const int32_t *counts_in, int32_t *counts_out, int64_t capacity,
                                       const float *rgb, int64_t rgb_bstride, int B, int H, int W, void *workspace,
                                       int32_t *overflow_flag, int32_t *assoc_out, void *stream) {
  GSX_CHECK_ARG(B >= 0 && H >= 2 && W >= 2, "gsx_fusion_merge_append: bad extents B=%d H=%d W=%d", B, H, W);
  if (B == 0) return 0;
  GSX_CHECK_ARG(map_geometry && map_colors && counts_in && counts_out, "gsx_fusion_merge_append: null map pointer");
  GSX_CHECK_ARG(counts_in != counts_out, "gsx_fusion_merge_append: counts_in and counts_out must not alias");
  GSX_CHECK_ARG(rgb && workspace && overflow_flag, "gsx_fusion_merge_append: null frame pointer");
  GSX_CHECK_ARG(aligned16(map_geometry) && aligned16(map_colors) && aligned16(workspace),
                "gsx_fusion_merge_append: map rows and workspace must be 16-byte aligned");
  GSX_CHECK_ARG(capacity <= 0x7fffffffll, "gsx_fusion_merge_append: capacity must fit int32 (counts are int32)");
  MergeArgs a{map_geometry, map_colors, with_ccounts ? 1 : 0, counts_in, counts_out, capacity, rgb, rgb_bstride, B, H,
              W, carve(workspace, B, H, W), overflow_flag, assoc_out};
  return launch_merge_append(a, (cudaStream_t)stream);
}

extern "C" int gsx_fusion_merge_append_bwd(const int32_t *assoc, const int32_t *counts_in, const float *map_geometry,
                                           const float *map_colors, int with_ccounts, int64_t capacity_in,
                                           const float *g_geometry, const float *g_colors, int64_t capacity_out,
                                           const float *gvertex, const float *gnormal, const float *rgb,
                                           const float *vertex, int B, int H, int W, double sigma,
                                           float *d_map_geometry, float *d_map_colors, float *d_gvertex,
                                           float *d_gnormal, float *d_rgb, float *d_vertex, void *stream) {
  GSX_CHECK_ARG(B >= 0 && H >= 2 && W >= 2, "gsx_fusion_merge_append_bwd: bad extents B=%d H=%d W=%d", B, H, W);
  if (B == 0) return 0;
  GSX_CHECK_ARG(assoc && counts_in && gvertex && gnormal && rgb && vertex, "gsx_fusion_merge_append_bwd: null input");
  GSX_CHECK_ARG(d_gvertex && d_gnormal && d_rgb && d_vertex, "gsx_fusion_merge_append_bwd: null frame gradient");
  GSX_CHECK_ARG(capacity_in == 0 || (map_geometry && map_colors && d_map_geometry && d_map_colors),
                "gsx_fusion_merge_append_bwd: null map pointer");
  GSX_CHECK_ARG(aligned16(map_geometry) && aligned16(map_colors) && aligned16(g_geometry) && aligned16(g_colors) &&
                    aligned16(d_map_geometry) && aligned16(d_map_colors),
                "gsx_fusion_merge_append_bwd: map rows must be 16-byte aligned");
  MergeBwdArgs a{assoc, counts_in, map_geometry, map_colors, with_ccounts ? 1 : 0, capacity_in, g_geometry, g_colors,
                 capacity_out, gvertex, gnormal, rgb, vertex, d_map_geometry, d_map_colors, d_gvertex, d_gnormal, d_rgb,
                 d_vertex, B, H * W, (float)(2.0 * (sigma * sigma))};
  cudaStream_t st = (cudaStream_t)stream;
  if (capacity_in > 0) {
    k_merge_bwd_rows<<<dim3((unsigned)((capacity_in + 255) / 256), (unsigned)B), 256, 0, st>>>(a);
    GSX_CHECK_LAUNCH("gsx_fusion_merge_append_bwd(rows)");
  }
  k_merge_bwd_pixels<<<dim3((unsigned)((a.P + 255) / 256), (unsigned)B), 256, 0, st>>>(a);
  GSX_CHECK_LAUNCH("gsx_fusion_merge_append_bwd(pixels)");
  return 0;
}
