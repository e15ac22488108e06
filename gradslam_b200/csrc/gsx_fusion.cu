// Fused PointFusion map update for sm_100a.
//   k_project_select  (K2+K3)  one thread per map point: project into the live camera, frustum / distance /
//                              normal tests, then a 128-bit atomic arg-min per pixel on the key
//                              (1/ccount, ray distance, point index).
//   k_merge_append    (K4)     one thread per pixel: confidence-weighted merge of the selected map point,
//                              or stable append of unmatched valid pixels (single-pass decoupled look-back
//                              scan, row-major order per batch element).  No float atomics anywhere.
// Reference op chains: gradslam/slam/fusionutils.py:198-722 (see include/gsx.h).
#include "gsx_common.cuh"
#include "gsx_exp.cuh"
#include "gsx_thresholds.h"
#include "../../include/gsx.h"

// Timing-only ablations (WRONG results; never set in a product build): where does the time go?
//   GSX_K4_ABLATE: 1 = matched rows are not read, 2 = matched rows are neither read nor written, 3 = no frame sample
//                  (no depth stencil, no vertex / normal math), 4 = arg-min records are not cleared
//   GSX_K2_ABLATE: 1 = no 128-bit CAS, 2 = no frame sample (depth gather + vertex / normal math)
#ifndef GSX_K4_ABLATE
#define GSX_K4_ABLATE 0
#endif
#ifndef GSX_K2_ABLATE
#define GSX_K2_ABLATE 0
#endif

namespace gsx {

constexpr int kBlock = 256;
#ifndef GSX_KPIX
#define GSX_KPIX 2
#endif
#ifndef GSX_K4_BLOCK
#define GSX_K4_BLOCK 256
#endif
constexpr int kMB = GSX_K4_BLOCK;         // threads per CTA of the merge/append kernel
constexpr int kTilePix = kMB * GSX_KPIX;  // pixels per merge tile (must equal kMergeTile)

// ---- workspace layout -----------------------------------------------------------------------------------
//   [0, B*P*16)                       U128 best[B][P]     complemented arg-min records (0 = empty)
//   then  uint64 tile_state[B][T]     (epoch<<34 | flag<<32 | value), T = ceil(P / kTilePix)
//   then  uint32 ticket[B]            dynamic tile ids (monotonic; tile = ticket - (epoch-1)*T)
//   then  uint64 stats[B][2]          running totals: {active map points (in frustum), merged points}
struct Workspace {
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
  return align_up(B * P * 16, 256) + align_up(B * tiles * 8, 256) + align_up((int64_t)B * 4, 256);
}

inline int64_t workspace_bytes(int B, int H, int W) {
  const int64_t P = (int64_t)H * W;
  const int64_t tiles = (P + kTilePix - 1) / kTilePix;
  return align_up(B * P * 16, 256) + align_up(B * tiles * 8, 256) + align_up((int64_t)B * 4, 256) +
         align_up((int64_t)B * 16, 256);
}

// ---- K2 + K3 ------------------------------------------------------------------------------------------
struct ProjectArgs {
  const float *pts, *nrm, *cc;  // geo32 layout: pts = geometry rows (B,cap,8), nrm = cc = null
  const int32_t *counts;
  int64_t cap;
  const float *poses;
  int64_t pose_bstride;
  const float *K;
  int64_t K_bstride;
  const float *gv, *gn;  // (B,H,W,3) materialised frame maps, or null: sample the depth image on the fly
  const float *depth;    // (B,H,W) live depth (used when gv/gn are null)
  int64_t depth_bstride;
  int B, H, W;
  float dist_th, dot_th, u_hi, v_hi;  // u_hi = float(W - 0.999), v_hi = float(H - 0.999)
  U128 *best;
  unsigned long long *stats;
  float d2_max;  // largest float x with sqrtf(x) < dist_th (-1 if none); used by GSX_K2_FASTTEST
};

#ifndef GSX_K2_MINB
#define GSX_K2_MINB 4
#endif
// Decision-exact shortcuts in K2 (round-2 candidate, OFF by default, not yet timed).  K2 only DECIDES with the frame
// normal and with sqrt(d2); only d2 itself enters the arg-min key.  So
//   * sqrtf(d2) < dist_th  becomes  d2 <= d2_max, with d2_max the largest float whose correctly rounded square root is
//     below dist_th (found on the host; sqrt is monotonic, so the decision is identical);
//   * the normal test n_frame . n_map > dot_th is first evaluated without normalising the frame normal
//     (((R c) . m) * rsqrt(|c|^2), no IEEE square root, no three IEEE divisions) and only when that value lies within a
//     guard band of dot_th - or the cross product is degenerate - is the canonical chain evaluated.
// The maps must stay bit-identical (tests/test_gpu_pointfusion.py, test_gpu_fullsize.py).
#ifndef GSX_K2_FASTTEST
#define GSX_K2_FASTTEST 0
#endif

struct MapPoint {  // everything K2 needs from one map row
  float px, py, pz, mx, my, mz, cc;
};
// kGeo: the "geo32" row layout under study for round 2 - geometry rows (px,py,pz,nx,ny,nz,cc,0) of exactly one 32-byte
// sector in `pts` (two 128-bit loads), colours in their own array; nrm / cc are unused.  Same values, same results.
template <bool kGeo>
__device__ __forceinline__ MapPoint load_map_point(const float *pts, const float *nrm, const float *cc, int64_t n) {
  MapPoint m;
  if (kGeo) {
    const float4 g0 = __ldg(reinterpret_cast<const float4 *>(pts + n * 8));
    const float4 g1 = __ldg(reinterpret_cast<const float4 *>(pts + n * 8 + 4));
    m.px = g0.x; m.py = g0.y; m.pz = g0.z;
    m.mx = g0.w; m.my = g1.x; m.mz = g1.y;
    m.cc = g1.z;
    return m;
  }
  m.px = __ldg(pts + n * 3);
  m.py = __ldg(pts + n * 3 + 1);
  m.pz = __ldg(pts + n * 3 + 2);
  m.mx = __ldg(nrm + n * 3);
  m.my = __ldg(nrm + n * 3 + 1);
  m.mz = __ldg(nrm + n * 3 + 2);
  m.cc = __ldg(cc + n);
  return m;
}

// The kernel is bound by (threads in flight) / (length of the dependent memory chain), not by bytes: ncu shows
// ~50 % issue utilisation and ~1.5 TB/s of DRAM traffic.  So the chain is kept as short as possible:
//   * the map row of the NEXT grid-stride iteration (position, normal, confidence) is fetched while the current
//     point is processed (software pipelining) - normal / confidence are read for every point, also the ~35 %
//     outside the frustum, which costs bytes but removes a round trip;
//   * the depth values under the projection (centre, right, below) are requested together;
//   * the result of the 128-bit CAS is only looked at one iteration later.
template <bool kFused, bool kGeo = false>
__global__ void __launch_bounds__(kBlock, GSX_K2_MINB) k_project_select(ProjectArgs a) {
  __shared__ Rigid s_pose, s_tinv;
  __shared__ float s_k[12];
  __shared__ KInv s_kinv;
  const int b = blockIdx.y;
  const int count = a.counts[b];
  if ((int64_t)blockIdx.x * kBlock >= count) return;
  if (threadIdx.x == 0) {
    s_pose = load_rigid(a.poses + b * a.pose_bstride);
    s_tinv = rigid_inverse(s_pose);
  }
  if (threadIdx.x >= 32 && threadIdx.x < 44) s_k[threadIdx.x - 32] = __ldg(a.K + b * a.K_bstride + (threadIdx.x - 32));
  if (threadIdx.x == 64) s_kinv = load_kinv(a.K + b * a.K_bstride);
  __syncthreads();
  const int P = a.H * a.W;
  const float *pts = a.pts + (int64_t)b * a.cap * (kGeo ? 8 : 3);
  const float *nrm = kGeo ? nullptr : a.nrm + (int64_t)b * a.cap * 3;
  const float *cc = kGeo ? nullptr : a.cc + (int64_t)b * a.cap;
  const float *gv = kFused ? nullptr : a.gv + (int64_t)b * P * 3;
  const float *gn = kFused ? nullptr : a.gn + (int64_t)b * P * 3;
  const float *dimg = a.depth + b * a.depth_bstride;
  U128 *best = a.best + (int64_t)b * P;
  unsigned int n_active = 0;
  const int64_t stride = (int64_t)gridDim.x * kBlock;
  int64_t n = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  U128 mine{0ull, 0ull}, old{0ull, 0ull};
  int pend_pix = -1;
  {
  MapPoint cur = load_map_point<kGeo>(pts, nrm, cc, n < count ? n : 0);
  for (; n < count; n += stride) {
    const MapPoint m = cur;
    const int64_t nn = n + stride;
    if (nn < count) cur = load_map_point<kGeo>(pts, nrm, cc, nn);  // in flight while this point is processed
    // world -> camera (pointclouds.py:526-573), then pinhole projection with the 4x4 K on the homogeneous
    // point (projutils.py:92-238): z == 0 divides by 1.
    const float3 q = rigid_apply(s_tinv, m.px, m.py, m.pz);
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
      float3 fv, fnm;
#if GSX_K2_FASTTEST
      bool fast_decided = false, fast_similar = false;  // the tests were already made
#endif
      if (kFused && GSX_K2_ABLATE == 2) {
        fv = make_float3(m.px, m.py, m.pz);
        fnm = make_float3(m.mx, m.my, m.mz);
#if GSX_K2_FASTTEST
      } else if (kFused) {
        // vertex exactly as frame_sample (it enters d2 and hence the key); the normal only decides
        const FrameSample f = frame_sample<false>(dimg, s_kinv, &s_pose, h, w, a.H, a.W);
        // (unconditional, so that the neighbour depths are requested together with the centre: one round trip)
        const float3 c = frame_cross(dimg, s_kinv, h, w, a.H, a.W, f.v);
        fv = f.gv;
        const float dx = fv.x - m.px, dy = fv.y - m.py, dz = fv.z - m.pz;
        const float d2f = (dx * dx + dy * dy) + dz * dz;
        bool similar = false;
        if (d2f <= a.d2_max) {
          const float vf = f.d > 0.0f ? 1.0f : 0.0f;
          const float c2 = (c.x * c.x + c.y * c.y) + c.z * c.z;
          bool decided = false;
          if (vf != 0.0f && c2 > 1e-30f && c2 < 1e30f) {
            const float3 rc = rotate(s_pose, c.x, c.y, c.z);
            const float approx = ((rc.x * m.mx + rc.y * m.my) + rc.z * m.mz) * rsqrtf(c2);
            const float guard = 1e-4f * (1.0f + (fabsf(m.mx) + fabsf(m.my)) + fabsf(m.mz));
            if (fabsf(approx - a.dot_th) > guard) {  // (false for NaN: falls through to the canonical chain)
              similar = approx > a.dot_th;
              decided = true;
            }
          }
          if (!decided) {
            const float3 nl = normalize_masked(c, vf);
            const float3 gn = rotate(s_pose, nl.x, nl.y, nl.z);
            similar = ((gn.x * m.mx + gn.y * m.my) + gn.z * m.mz) > a.dot_th;
          }
        }
        // hand the decision to the common code below: a normal that passes / fails the test by construction
        fnm = similar ? make_float3(m.mx, m.my, m.mz) : make_float3(0.f, 0.f, 0.f);
        fast_decided = true;
        fast_similar = similar;
#else
      } else if (kFused) {
        const FrameSample f = frame_sample<true>(dimg, s_kinv, &s_pose, h, w, a.H, a.W);
        fv = f.gv;
        fnm = f.gn;
#endif
      } else {
        const float *g = gv + (int64_t)pix * 3, *t = gn + (int64_t)pix * 3;
        fv = make_float3(__ldg(g), __ldg(g + 1), __ldg(g + 2));
        fnm = make_float3(__ldg(t), __ldg(t + 1), __ldg(t + 2));
      }
      // are_points_close (fusionutils.py:130): ||frame - map|| < dist_th
      const float dx = fv.x - m.px, dy = fv.y - m.py, dz = fv.z - m.pz;
      const float d2 = (dx * dx + dy * dy) + dz * dz;
      // are_normals_similar (fusionutils.py:187-195): n_frame . n_map > dot_th
      const float dot = (fnm.x * m.mx + fnm.y * m.my) + fnm.z * m.mz;
#if GSX_K2_FASTTEST
      live = fast_decided ? fast_similar : ((sqrtf(d2) < a.dist_th) && (dot > a.dot_th));
#else
      live = (sqrtf(d2) < a.dist_th) && (dot > a.dot_th);
#endif
      if (pend_pix >= 0) {  // settle the previous candidate's CAS before re-using the slot
        atomic_max_rec128_finish(best + pend_pix, mine, old);
        pend_pix = -1;
      }
      if (live) {
        // sort key of find_best_unique_correspondences (fusionutils.py:491-517): 1/(cc+1e-20), then the squared
        // distance (map - frame)^2 (== d2: squares are sign-independent), then n.
        const float inv_cc = 1.0f / (m.cc + 1e-20f);
        // positive floats order like their bit patterns; flip negatives so the order stays total.
        unsigned int kb = __float_as_uint(inv_cc);
        kb = (kb & 0x80000000u) ? ~kb : (kb | 0x80000000u);
        const unsigned int rb = __float_as_uint(d2) | 0x80000000u;  // d2 >= 0
        const unsigned long long hi = ((unsigned long long)kb << 32) | rb;
        mine = U128{~(unsigned long long)n, ~hi};
#if GSX_K2_ABLATE == 1
        if (n == -1) best[pix] = mine;  // never true: keeps the key computation alive
#else
        old = cas128(best + pix, U128{0ull, 0ull}, mine);  // optimistic: most pixels see a single candidate
        pend_pix = pix;
#endif
      }
    }
  }
  }
  if (pend_pix >= 0) atomic_max_rec128_finish(best + pend_pix, mine, old);
  // bookkeeping for the roofline's algorithmic-byte count: one atomic per warp
  n_active = __reduce_add_sync(0xffffffffu, n_active);
  if ((threadIdx.x & 31) == 0 && n_active) atomicAdd(a.stats + 2 * b, (unsigned long long)n_active);
}

// ---- K4 -------------------------------------------------------------------------------------------------
struct MergeArgs {
  float *pts, *nrm, *col, *cc;  // geo32 layout (geo = true): pts = geometry rows (B,cap,8), nrm = cc = null
  const int32_t *counts_in;
  int32_t *counts_out;
  int64_t cap;
  const float *depth;
  int64_t depth_bstride;
  const float *rgb;
  int64_t rgb_bstride;
  const float *K;
  int64_t K_bstride;
  const float *gv, *gn;  // materialised (B,H,W,3) maps to merge / append, or null: sample the depth on the fly
  const float *poses;    // (used when gv/gn are null)
  int64_t pose_bstride;
  int B, H, W;
  float two_sigma_sq;
  Workspace ws;
  unsigned int epoch;
  int32_t *overflow;
  int32_t *assoc;  // optional (B,P): +row+1 appended at `row`, -(row+1) merged into `row`, 0 untouched (kAssoc only)
};

// alpha = clamp(exp(-|v|^2 / 2 sigma^2), 1e-7, 1.01) (fusionutils.py:69-72).  The exponential is evaluated in double and
// rounded once: that is the correctly rounded float32 exp (up to 2^-29 odds), so the CUDA path and the CPU oracle agree
// bit for bit and no later threshold / arg-min decision can flip because of a 1-ulp difference in a confidence weight.
#ifndef GSX_K4_CTA_DIV
#define GSX_K4_CTA_DIV 1  // image row / column of a pixel from one division per CTA instead of one per pixel
#endif
#ifndef GSX_K4_FAST_EXP
#define GSX_K4_FAST_EXP 1  // reduced-range float64 exp (gsx_exp.cuh); 0: library exp().  Same bits either way.
#endif
__device__ __forceinline__ float confidence_exp(float sq_norm, float two_sigma_sq) {
  const float x = (-sq_norm) / two_sigma_sq;
#if GSX_K4_FAST_EXP
  if (!(x >= -17.0f)) return 0.0f;  // exp(x) < 4.2e-8: clamps to 1e-7 below (also NaN, like fmaxf(NaN, 1e-7f))
  return exp_f32_via_f64(x);
#else
  return (float)exp((double)x);
#endif
}
__device__ __forceinline__ float confidence_alpha(float sq_norm, float two_sigma_sq) {
  return fminf(fmaxf(confidence_exp(sq_norm, two_sigma_sq), 1e-7f), 1.01f);
}

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

#ifndef GSX_KPIX
#define GSX_KPIX 2
#endif
constexpr int kPix = GSX_KPIX;             // pixels per thread
constexpr int kMergeTile = kMB * kPix;  // pixels per CTA
static_assert(kMergeTile == kTilePix, "workspace tile size");

// exclusive prefix of the new-point counts of all preceding tiles (decoupled look-back, one warp, 32
// predecessors per step)
__device__ __forceinline__ unsigned int lookback_warp(const unsigned long long *state, int tile, unsigned int epoch,
                                                      int lane) {
  unsigned int excl = 0;
  for (int base = tile - 1; base >= 0; base -= 32) {
    const int j = base - lane;
    unsigned long long s = 0ull;
    if (j >= 0) {
      do {
        s = ld_acquire_u64(state + j);
      } while ((unsigned int)(s >> 34) != epoch);
    }
    const bool is_prefix = (j >= 0) && (((s >> 32) & 3ull) == kFlagPrefix);
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

#ifndef GSX_K4_GEO_MINB
#define GSX_K4_GEO_MINB GSX_K4_MINB  // occupancy target of the geo32 layout-study instantiation (spills 24 B at 4)
#endif
template <bool kFused, bool kDoMerge, bool kAssoc = false, bool kGeo = false>
__global__ void __launch_bounds__(kMB, kGeo ? GSX_K4_GEO_MINB : GSX_K4_MINB) k_merge_append(MergeArgs a) {
  __shared__ Rigid s_pose;
  __shared__ int s_tile, s_h0, s_w0;  // tile id; image row / column of the tile's first pixel
  __shared__ int s_warp_sums[kPix][kMB / 32];
  __shared__ int s_excl;
  __shared__ KInv s_k;
  // batch element varies fastest in the grid: CTAs resident at the same time belong to different elements, so
  // each element's look-back chain only sees ~1/B of the in-flight tiles
  const int b = blockIdx.x % a.B;
  const int T = a.ws.tiles;
  if (threadIdx.x == 0) {
    // dynamic tile id: tiles start in ticket order, so every predecessor of a running tile is running or done
    const unsigned int t = atomicAdd(a.ws.ticket + b, 1u);
    s_tile = (int)(t - (a.epoch - 1u) * (unsigned int)T);
    s_h0 = (s_tile * kMergeTile) / a.W;  // one division per CTA instead of one per pixel
    s_w0 = s_tile * kMergeTile - s_h0 * a.W;
  }
  if (threadIdx.x == 32) s_k = load_kinv(a.K + b * a.K_bstride);
  if (kFused && threadIdx.x == 64) s_pose = load_rigid(a.poses + b * a.pose_bstride);
  const int count_in = a.counts_in[b];  // loaded early: its latency hides behind everything below
  __syncthreads();
  const int tile = s_tile;
  const int P = a.H * a.W;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  U128 *best = a.ws.best + (int64_t)b * P;
  const float *depth = a.depth + b * a.depth_bstride;

  int pix[kPix];
  U128 rec[kPix];
  float d[kPix];
  bool matched[kPix], is_new[kPix];
  int warp_excl[kPix];
#pragma unroll
  for (int j = 0; j < kPix; ++j) {
    pix[j] = tile * kMergeTile + j * kMB + threadIdx.x;
    rec[j] = U128{0ull, 0ull};
    d[j] = 0.0f;
    if (pix[j] < P) {
      rec[j] = best[pix[j]];
      d[j] = __ldg(depth + pix[j]);
    }
  }
  int n_matched = 0;
#pragma unroll
  for (int j = 0; j < kPix; ++j) {
    matched[j] = (rec[j].lo | rec[j].hi) != 0ull;
    if (matched[j] && GSX_K4_ABLATE != 4) best[pix[j]] = U128{0ull, 0ull};  // leave the workspace clean for the next frame
    is_new[j] = (pix[j] < P) && (d[j] > 0.0f) && !matched[j];
    n_matched += matched[j] ? 1 : 0;
    // row-major order inside the tile: chunk j (256 consecutive pixels), then warp, then lane
    const unsigned int ballot = __ballot_sync(0xffffffffu, is_new[j]);
    warp_excl[j] = __popc(ballot & ((1u << lane) - 1u));
    if (lane == 0) s_warp_sums[j][warp] = __popc(ballot);
  }
  n_matched = __reduce_add_sync(0xffffffffu, n_matched);
  if (lane == 0 && n_matched) atomicAdd(a.ws.stats + 2 * b + 1, (unsigned long long)n_matched);
  __syncthreads();
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
  if (threadIdx.x == 0 && tile + 1 < T) st_release_u64(state + tile, pack_state(a.epoch, kFlagAgg, (unsigned)block_total));

  float *pts = a.pts + (int64_t)b * a.cap * (kGeo ? 8 : 3);  // geo32: geometry rows (px,py,pz,nx,ny,nz,cc,0)
  float *nrm = kGeo ? nullptr : a.nrm + (int64_t)b * a.cap * 3;
  float *col = a.col + (int64_t)b * a.cap * 3;
  float *cc = kGeo ? pts : (a.cc ? a.cc + (int64_t)b * a.cap : nullptr);  // geo32: only the null test is used
  const KInv k = s_k;
  const float *gvb = kFused ? nullptr : a.gv + (int64_t)b * P * 3;
  const float *gnb = kFused ? nullptr : a.gn + (int64_t)b * P * 3;
  const float *rgb = a.rgb + b * a.rgb_bstride;

  // per-pixel frame sample (loads of the 4 pixels are independent)
  float alpha[kPix];
  float3 fp[kPix], fn[kPix], fc[kPix];
#pragma unroll
  for (int j = 0; j < kPix; ++j) {
    if ((kDoMerge && matched[j]) || is_new[j]) {
      const float *c = rgb + (int64_t)pix[j] * 3;
      fc[j] = make_float3(__ldg(c), __ldg(c + 1), __ldg(c + 2));
      if (kFused) {
        int h, w;
        if (GSX_K4_CTA_DIV && a.W >= kMergeTile) {  // rows at least one tile wide: the tile wraps at most once
          h = s_h0;
          w = s_w0 + j * kMB + (int)threadIdx.x;
          if (w >= a.W) {
            w -= a.W;
            ++h;
          }
        } else {
          h = pix[j] / a.W;
          w = pix[j] - h * a.W;
        }
#if GSX_K4_ABLATE == 3
        FrameSample f;
        f.gv = make_float3((float)w, (float)h, d[j]);
        f.gn = make_float3(0.f, 0.f, 1.f);
        f.v = f.gv;
#else
        const FrameSample f = frame_sample<true>(depth, k, &s_pose, h, w, a.H, a.W);
#endif
        fp[j] = f.gv;
        fn[j] = f.gn;
        // alpha from the LOCAL vertex (fusionutils.py:657, 69-72)
        const float s = (f.v.x * f.v.x + f.v.y * f.v.y) + f.v.z * f.v.z;
        alpha[j] = confidence_alpha(s, a.two_sigma_sq);
      } else {
        const float *g = gvb + (int64_t)pix[j] * 3;
        const float *q = gnb + (int64_t)pix[j] * 3;
        fp[j] = make_float3(__ldg(g), __ldg(g + 1), __ldg(g + 2));
        fn[j] = make_float3(__ldg(q), __ldg(q + 1), __ldg(q + 2));
        const int h = pix[j] / a.W, w = pix[j] - h * a.W;
        const float3 v = backproject(k, (float)w, (float)h, d[j]);
        const float s = (v.x * v.x + v.y * v.y) + v.z * v.z;
        alpha[j] = confidence_alpha(s, a.two_sigma_sq);
      }
    }
  }
  // matched map rows: issue all loads first, then the arithmetic and the stores
  float mp[kPix][10];
#pragma unroll
  for (int j = 0; j < kPix; ++j) {
    if (kDoMerge && matched[j] && cc && GSX_K4_ABLATE != 2) {
      const int64_t n = (int64_t)(~rec[j].lo);
      if (kGeo) {
        const float4 g0 = *reinterpret_cast<const float4 *>(pts + n * 8);
        const float4 g1 = *reinterpret_cast<const float4 *>(pts + n * 8 + 4);
        mp[j][0] = g0.x; mp[j][1] = g0.y; mp[j][2] = g0.z;
        mp[j][3] = g0.w; mp[j][4] = g1.x; mp[j][5] = g1.y;
        mp[j][9] = g1.z;
#pragma unroll
        for (int q = 0; q < 3; ++q) mp[j][6 + q] = col[n * 3 + q];
        continue;
      }
#pragma unroll
      for (int q = 0; q < 3; ++q) {
#if GSX_K4_ABLATE == 1
        mp[j][q] = mp[j][3 + q] = mp[j][6 + q] = mp[j][9] = 1.0f;
        continue;
#endif
        mp[j][q] = pts[n * 3 + q];
        mp[j][3 + q] = nrm[n * 3 + q];
        mp[j][6 + q] = col[n * 3 + q];
      }
      if (GSX_K4_ABLATE != 1) mp[j][9] = cc[n];
    }
  }
#pragma unroll
  for (int j = 0; j < kPix; ++j) {
    if (kDoMerge && matched[j] && cc && GSX_K4_ABLATE != 2) {
      // confidence-weighted running mean (fusionutils.py:678-699); exactly one pixel owns this map row
      const int64_t n = (int64_t)(~rec[j].lo);
      const float c0 = mp[j][9];
      const float tot = c0 + alpha[j];
      const float inv = 1.0f / ((tot == 0.0f) ? 1.0f : tot);
      if (kGeo) {
        float4 g0, g1;
        g0.x = ((c0 * mp[j][0]) + (alpha[j] * fp[j].x)) * inv;
        g0.y = ((c0 * mp[j][1]) + (alpha[j] * fp[j].y)) * inv;
        g0.z = ((c0 * mp[j][2]) + (alpha[j] * fp[j].z)) * inv;
        g0.w = ((c0 * mp[j][3]) + (alpha[j] * fn[j].x)) * inv;
        g1.x = ((c0 * mp[j][4]) + (alpha[j] * fn[j].y)) * inv;
        g1.y = ((c0 * mp[j][5]) + (alpha[j] * fn[j].z)) * inv;
        g1.z = tot;
        g1.w = 0.0f;
        *reinterpret_cast<float4 *>(pts + n * 8) = g0;
        *reinterpret_cast<float4 *>(pts + n * 8 + 4) = g1;
        col[n * 3 + 0] = ((c0 * mp[j][6]) + (alpha[j] * fc[j].x)) * inv;
        col[n * 3 + 1] = ((c0 * mp[j][7]) + (alpha[j] * fc[j].y)) * inv;
        col[n * 3 + 2] = ((c0 * mp[j][8]) + (alpha[j] * fc[j].z)) * inv;
        continue;
      }
      pts[n * 3 + 0] = ((c0 * mp[j][0]) + (alpha[j] * fp[j].x)) * inv;
      pts[n * 3 + 1] = ((c0 * mp[j][1]) + (alpha[j] * fp[j].y)) * inv;
      pts[n * 3 + 2] = ((c0 * mp[j][2]) + (alpha[j] * fp[j].z)) * inv;
      if (kAssoc) a.assoc[(int64_t)b * P + pix[j]] = -(int32_t)(n + 1);
      nrm[n * 3 + 0] = ((c0 * mp[j][3]) + (alpha[j] * fn[j].x)) * inv;
      nrm[n * 3 + 1] = ((c0 * mp[j][4]) + (alpha[j] * fn[j].y)) * inv;
      nrm[n * 3 + 2] = ((c0 * mp[j][5]) + (alpha[j] * fn[j].z)) * inv;
      col[n * 3 + 0] = ((c0 * mp[j][6]) + (alpha[j] * fc[j].x)) * inv;
      col[n * 3 + 1] = ((c0 * mp[j][7]) + (alpha[j] * fc[j].y)) * inv;
      col[n * 3 + 2] = ((c0 * mp[j][8]) + (alpha[j] * fc[j].z)) * inv;
      cc[n] = tot;
    }
  }

  // decoupled look-back (warp 0): exclusive prefix of new-point counts over preceding tiles of this element
  if (warp == 0) {
    const unsigned int excl = lookback_warp(state, tile, a.epoch, lane);
    if (lane == 0) {
      if (tile + 1 < T) st_release_u64(state + tile, pack_state(a.epoch, kFlagPrefix, excl + (unsigned)block_total));
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
      if (kGeo && n < a.cap) {
        *reinterpret_cast<float4 *>(pts + n * 8) = make_float4(fp[j].x, fp[j].y, fp[j].z, fn[j].x);
        *reinterpret_cast<float4 *>(pts + n * 8 + 4) = make_float4(fn[j].y, fn[j].z, alpha[j], 0.0f);
        col[n * 3 + 0] = fc[j].x; col[n * 3 + 1] = fc[j].y; col[n * 3 + 2] = fc[j].z;
      } else if (n < a.cap) {
        pts[n * 3 + 0] = fp[j].x; pts[n * 3 + 1] = fp[j].y; pts[n * 3 + 2] = fp[j].z;
        nrm[n * 3 + 0] = fn[j].x; nrm[n * 3 + 1] = fn[j].y; nrm[n * 3 + 2] = fn[j].z;
        col[n * 3 + 0] = fc[j].x; col[n * 3 + 1] = fc[j].y; col[n * 3 + 2] = fc[j].z;
        if (cc) cc[n] = alpha[j];
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

int launch_project_select(const ProjectArgs &a, int64_t max_count, cudaStream_t stream, bool geo = false) {
  if (a.B == 0 || max_count <= 0) return 0;
  const int64_t chunk = (int64_t)kBlock;
  int64_t bx = (max_count + chunk - 1) / chunk;
#ifndef GSX_K2_CTAS_PER_SM
#define GSX_K2_CTAS_PER_SM 8
#endif
  const int64_t cap_blocks = (int64_t)kNumSMs * GSX_K2_CTAS_PER_SM;  // grid-stride beyond this many CTAs per SM
  if (bx * a.B > cap_blocks) bx = (cap_blocks + a.B - 1) / a.B;
  if (bx < 1) bx = 1;
  if (geo)  // layout study: fused frame sampling only
    k_project_select<true, true><<<dim3((unsigned)bx, (unsigned)a.B), kBlock, 0, stream>>>(a);
  else if (a.gv)
    k_project_select<false><<<dim3((unsigned)bx, (unsigned)a.B), kBlock, 0, stream>>>(a);
  else
    k_project_select<true><<<dim3((unsigned)bx, (unsigned)a.B), kBlock, 0, stream>>>(a);
  GSX_CHECK_LAUNCH("gsx_fusion_project_select");
  return 0;
}

int launch_merge_append(const MergeArgs &a, cudaStream_t stream, bool geo = false) {
  if (a.B == 0) return 0;
  const dim3 grid((unsigned)(a.ws.tiles * a.B));
  if (geo) k_merge_append<true, true, false, true><<<grid, kMB, 0, stream>>>(a);  // layout study
  else if (a.assoc) k_merge_append<false, true, true><<<grid, kMB, 0, stream>>>(a);  // differentiable forward (maps given)
  else if (a.gv) k_merge_append<false, true><<<grid, kMB, 0, stream>>>(a);
  else k_merge_append<true, true><<<grid, kMB, 0, stream>>>(a);
  GSX_CHECK_LAUNCH("gsx_fusion_merge_append");
  return 0;
}

// ---- K4 backward ------------------------------------------------------------------------------------------
// The differentiable forward (k_merge_append<false, true, true>) leaves, per pixel, where its sample went:
// merged into map row n (assoc = -(n+1)), appended as row n (assoc = n+1) or dropped (0).  With the pre-merge map
// and the frame values the backward is a pure per-pixel / per-row map - no atomics, no scan:
//   merged   out = (c*m + a*f) * inv,  inv = 1/(c+a)   (fusionutils.py:678-699)
//            d m = g*c*inv      d f = g*a*inv      d c += g*(m*inv - num*inv^2)      d a += g*(f*inv - num*inv^2)
//            cc_out = c + a  =>  d c += g_cc,  d a += g_cc
//   appended out = f, cc_out = a   =>  d f = g,  d a = g_cc
//   alpha    a = clamp(exp(-|v|^2 / 2 sigma^2), 1e-7, 1.01) (fusionutils.py:69-72)  =>  d v = d a * e * (-2 v / 2 sigma^2)
//            inside the clamp range, 0 outside.
struct MergeBwdArgs {
  const int32_t *assoc;
  const int32_t *counts_in;
  const float *pts, *nrm, *col, *cc;  // pre-merge map (B, cap_in, .)
  int64_t cap_in;
  const float *g_pts, *g_nrm, *g_col, *g_cc;  // upstream gradients (B, cap_out, .); null = zero
  int64_t cap_out;
  const float *gv, *gn, *rgb, *vloc;  // frame values (B, P, 3)
  float *d_pts, *d_nrm, *d_col, *d_cc;  // (B, cap_in, .)
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
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    a.d_pts[ri * 3 + q] = (live && a.g_pts) ? a.g_pts[ro * 3 + q] : 0.0f;
    a.d_nrm[ri * 3 + q] = (live && a.g_nrm) ? a.g_nrm[ro * 3 + q] : 0.0f;
    a.d_col[ri * 3 + q] = (live && a.g_col) ? a.g_col[ro * 3 + q] : 0.0f;
  }
  if (a.d_cc) a.d_cc[ri] = (live && a.g_cc) ? a.g_cc[ro] : 0.0f;
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
    for (int q = 0; q < 3; ++q) {
      g[q] = a.g_pts ? a.g_pts[ro * 3 + q] : 0.0f;
      g[3 + q] = a.g_nrm ? a.g_nrm[ro * 3 + q] : 0.0f;
      g[6 + q] = a.g_col ? a.g_col[ro * 3 + q] : 0.0f;
    }
    const float gcc = a.g_cc ? a.g_cc[ro] : 0.0f;
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
      const float c0 = a.cc[ri];
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
        m[q] = a.pts[ri * 3 + q];
        m[3 + q] = a.nrm[ri * 3 + q];
        m[6 + q] = a.col[ri * 3 + q];
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
      for (int q = 0; q < 3; ++q) {
        a.d_pts[ri * 3 + q] = dm[q];
        a.d_nrm[ri * 3 + q] = dm[3 + q];
        a.d_col[ri * 3 + q] = dm[6 + q];
        dgv[q] = df[q];
        dgn[q] = df[3 + q];
        dc[q] = df[6 + q];
      }
      a.d_cc[ri] = d_c0;
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

// One frame (K2 + K4, frame geometry sampled on the fly) for the batch elements [b0, b0 + nb) of a B_total-element
// problem, on `st`.  All pointers are the FULL-batch base pointers; batch elements are independent, so disjoint groups
// may run concurrently on different streams (gsx_pointfusion_sequence_gt).
int fusion_frame_group(float *pts, float *nrm, float *col, float *cc, const int32_t *cin, int32_t *cout, int64_t cap,
                       int64_t max_count, const float *poses, int64_t pose_bs, const float *K, int64_t K_bs,
                       const float *depth, int64_t d_bs, const float *rgb, int64_t rgb_bs, int B_total, int b0, int nb,
                       int H, int W, float dist_th, float dot_th, double sigma, void *workspace, uint32_t epoch,
                       int32_t *overflow, cudaStream_t st, bool geo) {
  const int64_t P = (int64_t)H * W;
  Workspace ws = carve(workspace, B_total, H, W);
  ws.best += (int64_t)b0 * P;
  ws.tile_state += (int64_t)b0 * ws.tiles;
  ws.ticket += b0;
  ws.stats += 2 * b0;
  // geo32 layout study (geo = true): pts = geometry rows (B,cap,8) holding normals and counts too; nrm = cc = null
  float *gp = pts + (int64_t)b0 * cap * (geo ? 8 : 3), *gn = geo ? nullptr : nrm + (int64_t)b0 * cap * 3;
  float *gc = col + (int64_t)b0 * cap * 3;
  float *gcc = geo ? gp : (cc ? cc + (int64_t)b0 * cap : nullptr);
  const float *gposes = poses + (int64_t)b0 * pose_bs, *gK = K + (int64_t)b0 * K_bs;
  const float *gdepth = depth + (int64_t)b0 * d_bs, *grgb = rgb + (int64_t)b0 * rgb_bs;
  if (max_count > 0 && gcc) {
    ProjectArgs pa{gp, gn, gcc, cin + b0, cap, gposes, pose_bs, gK, K_bs, nullptr, nullptr, gdepth, d_bs, nb, H, W,
                   dist_th, dot_th, (float)(W - 0.999), (float)(H - 0.999), ws.best, ws.stats,
                   sqrt_lt_threshold(dist_th)};
    const int rc = launch_project_select(pa, max_count, st, geo);
    if (rc) return rc;
  }
  MergeArgs ma{gp, gn, gc, gcc, cin + b0, cout + b0, cap, gdepth, d_bs, grgb, rgb_bs, gK, K_bs, nullptr, nullptr,
               gposes, pose_bs, nb, H, W, (float)(2.0 * (sigma * sigma)), ws, epoch, overflow, nullptr};
  return launch_merge_append(ma, st, geo);
}

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

extern "C" int gsx_fusion_project_select(const float *map_points, const float *map_normals,
                                         const float *map_ccounts, const int32_t *counts, int64_t capacity,
                                         int64_t max_count, const float *poses, int64_t pose_bstride,
                                         const float *intrinsics, int64_t K_bstride, const float *depth,
                                         int64_t depth_bstride, const float *gvertex, const float *gnormal, int B,
                                         int H, int W, float dist_th, float dot_th, void *workspace, void *stream) {
  GSX_CHECK_ARG(B >= 0 && H >= 2 && W >= 2, "gsx_fusion_project_select: bad extents B=%d H=%d W=%d", B, H, W);
  if (max_count <= 0 || B == 0) return 0;
  GSX_CHECK_ARG(map_points && map_normals && map_ccounts && counts, "gsx_fusion_project_select: null map pointer");
  GSX_CHECK_ARG(poses && intrinsics && workspace, "gsx_fusion_project_select: null frame pointer");
  GSX_CHECK_ARG((gvertex && gnormal) || (depth && !gvertex && !gnormal),
                "gsx_fusion_project_select: pass both frame maps, or neither together with the depth image");
  GSX_CHECK_ARG(max_count <= capacity, "gsx_fusion_project_select: max_count %lld > capacity %lld",
                (long long)max_count, (long long)capacity);
  const Workspace ws = carve(workspace, B, H, W);
  ProjectArgs a{map_points, map_normals, map_ccounts, counts, capacity, poses, pose_bstride, intrinsics, K_bstride,
                gvertex, gnormal, depth, depth_bstride, B, H, W, dist_th, dot_th, (float)(W - 0.999), (float)(H - 0.999), ws.best,
                ws.stats, sqrt_lt_threshold(dist_th)};
  return launch_project_select(a, max_count, (cudaStream_t)stream);
}

extern "C" int gsx_fusion_merge_append(float *map_points, float *map_normals, float *map_colors,
                                       float *map_ccounts, const int32_t *counts_in, int32_t *counts_out,
                                       int64_t capacity, const float *depth, int64_t depth_bstride,
                                       const float *rgb, int64_t rgb_bstride, const float *intrinsics,
                                       int64_t K_bstride, const float *poses, int64_t pose_bstride,
                                       const float *gvertex, const float *gnormal, int B, int H, int W,
                                       double sigma, void *workspace, uint32_t epoch, int32_t *overflow_flag,
                                       void *stream) {
  GSX_CHECK_ARG(B >= 0 && H >= 2 && W >= 2, "gsx_fusion_merge_append: bad extents B=%d H=%d W=%d", B, H, W);
  if (B == 0) return 0;
  GSX_CHECK_ARG(map_points && map_normals && map_colors && counts_in && counts_out,
                "gsx_fusion_merge_append: null map pointer");  // map_ccounts may be NULL (aggregation-only maps)
  GSX_CHECK_ARG(counts_in != counts_out, "gsx_fusion_merge_append: counts_in and counts_out must not alias");
  GSX_CHECK_ARG(depth && rgb && intrinsics && workspace && overflow_flag, "gsx_fusion_merge_append: null frame pointer");
  GSX_CHECK_ARG((gvertex && gnormal) || (poses && !gvertex && !gnormal),
                "gsx_fusion_merge_append: pass both frame maps, or neither together with the poses");
  GSX_CHECK_ARG(epoch >= 1 && epoch < (1u << 30), "gsx_fusion_merge_append: epoch out of range");
  GSX_CHECK_ARG(capacity <= 0x7fffffffll, "gsx_fusion_merge_append: capacity must fit int32 (counts are int32)");
  const Workspace ws = carve(workspace, B, H, W);
  MergeArgs a{map_points, map_normals, map_colors, map_ccounts, counts_in, counts_out, capacity, depth, depth_bstride,
              rgb, rgb_bstride, intrinsics, K_bstride, gvertex, gnormal, poses, pose_bstride, B, H, W,
              (float)(2.0 * (sigma * sigma)), ws, epoch, overflow_flag, nullptr};
  return launch_merge_append(a, (cudaStream_t)stream);
}

extern "C" int gsx_fusion_merge_append_fwd(float *map_points, float *map_normals, float *map_colors,
                                           float *map_ccounts, const int32_t *counts_in, int32_t *counts_out,
                                           int64_t capacity, const float *depth, int64_t depth_bstride,
                                           const float *rgb, int64_t rgb_bstride, const float *intrinsics,
                                           int64_t K_bstride, const float *gvertex, const float *gnormal, int B, int H,
                                           int W, double sigma, void *workspace, uint32_t epoch,
                                           int32_t *overflow_flag, int32_t *assoc_out, void *stream) {
  GSX_CHECK_ARG(B >= 0 && H >= 2 && W >= 2, "gsx_fusion_merge_append_fwd: bad extents B=%d H=%d W=%d", B, H, W);
  if (B == 0) return 0;
  GSX_CHECK_ARG(map_points && map_normals && map_colors && counts_in && counts_out,
                "gsx_fusion_merge_append_fwd: null map pointer");
  GSX_CHECK_ARG(counts_in != counts_out, "gsx_fusion_merge_append_fwd: counts_in and counts_out must not alias");
  GSX_CHECK_ARG(depth && rgb && intrinsics && workspace && overflow_flag && gvertex && gnormal && assoc_out,
                "gsx_fusion_merge_append_fwd: null frame pointer");
  GSX_CHECK_ARG(epoch >= 1 && epoch < (1u << 30), "gsx_fusion_merge_append_fwd: epoch out of range");
  GSX_CHECK_ARG(capacity <= 0x7fffffffll, "gsx_fusion_merge_append_fwd: capacity must fit int32 (counts are int32)");
  const Workspace ws = carve(workspace, B, H, W);
  MergeArgs a{map_points, map_normals, map_colors, map_ccounts, counts_in, counts_out, capacity, depth, depth_bstride,
              rgb, rgb_bstride, intrinsics, K_bstride, gvertex, gnormal, nullptr, 0, B, H, W,
              (float)(2.0 * (sigma * sigma)), ws, epoch, overflow_flag, assoc_out};
  return launch_merge_append(a, (cudaStream_t)stream);
}

extern "C" int gsx_fusion_merge_append_bwd(const int32_t *assoc, const int32_t *counts_in, const float *map_points,
                                           const float *map_normals, const float *map_colors, const float *map_ccounts,
                                           int64_t capacity_in, const float *g_points, const float *g_normals,
                                           const float *g_colors, const float *g_ccounts, int64_t capacity_out,
                                           const float *gvertex, const float *gnormal, const float *rgb,
                                           const float *vertex, int B, int H, int W, double sigma, float *d_map_points,
                                           float *d_map_normals, float *d_map_colors, float *d_map_ccounts,
                                           float *d_gvertex, float *d_gnormal, float *d_rgb, float *d_vertex,
                                           void *stream) {
  GSX_CHECK_ARG(B >= 0 && H >= 2 && W >= 2, "gsx_fusion_merge_append_bwd: bad extents B=%d H=%d W=%d", B, H, W);
  if (B == 0) return 0;
  GSX_CHECK_ARG(assoc && counts_in && gvertex && gnormal && rgb && vertex, "gsx_fusion_merge_append_bwd: null input");
  GSX_CHECK_ARG(d_gvertex && d_gnormal && d_rgb && d_vertex, "gsx_fusion_merge_append_bwd: null frame gradient");
  GSX_CHECK_ARG(capacity_in == 0 || (map_points && map_normals && map_colors && d_map_points && d_map_normals &&
                                     d_map_colors),
                "gsx_fusion_merge_append_bwd: null map pointer");
  GSX_CHECK_ARG((map_ccounts == nullptr) == (d_map_ccounts == nullptr),
                "gsx_fusion_merge_append_bwd: ccounts and their gradient must both be given or both be null");
  MergeBwdArgs a{assoc, counts_in, map_points, map_normals, map_colors, map_ccounts, capacity_in, g_points, g_normals,
                 g_colors, g_ccounts, capacity_out, gvertex, gnormal, rgb, vertex, d_map_points, d_map_normals,
                 d_map_colors, d_map_ccounts, d_gvertex, d_gnormal, d_rgb, d_vertex, B, H * W,
                 (float)(2.0 * (sigma * sigma))};
  cudaStream_t st = (cudaStream_t)stream;
  if (capacity_in > 0) {
    k_merge_bwd_rows<<<dim3((unsigned)((capacity_in + 255) / 256), (unsigned)B), 256, 0, st>>>(a);
    GSX_CHECK_LAUNCH("gsx_fusion_merge_append_bwd(rows)");
  }
  k_merge_bwd_pixels<<<dim3((unsigned)((a.P + 255) / 256), (unsigned)B), 256, 0, st>>>(a);
  GSX_CHECK_LAUNCH("gsx_fusion_merge_append_bwd(pixels)");
  return 0;
}
