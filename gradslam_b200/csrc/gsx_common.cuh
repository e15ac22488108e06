// Shared device helpers for libgsx (sm_100a).  Compiled with -fmad=false: every product and sum below
// is rounded separately to fp32, in the association order written, so results are bit-identical to the
// CPU oracle's canonical arithmetic (oracle/gsx_oracle.py).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>

namespace gsx {

void set_error(const char *fmt, ...);

#define GSX_CHECK_ARG(cond, ...)   \
  do {                             \
    if (!(cond)) {                 \
      gsx::set_error(__VA_ARGS__); \
      return 1;                    \
    }                              \
  } while (0)

#define GSX_CHECK_LAUNCH(name)                                                    \
  do {                                                                            \
    cudaError_t e_ = cudaGetLastError();                                          \
    if (e_ != cudaSuccess) {                                                      \
      gsx::set_error("%s: launch failed: %s", name, cudaGetErrorString(e_));     \
      return 2;                                                                   \
    }                                                                             \
  } while (0)

constexpr int kNumSMs = 148;  // B200

struct Rigid {  // row-major rotation + translation of a 4x4 rigid transform
  float r[9];
  float t[3];
};

__device__ __forceinline__ Rigid load_rigid(const float *__restrict__ T) {
  Rigid a;
  a.r[0] = __ldg(T + 0); a.r[1] = __ldg(T + 1); a.r[2] = __ldg(T + 2);  a.t[0] = __ldg(T + 3);
  a.r[3] = __ldg(T + 4); a.r[4] = __ldg(T + 5); a.r[5] = __ldg(T + 6);  a.t[1] = __ldg(T + 7);
  a.r[6] = __ldg(T + 8); a.r[7] = __ldg(T + 9); a.r[8] = __ldg(T + 10); a.t[2] = __ldg(T + 11);
  return a;
}

__device__ __forceinline__ float dot3(float a0, float a1, float a2, float b0, float b1, float b2) {
  return (a0 * b0 + a1 * b1) + a2 * b2;
}

// [R^T, (-R^T) t]   (kornia inverse_transformation as called at gradslam/slam/fusionutils.py:249)
__device__ __forceinline__ Rigid rigid_inverse(const Rigid &a) {
  Rigid o;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) o.r[i * 3 + j] = a.r[j * 3 + i];
    o.t[i] = dot3(-a.r[0 * 3 + i], -a.r[1 * 3 + i], -a.r[2 * 3 + i], a.t[0], a.t[1], a.t[2]);
  }
  return o;
}

__device__ __forceinline__ float3 rigid_apply(const Rigid &a, float x, float y, float z) {
  float3 q;
  q.x = dot3(a.r[0], a.r[1], a.r[2], x, y, z) + a.t[0];
  q.y = dot3(a.r[3], a.r[4], a.r[5], x, y, z) + a.t[1];
  q.z = dot3(a.r[6], a.r[7], a.r[8], x, y, z) + a.t[2];
  return q;
}

__device__ __forceinline__ float3 rotate(const Rigid &a, float x, float y, float z) {
  float3 q;
  q.x = dot3(a.r[0], a.r[1], a.r[2], x, y, z);
  q.y = dot3(a.r[3], a.r[4], a.r[5], x, y, z);
  q.z = dot3(a.r[6], a.r[7], a.r[8], x, y, z);
  return q;
}

// closed-form inverse intrinsics (gradslam/geometry/projutils.py:405-450): eps added to fx, fy.
struct KInv {
  float k00, k02, k11, k12;
};
__device__ __forceinline__ KInv load_kinv(const float *__restrict__ K) {
  const float fx = __ldg(K + 0), fy = __ldg(K + 5), cx = __ldg(K + 2), cy = __ldg(K + 6);
  KInv k;
  k.k00 = 1.0f / (fx + 1e-6f);
  k.k11 = 1.0f / (fy + 1e-6f);
  k.k02 = (-1.0f * cx) / (fx + 1e-6f);
  k.k12 = (-1.0f * cy) / (fy + 1e-6f);
  return k;
}

// local vertex of pixel (u=w, v=h) with depth d, zeroed where d <= 0 (rgbdimages.py:672-679)
__device__ __forceinline__ float3 backproject(const KInv &k, float u, float v, float d) {
  const float vf = d > 0.0f ? 1.0f : 0.0f;
  float3 p;
  p.x = ((k.k00 * u + k.k02) * d) * vf;
  p.y = ((k.k11 * v + k.k12) * d) * vf;
  p.z = d * vf;
  return p;
}

// Vertex and normal of pixel (h,w) of depth image `dimg`, in the camera frame and (if pose != nullptr) in the
// world frame: the whole op chain of gradslam/structures/rgbdimages.py:643-762 for one pixel.  Used by K1 (to
// materialise maps) and, on the fly, by the fusion kernels, so both see bit-identical values.
//   v  = K^-1 (w,h,1) d, zeroed where d <= 0            n  = normalize(dh x dv) * valid(centre)
//   dh = forward difference along w (last column re-uses its neighbour's), dv along h likewise
//   gv = (R v + t) * valid                              gn = R n
struct FrameSample {
  float3 v, n, gv, gn;
  float d;
};

// a x b and |c| rounded exactly like the reference's CPU build (torch.cross / Tensor.norm on float32,
// rgbdimages.py:733-734): cross.x = fma(a.y, b.z, -(a.z * b.y)) - first product exact, second rounded - and
// |c| = sqrt(fma(c.z, c.z, fma(c.y, c.y, c.x * c.x))).  Where a valid pixel's right AND lower neighbours are both
// missing, a == b and the contracted cross product is rounding residue rather than 0; normalised, it decides whether
// map points match there, so the kernels and the oracle (oracle/normal_fma.c) reproduce it bit for bit.  The library
// is built with -fmad=false: only these explicit fused operations are fused.
__device__ __forceinline__ float3 cross_ref(float ax, float ay, float az, float bx, float by, float bz) {
  return make_float3(__fmaf_rn(ay, bz, -__fmul_rn(az, by)), __fmaf_rn(az, bx, -__fmul_rn(ax, bz)),
                     __fmaf_rn(ax, by, -__fmul_rn(ay, bx)));
}
__device__ __forceinline__ float norm_ref(const float3 &c) {
  return __fsqrt_rn(__fmaf_rn(c.z, c.z, __fmaf_rn(c.y, c.y, __fmul_rn(c.x, c.x))));
}

// un-normalised local normal dh x dv of pixel (h,w) given its own local vertex v
__device__ __forceinline__ float3 frame_cross(const float *__restrict__ dimg, const KInv &k, int h, int w, int H, int W,
                                              const float3 &v) {
  const int wa = (w < W - 1) ? w : w - 1;
  const int ha = (h < H - 1) ? h : h - 1;
  const float dr = __ldg(dimg + h * W + wa + 1);
  const float db = __ldg(dimg + (ha + 1) * W + w);
  const float3 a1 = backproject(k, (float)(wa + 1), (float)h, dr);
  const float3 b1 = backproject(k, (float)w, (float)(ha + 1), db);
  const float3 a0 = (wa == w) ? v : backproject(k, (float)wa, (float)h, __ldg(dimg + h * W + wa));
  const float3 b0 = (ha == h) ? v : backproject(k, (float)w, (float)ha, __ldg(dimg + ha * W + w));
  const float dhx = a1.x - a0.x, dhy = a1.y - a0.y, dhz = a1.z - a0.z;
  const float dvx = b1.x - b0.x, dvy = b1.y - b0.y, dvz = b1.z - b0.z;
  return cross_ref(dhx, dhy, dhz, dvx, dvy, dvz);
}

// normalize(c) * vf, the zero vector staying zero (rgbdimages.py:731-743)
__device__ __forceinline__ float3 normalize_masked(const float3 &c, float vf) {
  const float nrm = norm_ref(c);
  const float den = (nrm == 0.0f) ? 1.0f : nrm;
  return make_float3((c.x / den) * vf, (c.y / den) * vf, (c.z / den) * vf);
}

// local normal of pixel (h,w) given its own local vertex v and validity vf (1 or 0)
__device__ __forceinline__ float3 frame_normal(const float *__restrict__ dimg, const KInv &k, int h, int w, int H, int W,
                                               const float3 &v, float vf) {
  return normalize_masked(frame_cross(dimg, k, h, w, H, W, v), vf);
}

// The five depth values the sample of pixel (h,w) depends on: centre, the two ends of its horizontal difference
// (w_a, w_a+1) and of its vertical difference (h_a, h_a+1), where w_a = min(w, W-2), h_a = min(h, H-2).
struct DepthStencil {
  float c, l, r, u, d;
};
__device__ __forceinline__ DepthStencil load_stencil(const float *__restrict__ dimg, int h, int w, int H, int W) {
  const int wa = (w < W - 1) ? w : w - 1;
  const int ha = (h < H - 1) ? h : h - 1;
  DepthStencil s;
  s.c = __ldg(dimg + h * W + w);
  s.l = __ldg(dimg + h * W + wa);
  s.r = __ldg(dimg + h * W + wa + 1);
  s.u = __ldg(dimg + ha * W + w);
  s.d = __ldg(dimg + (ha + 1) * W + w);
  return s;
}
// frame_sample<true> evaluated from an already loaded stencil (bit-identical arithmetic)
__device__ __forceinline__ FrameSample frame_sample_from(const DepthStencil &t, const KInv &k, const Rigid *pose, int h,
                                                         int w, int H, int W) {
  FrameSample s;
  const int wa = (w < W - 1) ? w : w - 1;
  const int ha = (h < H - 1) ? h : h - 1;
  s.d = t.c;
  const float vf = t.c > 0.0f ? 1.0f : 0.0f;
  s.v = backproject(k, (float)w, (float)h, t.c);
  const float3 a1 = backproject(k, (float)(wa + 1), (float)h, t.r);
  const float3 b1 = backproject(k, (float)w, (float)(ha + 1), t.d);
  const float3 a0 = (wa == w) ? s.v : backproject(k, (float)wa, (float)h, t.l);
  const float3 b0 = (ha == h) ? s.v : backproject(k, (float)w, (float)ha, t.u);
  const float dhx = a1.x - a0.x, dhy = a1.y - a0.y, dhz = a1.z - a0.z;
  const float dvx = b1.x - b0.x, dvy = b1.y - b0.y, dvz = b1.z - b0.z;
  s.n = normalize_masked(cross_ref(dhx, dhy, dhz, dvx, dvy, dvz), vf);
  if (pose) {
    s.gv = rigid_apply(*pose, s.v.x, s.v.y, s.v.z);
    s.gv.x *= vf; s.gv.y *= vf; s.gv.z *= vf;
    s.gn = rotate(*pose, s.n.x, s.n.y, s.n.z);
  } else {
    s.gv = s.v;
    s.gn = s.n;
  }
  return s;
}

template <bool kWantNormal>
__device__ __forceinline__ FrameSample frame_sample(const float *__restrict__ dimg, const KInv &k, const Rigid *pose,
                                                    int h, int w, int H, int W) {
  FrameSample s;
  const float dc = __ldg(dimg + h * W + w);
  s.d = dc;
  const float vf = dc > 0.0f ? 1.0f : 0.0f;
  s.v = backproject(k, (float)w, (float)h, dc);
  s.n = kWantNormal ? frame_normal(dimg, k, h, w, H, W, s.v, vf) : make_float3(0.f, 0.f, 0.f);
  if (pose) {
    s.gv = rigid_apply(*pose, s.v.x, s.v.y, s.v.z);
    s.gv.x *= vf; s.gv.y *= vf; s.gv.z *= vf;
    s.gn = kWantNormal ? rotate(*pose, s.n.x, s.n.y, s.n.z) : s.n;
  } else {
    s.gv = s.v;
    s.gn = s.n;
  }
  return s;
}

// 128-bit record used for the per-pixel arg-min.  The workspace is zero-initialised and zero means
// "no candidate", so a key (hi:lo) is stored as its bitwise complement and the arg-min over keys
// becomes an atomic MAX over the stored 128-bit unsigned integers.
struct __align__(16) U128 {
  unsigned long long lo, hi;
};

__device__ __forceinline__ U128 cas128(U128 *addr, U128 expected, U128 desired) {
  U128 old;
  asm volatile(
      "{\n\t.reg .b128 e, d, o;\n\t"
      "mov.b128 e, {%2, %3};\n\t"
      "mov.b128 d, {%4, %5};\n\t"
      "atom.global.relaxed.gpu.cas.b128 o, [%6], e, d;\n\t"
      "mov.b128 {%0, %1}, o;\n\t}"
      : "=l"(old.lo), "=l"(old.hi)
      : "l"(expected.lo), "l"(expected.hi), "l"(desired.lo), "l"(desired.hi), "l"(addr)
      : "memory");
  return old;
}

// ---- L2 residency hints (createpolicy + .L2::cache_hint) -----------------------------------------------------------
// The per-frame arg-min records are hit by scattered read-modify-writes; if their sectors have left the L2 every one of
// them becomes a DRAM read + write-back.  Streaming data (map rows, colours) is therefore loaded evict-first and the
// records are touched evict-last, so the records of the frame in flight stay resident.
__device__ __forceinline__ unsigned long long l2_policy_evict_last() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ float4 ldg128_hint(const float *addr, unsigned long long pol) {
  float4 v;
  asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(addr), "l"(pol));
  return v;
}
__device__ __forceinline__ float2 ldg64_hint(const float *addr, unsigned long long pol) {
  float2 v;
  asm volatile("ld.global.nc.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;" : "=f"(v.x), "=f"(v.y) : "l"(addr), "l"(pol));
  return v;
}
__device__ __forceinline__ void stg128_hint(float *addr, const float4 &v, unsigned long long pol) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w), "l"(pol)
               : "memory");
}
__device__ __forceinline__ U128 load128_relaxed(const U128 *addr) {
  U128 v;
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.lo), "=l"(v.hi) : "l"(addr) : "memory");
  return v;
}

__device__ __forceinline__ bool rec_greater(const U128 &a, const U128 &b) {
  return a.hi > b.hi || (a.hi == b.hi && a.lo > b.lo);
}

// Finishes an arg-min update whose first (optimistic, expected = empty) CAS returned `old`.
__device__ __forceinline__ void atomic_max_rec128_finish(U128 *addr, const U128 &mine, U128 old) {
  U128 cur{0ull, 0ull};
  while (!(old.hi == cur.hi && old.lo == cur.lo)) {  // the CAS did not take effect
    cur = old;
    if (!rec_greater(mine, cur)) return;  // somebody better is already there
    old = cas128(addr, cur, mine);
  }
}

// arg-min over keys == max over complemented records
__device__ __forceinline__ void atomic_min_key128(U128 *addr, unsigned long long key_hi, unsigned long long key_lo) {
  const U128 mine{~key_lo, ~key_hi};
  U128 cur{0ull, 0ull};  // optimistic: most pixels see a single candidate, so expect "empty" first
  while (mine.hi > cur.hi || (mine.hi == cur.hi && mine.lo > cur.lo)) {
    const U128 old = cas128(addr, cur, mine);
    if (old.hi == cur.hi && old.lo == cur.lo) break;
    cur = old;
  }
}

}  // namespace gsx
