// rt_common.cuh — device-side data layout and small math helpers of the B200
// path-tracer core. Everything here is `__host__ __device__` so that
// tests/emu can single-step the per-lane code on the CPU while debugging; the
// product library only ever runs it on the GPU.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <string.h>

#include "../../include/rt_capi.h"

#define RT_HD __host__ __device__ __forceinline__
#define RT_D __device__ __forceinline__

#ifdef __CUDA_ARCH__
#define RT_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#define RT_FMIN(a, b) fminf((a), (b))
#define RT_FMAX(a, b) fmaxf((a), (b))
#define RT_F2I(x) __float_as_int(x)
#define RT_I2F(x) __int_as_float(x)
static __device__ __forceinline__ float rt_rsqrt_approx(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
#define RT_RSQRT(x) rt_rsqrt_approx(x)
#define RT_RCP(x) __frcp_rn(x)
// sqrt.approx (MUFU, ~1 ulp): every use is either statistical (samplers) or bounded at 1e-5
// relative (ray parameters); the IEEE sequence is 8 instructions and a slow-path branch
static __device__ __forceinline__ float rt_sqrt_approx(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
#define RT_SQRT(x) rt_sqrt_approx(x)
#define RT_MULHI(a, b) __umulhi((a), (b))
// fast division (MUFU.RCP + multiply, <= 2 ulp) for quantities whose parity bound is 1e-5 relative
// or statistical: the IEEE sequence costs ~12 instructions and a slow-path branch per use
// (rcp.approx.ftz is ONE MUFU; __fdividef adds a denormal-rescue sequence of 4 instructions)
static __device__ __forceinline__ float rt_rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
#define RT_FDIV(a, b) ((a) * rt_rcp_approx(b))
#else
static inline int rt_host_f2i(float x) { int i; memcpy(&i, &x, 4); return i; }
static inline float rt_host_i2f(int i) { float x; memcpy(&x, &i, 4); return x; }
#define RT_FMA(a, b, c) fmaf((a), (b), (c))
#define RT_FMIN(a, b) fminf((a), (b))
#define RT_FMAX(a, b) fmaxf((a), (b))
#define RT_F2I(x) rt_host_f2i(x)
#define RT_I2F(x) rt_host_i2f(x)
#define RT_RSQRT(x) (1.0f / sqrtf(x))
#define RT_RCP(x) (1.0f / (x))
#define RT_SQRT(x) sqrtf(x)
#define RT_MULHI(a, b) ((uint32_t)(((uint64_t)(a) * (uint64_t)(b)) >> 32))
#define RT_FDIV(a, b) ((a) / (b))
#endif

// Rarely executed or large straight-line code lives in out-of-line functions: the big general kernels are bound by
// instruction fetch, and register allocation of the hot loop is not shaped by cold code. Arguments and
// results travel BY VALUE (a reference would pin the caller's variables in local memory).
#ifdef __CUDA_ARCH__
#define RT_COLD __device__ __noinline__
#else
#define RT_COLD static inline
#endif

#define RT_NODE_SHIFT 5 // log2(sizeof packed node): node links are byte offsets

struct V3f {
  float x, y, z;
};
RT_HD V3f v3(float x, float y, float z) { V3f r = {x, y, z}; return r; }
RT_HD V3f operator+(V3f a, V3f b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
RT_HD V3f operator-(V3f a, V3f b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
RT_HD V3f operator-(V3f a) { return v3(-a.x, -a.y, -a.z); }
RT_HD V3f operator*(float t, V3f a) { return v3(t * a.x, t * a.y, t * a.z); }
RT_HD V3f operator*(V3f a, V3f b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
RT_HD float dot(V3f a, V3f b) { return RT_FMA(a.z, b.z, RT_FMA(a.y, b.y, a.x * b.x)); }
RT_HD V3f cross(V3f u, V3f w) {
  return v3(RT_FMA(u.y, w.z, -u.z * w.y), RT_FMA(u.z, w.x, -u.x * w.z), RT_FMA(u.x, w.y, -u.y * w.x));
}
// a + t*b
RT_HD V3f madd(V3f a, float t, V3f b) { return v3(RT_FMA(t, b.x, a.x), RT_FMA(t, b.y, a.y), RT_FMA(t, b.z, a.z)); }
RT_HD V3f normalize(V3f a) { float s = RT_RSQRT(dot(a, a)); return s * a; }
RT_HD V3f v3_from(const float *p) { return v3(p[0], p[1], p[2]); }
RT_HD V3f xyz(float4 q) { return v3(q.x, q.y, q.z); }

// ---------------------------------------------------------------- scene on device
// All arrays live in HBM (uploaded once per scene) and are staged into shared
// memory by the render kernel when they fit ("data layout").
//
//  nodes   float4[2*n_nodes]   {bmin.xyz, escape} {bmax.xyz, payload}   32 B/node
//                              escape / child links are BYTE offsets (index << 5)
//                              payload >= 0: first child; < 0: ~(first<<3 | count-1) into leaf_prims
//  leaf_prims int32[n_small]   RT_PRIM_IDs in depth-first leaf order
//  sph     float4[n_spheres]   {c0.xyz, radius}                         16 B
//  sph_k   float[n_spheres]    |c0|^2 - r^2 (big-sphere intersection, intersect.cuh)
//  sph_mv  float4[n_spheres]   {c1-c0, 1/(time1-time0)} (only if any sphere moves;
//                              .w == 0 marks a static sphere; time0 in sph_t0)
//  tri     float4[4*n_tris]    {N.xyz (un-normalised), v0.N}, then three edge
//                              planes {m_e.xyz, k_e}: inside <=> m_e.p - k_e >= 0
//  tri_n   float4[n_tris]      {unit N.xyz, 0} (shading only)
//  quad    float4[2*n_quads]   {k, a0, a1, as_float(axis)} {b0, b1, 0, 0}
//  *_mat   int32 per primitive material index
//  mats    float4[2*n_mats]    {albedo.rgb, as_float(type | texture<<8)} {albedo2.rgb, param}
//  big     int32[n_big]        RT_PRIM_IDs tested for every ray before traversal (spheres, triangles)
//  bigq    float4[2*n_bigq]    always-tested RECTS (room walls), decoded: {k, a0, a1, as_float(axis)}
//                              {b0, b1, as_float(RT_PRIM_ID), 0} - no id fetch / type dispatch per ray
//  media   float4[4*n_media]   {p0.xyz, as_float(shape)} {p1.xyz, -1/density} {offset.xyz, as_float(material)}
//                              {sin_y, cos_y, 0, 0}            (global memory; not in the BVH)
//  perlin_vec float4[256*n_perlin], perlin_perm uint8[768*n_perlin] (x, y, z tables)
//  images  DevImage[n_images]  8-bit RGB rows, top row first
//  inst    float4[4*n_inst]    per rt_instance: three rows {Rinv_k.xyz, tinv_k} of the WORLD -> OBJECT map
//                              (p_obj = Rinv p_world + tinv; rigid, so t is preserved and normals go back with
//                              the transpose), then {as_float(byte offset of the object's root node),
//                              as_float(group), 0, 0}, then the quantisation frame of the object's tree for the
//                              16-bit shared-memory node records: {qbase.xyz, 0} {qscale.xyz, 0} (RT_INST_STRIDE = 6)
//  box     float4[2*n_boxes]   {p0.xyz, as_float(first rect)} {p1.xyz, 0}: six consecutive rects that are the sides of
//                              an axis-aligned box in box.h's order (xy@z1, xy@z0, xz@y1, xz@y0, yz@x1, yz@x0). The
//                              builder gives the box ONE leaf (RT_PRIM_BOX) instead of six; the test evaluates the six
//                              rect formulas on the same numbers, so hits (rect id, t) are bit-identical.
// Two-level scenes (n_inst > 0): `nodes` holds the top-level tree first (root = node 0; its leaves name world
// primitives or RT_PRIM_INSTANCE ids) and then one tree per group in OBJECT space; a link that leaves a group's
// tree is RT_POP_LINK(n_nodes) = "back to the top level" (intersect.cuh, render_kernels.cuh).
struct DevImage {
  const uint8_t *rgb;
  int width, height;
};

struct DevScene {
  const float4 *nodes;
  const float4 *sph;
  const float *sph_k;   // |c0|^2 - r^2 per sphere (double precision on the host, rounded once)
  const float4 *sph_mv;
  const float *sph_t0;
  const float4 *tri;
  const float4 *tri_n;
  const float4 *quad;
  const int32_t *sph_mat;
  const int32_t *tri_mat;
  const int32_t *quad_mat;
  const float4 *mats;
  const int32_t *big;
  const int32_t *leaf_prims; // RT_PRIM_IDs in depth-first leaf order (leaves hold ranges of it)
  int n_nodes, n_spheres, n_tris, n_quads, n_mats, n_big;
  int any_moving;
  int node_stride; // bytes between the four quadrant-ordered copies of `nodes` (copy 0 first)
  const float4 *bigq;
  int n_bigq;
  const float4 *media;
  const float4 *perlin_vec;
  const uint8_t *perlin_perm;
  const DevImage *images;
  int n_media, n_perlin, n_images;
  const float4 *box;
  int n_boxes;
  const float4 *inst;
  const int32_t *groups; // rt_group[n_groups] as 8 ints each (brute-force parity hook)
  int n_inst;
};

#define RT_INST_STRIDE 6 // float4 per instance record
// link value that ends an object's (bottom-level) tree: one node past "traversal finished"
#define RT_POP_LINK(n_nodes) (((n_nodes) << RT_NODE_SHIFT) + (1 << RT_NODE_SHIFT))

// world -> object for an instance record (rigid: directions only rotate, t is preserved)
RT_HD V3f inst_point_to_object(const float4 *rec, V3f p) {
  const float4 a = rec[0], b = rec[1], c = rec[2];
  return v3(RT_FMA(a.z, p.z, RT_FMA(a.y, p.y, RT_FMA(a.x, p.x, a.w))), RT_FMA(b.z, p.z, RT_FMA(b.y, p.y, RT_FMA(b.x, p.x, b.w))),
            RT_FMA(c.z, p.z, RT_FMA(c.y, p.y, RT_FMA(c.x, p.x, c.w))));
}
RT_HD V3f inst_vector_to_object(const float4 *rec, V3f d) {
  const float4 a = rec[0], b = rec[1], c = rec[2];
  return v3(RT_FMA(a.z, d.z, RT_FMA(a.y, d.y, a.x * d.x)), RT_FMA(b.z, d.z, RT_FMA(b.y, d.y, b.x * d.x)),
            RT_FMA(c.z, d.z, RT_FMA(c.y, d.y, c.x * d.x)));
}
// object -> world for a direction / normal: the transpose of the rotation rows
RT_HD V3f inst_vector_to_world(const float4 *rec, V3f n) {
  const float4 a = rec[0], b = rec[1], c = rec[2];
  return v3(RT_FMA(c.x, n.z, RT_FMA(b.x, n.y, a.x * n.x)), RT_FMA(c.y, n.z, RT_FMA(b.y, n.y, a.y * n.x)),
            RT_FMA(c.z, n.z, RT_FMA(b.z, n.y, a.z * n.x)));
}

struct DevCamera {
  V3f origin, llc, horizontal, vertical, u, v;
  float lens_radius, time0, time1;
};

struct Ray {
  V3f o, d;
  float tm;
};
