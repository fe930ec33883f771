// intersect.cuh — ray/primitive and ray/box tests + the threaded-BVH closest-hit
// query. Replaces the reference's virtual `hittable::hit` family:
//   sphere::hit         rt_in_one_weekend/sphere.h:21-55, accelerated-rt-cuda/sphere.h:20-49,
//                       rt_next_week/cuda/sphere.h:43-77
//   moving_sphere::hit  rt_next_week/cuda/moving_sphere.h:39-72
//   triangle::hit       triangles/cuda/include/triangle.h:102-215
//   xy/xz/yz_rect::hit  rt_next_week/cuda/aarect.h:38-65,93-121,149-176
//   aabb::hit           rt_next_week/cuda/aabb.h:33-50
//   hittable_list::hit  hittable_list.h:20-34     bvh_node::hit  rt_next_week/cuda/bvh.h:78-137
// The arithmetic is re-derived for accuracy in FP32 (see DESIGN.md §"closest hit"):
// results must agree with the reference's DOUBLE renderer to 1e-5 relative in t,
// which the reference's own float formulas do not achieve (up to 3.6e-4 measured).
#pragma once
#include "rt_common.cuh"

struct RayPre {
  V3f inv_d;  // 1/d (inf where d == 0)
  V3f ood;    // o/d
  float inv_a; // 1/(d.d)
};

// A direction component of (+-)0 would make  b/d - o/d  = inf - inf = NaN for planes on the
// same side of the origin, and the NaN-dropping min/max would then lose the slab's far
// plane: nudge such components to +-1e-20 (1/d = +-1e20 stays finite and keeps the sign of b - o).
RT_HD float rt_safe_dir(float d) { return fabsf(d) < 1e-20f ? copysignf(1e-20f, d) : d; }

RT_HD RayPre ray_precompute(const Ray &r) {
  RayPre p;
  p.inv_d = v3(1.0f / rt_safe_dir(r.d.x), 1.0f / rt_safe_dir(r.d.y), 1.0f / rt_safe_dir(r.d.z));
  p.ood = v3(r.o.x * p.inv_d.x, r.o.y * p.inv_d.y, r.o.z * p.inv_d.z);
  p.inv_a = 1.0f / dot(r.d, r.d);
  return p;
}

// Same with hardware reciprocals (MUFU.RCP, <= 2 ulp) for the render kernel: the slab test is
// conservative by construction (padded boxes) and a 2-ulp error in 1/(d.d) moves t by 2e-7
// relative, well inside the 1e-5 parity bound; the parity hook keeps the IEEE version.
RT_HD RayPre ray_precompute_fast(const Ray &r) {
#ifdef __CUDA_ARCH__
  RayPre p;
  p.inv_d = v3(rt_rcp_approx(rt_safe_dir(r.d.x)), rt_rcp_approx(rt_safe_dir(r.d.y)), rt_rcp_approx(rt_safe_dir(r.d.z)));
  p.ood = v3(r.o.x * p.inv_d.x, r.o.y * p.inv_d.y, r.o.z * p.inv_d.z);
  p.inv_a = rt_rcp_approx(dot(r.d, r.d));
  return p;
#else
  return ray_precompute(r);
#endif
}

struct HitAcc {
  float t;    // closest accepted t so far (starts at t_max)
  int32_t id; // RT_PRIM_ID or -1
};

// Interval + tie rule. closed: the reference accepts roots on [t_min, t_max] and the
// LATER list item wins an exact tie (sphere.h:37-43 with hittable_list.h:26-31);
// open: (t_min, t_max), the FIRST wins (accelerated-rt-cuda/sphere.h:29). List
// order == primitive-id order, which makes the result independent of BVH order.
RT_HD bool accept_root(float root, float t_min, bool closed, const HitAcc &h, int32_t id) {
  bool lo = closed ? (root >= t_min) : (root > t_min);
  bool better = (root < h.t) || (root == h.t && (closed ? (id > h.id) : (id < h.id)));
  return lo && better;
}

// Sphere, FP32, cancellation-free: with g = c - o, q = (g.d)/(d.d) the parameter of
// closest approach and l = g - q d the perpendicular from the ray to the centre,
//   disc = r^2 - |l|^2,   roots = q -/+ sqrt(disc/(d.d)).
// (|l| <= r whenever there is a hit, so no large numbers are subtracted.)
// Shared root-selection logic, written without branches (predicate arithmetic only): in the
// render kernel a handful of lanes run this while the rest of the warp waits, so every
// divergent side costs a full warp instruction (ncu r1h: the branchy version averaged 5-9
// active lanes). r1 <= r2 are the two roots; `valid` = the discriminant test passed.
RT_HD void take_sphere_roots(bool valid, float r1, float r2, bool closed, float t_min, HitAcc &h, int32_t id) {
  const bool tie_wins = closed ? (id > h.id) : (id < h.id);
  const bool lo1 = closed ? (r1 >= t_min) : (r1 > t_min);
  const bool lo2 = closed ? (r2 >= t_min) : (r2 > t_min);
  const bool ok1 = lo1 & ((r1 < h.t) | ((r1 == h.t) & tie_wins));
  // the near root was inside the interval (so the reference stops there) even if it lost a tie
  const bool in1 = lo1 & (closed ? (r1 <= h.t) : (r1 < h.t));
  const bool ok2 = (!in1) & lo2 & ((r2 < h.t) | ((r2 == h.t) & tie_wins));
  const bool take = valid & (ok1 | ok2);
  const float t = ok1 ? r1 : r2;
  h.t = take ? t : h.t;
  h.id = take ? id : h.id;
}

RT_HD void hit_sphere(float4 s, V3f center, bool closed, const Ray &r, const RayPre &pre, float t_min, HitAcc &h,
                      int32_t id) {
  V3f g = center - r.o;
  float q = dot(g, r.d) * pre.inv_a;
  V3f l = madd(g, -q, r.d);
  float disc = RT_FMA(s.w, s.w, -dot(l, l));
  const bool valid = closed ? (disc >= 0.0f) : (disc > 0.0f);
  float sq = RT_SQRT(RT_FMAX(disc, 0.0f) * pre.inv_a);
  take_sphere_roots(valid, q - sq, q + sq, closed, t_min, h, id);
}

// Oversized spheres (the r = 1000 ground, always-tested "big" list). With |c - o| ~ 1e3
// the subtraction c - o alone loses 3e-5 of the height above the surface in FP32, so
// the quadratic is expanded around the ORIGIN instead:  f = |o-c|^2 - r^2 =
// |o|^2 - 2 o.c + K  with K = |c|^2 - r^2 precomputed in double on the host (0 for the
// reference's ground sphere): no large terms cancel for origins near the surface.
// Roots from the numerically stable pair  s = bp + sign(bp) sqrt(bp^2 - a f):
// t_a = f / s, t_b = s / a.   (bp = (c-o).d, a = d.d)
RT_HD void hit_sphere_big(float4 s, float K, bool closed, const Ray &r, const RayPre &pre, float t_min, HitAcc &h,
                          int32_t id) {
  const V3f c = xyz(s);
  const float f = RT_FMA(-2.0f, dot(r.o, c), dot(r.o, r.o)) + K;
  const float a = dot(r.d, r.d);
  const float bp = dot(c, r.d) - dot(r.o, r.d);
  const float disc = RT_FMA(bp, bp, -a * f);
  const bool valid = closed ? (disc >= 0.0f) : (disc > 0.0f);
  const float sq = RT_SQRT(RT_FMAX(disc, 0.0f));
  const float s1 = bp + (bp >= 0.0f ? sq : -sq);
  const float ta = (s1 != 0.0f) ? RT_FDIV(f, s1) : 0.0f;
  const float tb = s1 * pre.inv_a;
  take_sphere_roots(valid, RT_FMIN(ta, tb), RT_FMAX(ta, tb), closed, t_min, h, id);
}

// moving_sphere::center — moving_sphere.h:34-36. mv = {c1-c0, 1/(time1-time0)}.
RT_HD V3f sphere_center_at(float4 s, float4 mv, float time0, float tm) {
  float f = (tm - time0) * mv.w;
  return v3(RT_FMA(f, mv.x, s.x), RT_FMA(f, mv.y, s.y), RT_FMA(f, mv.z, s.z));
}

// Triangle in plane form. The reference tests  |N.d| < 0.01 (un-normalised N and d,
// triangle.h:123-128), t = (v0-o).N / N.d, t >= 0, t in [t_min,t_max], then three
// edge tests dot(N^, cross(edge, p - vert)) >= 0 (:172-202). With
// m_e = cross(N^, edge), k_e = m_e.vert the edge test is m_e.p - k_e >= 0.
RT_HD void hit_triangle(float4 pl, float4 e0, float4 e1, float4 e2, const Ray &r, float t_min, HitAcc &h,
                        int32_t id) {
  // branch-free: a rejected candidate leaves NaN/inf in t or p, and every comparison below is
  // written so that NaN fails it
  const V3f n = xyz(pl);
  const float nd = dot(n, r.d);
  const float t = RT_FDIV(pl.w - dot(n, r.o), nd);
  bool ok = (fabsf(nd) >= 0.01f) & (t >= 0.0f) & accept_root(t, t_min, true, h, id);
  const V3f p = madd(r.o, t, r.d);
  ok = ok & (dot(xyz(e0), p) - e0.w >= 0.0f) & (dot(xyz(e1), p) - e1.w >= 0.0f) & (dot(xyz(e2), p) - e2.w >= 0.0f);
  h.t = ok ? t : h.t;
  h.id = ok ? id : h.id;
}

// Axis-aligned rectangle: q0 = {k, a0, a1, as_float(axis)}, q1 = {b0, b1, -, -}.
// t = (k - o_axis) / d_axis with the ray's precomputed reciprocal (the subtraction comes first: a
// ray leaving this very plane gets t = 0 exactly). The axis is switched on, not selected by
// arithmetic: for the always-tested walls it is uniform across the warp.
RT_HD void hit_quad(float4 q0, float4 q1, const Ray &r, const RayPre &pre, float t_min, HitAcc &h, int32_t id) {
  const int ax = RT_F2I(q0.w);
  float t, a, b; // in-plane coordinates: axis 0 -> (y,z), 1 -> (x,z), 2 -> (x,y)
  if (ax == 0) {
    t = (q0.x - r.o.x) * pre.inv_d.x;
    a = RT_FMA(t, r.d.y, r.o.y); b = RT_FMA(t, r.d.z, r.o.z);
  } else if (ax == 1) {
    t = (q0.x - r.o.y) * pre.inv_d.y;
    a = RT_FMA(t, r.d.x, r.o.x); b = RT_FMA(t, r.d.z, r.o.z);
  } else {
    t = (q0.x - r.o.z) * pre.inv_d.z;
    a = RT_FMA(t, r.d.x, r.o.x); b = RT_FMA(t, r.d.y, r.o.y);
  }
  const bool inside = (a >= q0.y) & (a <= q0.z) & (b >= q1.x) & (b <= q1.y);
  const bool take = inside & accept_root(t, t_min, true, h, id);
  h.t = take ? t : h.t;
  h.id = take ? id : h.id;
}

// box(p0, p1) = six rects in box.h's order (rt_next_week/cuda/box.h:41-58: xy@z1, xy@z0, xz@y1, xz@y0, yz@x1,
// yz@x0), rects first .. first + 5 of the scene. The six rect tests run on the same numbers the six rects hold
// (scene_flatten.h recognises a box by exact equality), in rect order, with the rects' ids: the result is
// bit-identical to six separate leaves, at one leaf visit instead of up to six.
// (Inline. Out-of-line copies of primitive tests were measured and rejected: with a second kind of call next to
// the out-of-line Philox in the kernel, the rt_next_week final scene rendered 30 % slower.)
RT_HD HitAcc hit_box_sides(float4 b0, float4 b1, const Ray &r, const RayPre &pre, float t_min, HitAcc h) {
  const int first = RT_F2I(b0.w);
  const float lo[3] = {b0.x, b0.y, b0.z}, hi[3] = {b1.x, b1.y, b1.z};
  const float o[3] = {r.o.x, r.o.y, r.o.z}, d[3] = {r.d.x, r.d.y, r.d.z}, inv[3] = {pre.inv_d.x, pre.inv_d.y, pre.inv_d.z};
#ifdef RT_BOX_SEQUENTIAL // the six rect tests one after the other against the running hit (the form this replaces; A/B builds)
#pragma unroll
  for (int j = 0; j < 6; j++) {
    const int ax = 2 - (j >> 1), ia = ax == 0 ? 1 : 0, ib = ax == 2 ? 1 : 2;
    const float4 q0 = make_float4((j & 1) ? lo[ax] : hi[ax], lo[ia], hi[ia], RT_I2F(ax));
    const float4 q1 = make_float4(lo[ib], hi[ib], 0.f, 0.f);
    hit_quad(q0, q1, r, pre, t_min, h, RT_PRIM_ID(RT_PRIM_QUAD, first + j));
  }
  return h;
#else
  // The same six rect formulas in the same order, but the sides first compete among themselves - the smallest
  // t >= t_min inside its rect wins, the LATER side on an exact tie (ids grow with j and the closed rule lets the later
  // list item win) - and only the winner meets the running hit, under the list rule. Sequentially updating h through
  // the six sides gives the same (id, t): the minimum over {h, valid sides} with the same tie-breaks. Per side this
  // saves the comparison against h.t / h.id and the id arithmetic; few lanes run this code (3-4 of 32 in the
  // rt_next_week final scene), so its length is what counts.
  float bt = INFINITY;
  int bj = -1;
#pragma unroll
  for (int j = 0; j < 6; j++) {
    const int ax = 2 - (j >> 1), ia = ax == 0 ? 1 : 0, ib = ax == 2 ? 1 : 2;
    const float k = (j & 1) ? lo[ax] : hi[ax];
    const float t = (k - o[ax]) * inv[ax]; // hit_quad's own expressions: t, then the in-plane point by FMA
    const float a = RT_FMA(t, d[ia], o[ia]), b = RT_FMA(t, d[ib], o[ib]);
    const bool ok = (a >= lo[ia]) & (a <= hi[ia]) & (b >= lo[ib]) & (b <= hi[ib]) & (t >= t_min) & (t <= bt);
    bt = ok ? t : bt;
    bj = ok ? j : bj;
  }
  const int32_t id = RT_PRIM_ID(RT_PRIM_QUAD, first + bj);
  const bool take = (bj >= 0) & ((bt < h.t) | ((bt == h.t) & (id > h.id)));
  h.t = take ? bt : h.t;
  h.id = take ? id : h.id;
  return h;
#endif
}

// Slab test against a packed node box (closed interval). 1/d is finite (rt_safe_dir), so the
// only NaN left is 0 * huge - 0 * huge = 0 handled exactly; fminf/fmaxf would drop any other.
RT_HD bool hit_box(float4 lo, float4 hi, const RayPre &pre, float t_min, float t_max) {
  float x0 = RT_FMA(lo.x, pre.inv_d.x, -pre.ood.x), x1 = RT_FMA(hi.x, pre.inv_d.x, -pre.ood.x);
  float y0 = RT_FMA(lo.y, pre.inv_d.y, -pre.ood.y), y1 = RT_FMA(hi.y, pre.inv_d.y, -pre.ood.y);
  float z0 = RT_FMA(lo.z, pre.inv_d.z, -pre.ood.z), z1 = RT_FMA(hi.z, pre.inv_d.z, -pre.ood.z);
  float tn = RT_FMAX(RT_FMAX(RT_FMIN(x0, x1), RT_FMIN(y0, y1)), RT_FMAX(RT_FMIN(z0, z1), t_min));
  float tf = RT_FMIN(RT_FMIN(RT_FMAX(x0, x1), RT_FMAX(y0, y1)), RT_FMIN(RT_FMAX(z0, z1), t_max));
  return tn <= tf;
}

// Same test for a node copy whose x and z planes were pre-swapped for the ray's direction
// signs (lo = entry plane, hi = exit plane on those axes): 6 instead of 10 min/max.
RT_HD bool hit_box_xz_sorted(float4 lo, float4 hi, const RayPre &pre, float t_min, float t_max) {
  float x0 = RT_FMA(lo.x, pre.inv_d.x, -pre.ood.x), x1 = RT_FMA(hi.x, pre.inv_d.x, -pre.ood.x);
  float y0 = RT_FMA(lo.y, pre.inv_d.y, -pre.ood.y), y1 = RT_FMA(hi.y, pre.inv_d.y, -pre.ood.y);
  float z0 = RT_FMA(lo.z, pre.inv_d.z, -pre.ood.z), z1 = RT_FMA(hi.z, pre.inv_d.z, -pre.ood.z);
  float tn = RT_FMAX(RT_FMAX(x0, z0), RT_FMAX(RT_FMIN(y0, y1), t_min));
  float tf = RT_FMIN(RT_FMIN(x1, z1), RT_FMIN(RT_FMAX(y0, y1), t_max));
  return tn <= tf;
}

// One primitive by RT_PRIM_ID. PROFILE selects the interval rule of static spheres.
// SPH: a profile-2 scene that holds spheres only (config 4: moving spheres, checker ground) - no type dispatch.
template <int PROFILE, bool GENERAL, bool BIG, bool SPH = false>
RT_HD void hit_prim(const DevScene &S, int32_t id, const Ray &r, const RayPre &pre, float t_min, HitAcc &h) {
  if (!GENERAL) {
    float4 s = S.sph[id];
    // K = NaN: the flattening step found the ordinary form more accurate for this sphere (scene_flatten.h)
    const float K = BIG ? S.sph_k[id] : 0.f;
    if (BIG && K == K) hit_sphere_big(s, K, PROFILE == 0, r, pre, t_min, h, id);
    else hit_sphere(s, xyz(s), PROFILE == 0, r, pre, t_min, h, id);
    return;
  }
  const int type = SPH ? (int)RT_PRIM_SPHERE : RT_PRIM_TYPE_OF(id), idx = SPH ? id : RT_PRIM_INDEX_OF(id);
  if (type == RT_PRIM_SPHERE) {
    float4 s = S.sph[idx];
    V3f c = xyz(s);
    bool closed = PROFILE == 0;
    bool moving = false;
    if (S.any_moving) {
      float4 mv = S.sph_mv[idx];
      if (mv.w != 0.0f) { c = sphere_center_at(s, mv, S.sph_t0[idx], r.tm); closed = true; moving = true; }
    }
    const float K = (BIG && !moving) ? S.sph_k[idx] : 0.f;
    if (BIG && !moving && K == K) hit_sphere_big(s, K, closed, r, pre, t_min, h, id);
    else hit_sphere(s, c, closed, r, pre, t_min, h, id);
  } else if (type == RT_PRIM_TRIANGLE) {
    const float4 *t = S.tri + 4 * idx;
    hit_triangle(t[0], t[1], t[2], t[3], r, t_min, h, id);
  } else if (type == RT_PRIM_QUAD) {
    const float4 *q = S.quad + 2 * idx;
    hit_quad(q[0], q[1], r, pre, t_min, h, id);
  } else { // RT_PRIM_BOX: six rects of a box in one leaf
    const float4 *b = S.box + 2 * idx;
    h = hit_box_sides(b[0], b[1], r, pre, t_min, h);
  }
}

struct TraceCounters {
  unsigned box_tests, prim_tests;
};

// Closest hit: always-tested big primitives first (they shrink t_max early), then
// the threaded BVH — no stack: every node carries the index to continue with when
// its box is missed (or its subtree is done).  while-while form: lanes first walk
// to their next leaf, then the warp tests primitives together.
// INST (two-level scenes): a top-level leaf may name an rt_instance; the ray is moved into the object's space
// (rigid: t is preserved, rt_next_week/cuda/hittable.h:66-79,156-190), the object's tree is walked until its end
// link RT_POP_LINK, and the top-level walk resumes with the world ray. *hit_inst = instance of the closest hit (-1: world).
template <int PROFILE, bool GENERAL, bool COUNT, bool INST = false>
RT_HD HitAcc trace_closest(const DevScene &S, const Ray &r_world, float t_min, float t_max, TraceCounters *cnt,
                           int *hit_inst = nullptr) {
  HitAcc h;
  h.t = t_max;
  h.id = -1;
  Ray r = r_world;
  RayPre pre = ray_precompute(r);
  for (int i = 0; i < S.n_big; i++) {
    int32_t id = S.big[i];
    if (COUNT) cnt->prim_tests++;
    bool is_sphere = !GENERAL || RT_PRIM_TYPE_OF(id) == RT_PRIM_SPHERE;
    if (is_sphere) hit_prim<PROFILE, GENERAL, true>(S, id, r, pre, t_min, h);
    else hit_prim<PROFILE, GENERAL, false>(S, id, r, pre, t_min, h);
  }
  if (GENERAL) {
    for (int i = 0; i < S.n_bigq; i++) {
      if (COUNT) cnt->prim_tests++;
      const float4 q0 = S.bigq[2 * i], q1 = S.bigq[2 * i + 1];
      hit_quad(q0, q1, r, pre, t_min, h, RT_F2I(q1.z));
    }
  }
  int node = 0; // byte offset of the current node
  const int n_nodes = S.n_nodes << RT_NODE_SHIFT;
  const int pop = RT_POP_LINK(S.n_nodes);
  int first = 0, left = 0; // pending primitives of the last hit leaf: leaf_prims[first .. first+left)
  int cur_inst = -1, best_inst = -1, node_top = 0, first_top = 0, left_top = 0;
  for (;;) {
    if (left == 0) {
      if (INST && node == pop) { // the object's tree is done: back to the top level with the world ray
        r = r_world;
        pre = ray_precompute(r);
        node = node_top; first = first_top; left = left_top;
        cur_inst = -1;
        continue;
      }
      if (!((unsigned)node < (unsigned)n_nodes)) break;
      while ((unsigned)node < (unsigned)n_nodes) {
        const float4 *np = (const float4 *)((const char *)S.nodes + node);
        float4 lo = np[0], hi = np[1];
        if (COUNT) cnt->box_tests++;
        int esc = RT_F2I(lo.w), pay = RT_F2I(hi.w);
        if (hit_box(lo, hi, pre, t_min, h.t)) {
          if (pay >= 0) { node = pay; continue; }
          first = (~pay) >> 3;
          left = ((~pay) & 7) + 1;
          node = esc;
          break;
        }
        node = esc;
      }
    }
    if (left > 0) { // one primitive per outer iteration keeps the warp converged here
      const int32_t id = S.leaf_prims[first];
      first++;
      left--;
      if (INST && RT_PRIM_TYPE_OF(id) == RT_PRIM_INSTANCE) {
        const float4 *rec = S.inst + RT_INST_STRIDE * RT_PRIM_INDEX_OF(id);
        node_top = node; first_top = first; left_top = left;
        r.o = inst_point_to_object(rec, r_world.o);
        r.d = inst_vector_to_object(rec, r_world.d);
        pre = ray_precompute(r);
        node = RT_F2I(rec[3].x);
        first = 0; left = 0;
        cur_inst = RT_PRIM_INDEX_OF(id);
      } else {
        if (COUNT) cnt->prim_tests++;
        const HitAcc old = h;
        hit_prim<PROFILE, GENERAL, false>(S, id, r, pre, t_min, h);
        if (INST && (h.id != old.id || h.t != old.t)) best_inst = cur_inst;
      }
    }
  }
  if (INST && hit_inst) *hit_inst = best_inst;
  return h;
}

// Brute force over the flattened arrays in list order (parity hook; the reference's
// hittable_list::hit). Spheres in the big list still use the big-sphere formula so that the
// result is identical to the BVH path's. Two-level scenes: the world primitives (those no group owns), then
// every instance's object with the ray moved into its space.
template <int PROFILE, bool GENERAL, bool INST = false>
RT_HD HitAcc trace_brute(const DevScene &S, const uint8_t *is_big, const Ray &r, float t_min, float t_max,
                         const uint8_t *grouped = nullptr, int *hit_inst = nullptr) {
  HitAcc h;
  h.t = t_max;
  h.id = -1;
  RayPre pre = ray_precompute(r);
  for (int i = 0; i < S.n_spheres; i++) {
    if (INST && grouped[i]) continue;
    int32_t id = RT_PRIM_ID(RT_PRIM_SPHERE, i);
    if (is_big[i]) hit_prim<PROFILE, GENERAL, true>(S, id, r, pre, t_min, h);
    else hit_prim<PROFILE, GENERAL, false>(S, id, r, pre, t_min, h);
  }
  if (GENERAL) {
    for (int i = 0; i < S.n_tris; i++)
      if (!(INST && grouped[S.n_spheres + i])) hit_prim<PROFILE, true, false>(S, RT_PRIM_ID(RT_PRIM_TRIANGLE, i), r, pre, t_min, h);
    for (int i = 0; i < S.n_quads; i++)
      if (!(INST && grouped[S.n_spheres + S.n_tris + i])) hit_prim<PROFILE, true, false>(S, RT_PRIM_ID(RT_PRIM_QUAD, i), r, pre, t_min, h);
  }
  if (INST) {
    int best_inst = -1;
    for (int k = 0; k < S.n_inst; k++) {
      const float4 *rec = S.inst + RT_INST_STRIDE * k;
      const int32_t *G = S.groups + 8 * RT_F2I(rec[3].y); // rt_group: first/count of spheres, triangles, quads
      Ray ro = r;
      ro.o = inst_point_to_object(rec, r.o);
      ro.d = inst_vector_to_object(rec, r.d);
      const RayPre po = ray_precompute(ro);
      const HitAcc old = h;
      for (int i = 0; i < G[1]; i++) hit_prim<PROFILE, true, false>(S, RT_PRIM_ID(RT_PRIM_SPHERE, G[0] + i), ro, po, t_min, h);
      for (int i = 0; i < G[3]; i++) hit_prim<PROFILE, true, false>(S, RT_PRIM_ID(RT_PRIM_TRIANGLE, G[2] + i), ro, po, t_min, h);
      for (int i = 0; i < G[5]; i++) hit_prim<PROFILE, true, false>(S, RT_PRIM_ID(RT_PRIM_QUAD, G[4] + i), ro, po, t_min, h);
      if (h.id != old.id || h.t != old.t) best_inst = k;
    }
    if (hit_inst) *hit_inst = best_inst;
  }
  return h;
}
