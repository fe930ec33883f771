// bvh_build.cuh — per-thread bodies of the GPU LBVH builder (one `__host__ __device__`
// function per kernel so that tests/emu can run them serially on the CPU).
//
// Replaces `bvh_node::bvh_node` (rt_next_week/cuda/bvh.h:139-196): one GPU thread,
// recursive `new`, random axis, sequential thrust::sort per level, median split.
// Here: primitive boxes -> oversized-primitive split-off -> 30-bit Morton codes of
// box centroids (unique 64-bit keys: code<<32 | primitive) -> parallel sort ->
// Karras 2012 hierarchy -> bottom-up box fit fused with SAH tree rotations ->
// depth-first threaded layout (escape indices) packed in 32-byte nodes.
#pragma once
#include "rt_common.cuh"

#ifdef __CUDA_ARCH__
#define RT_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#define RT_ATOMIC_MIN(p, v) atomicMin((p), (v))
#define RT_ATOMIC_MAX(p, v) atomicMax((p), (v))
#define RT_FENCE() __threadfence()
#define RT_LDCG4(p) __ldcg(p)
#define RT_LDCG_I(p) __ldcg(p)
#define RT_CLZ64(x) __clzll((long long)(x))
#else
static inline int rt_host_atomic_add(int *p, int v) { int o = *p; *p += v; return o; }
#define RT_ATOMIC_ADD(p, v) rt_host_atomic_add((p), (v))
#define RT_ATOMIC_MIN(p, v) (*(p) = (*(p) < (v)) ? *(p) : (v))
#define RT_ATOMIC_MAX(p, v) (*(p) = (*(p) > (v)) ? *(p) : (v))
#define RT_FENCE()
#define RT_LDCG4(p) (*(p))
#define RT_LDCG_I(p) (*(p))
#define RT_CLZ64(x) ((x) ? __builtin_clzll(x) : 64)
#endif

// order-preserving float <-> int map for atomicMin/Max on floats
RT_HD int f2ord(float f) { int i = RT_F2I(f); return i >= 0 ? i : i ^ 0x7fffffff; }
RT_HD float ord2f(int i) { return RT_I2F(i >= 0 ? i : i ^ 0x7fffffff); }

struct BuildBounds { // ordered-int encoded
  int lo[3], hi[3];
};

struct BuildArrays {
  // primitives of this build unit, unified index gid: spheres, then triangles, then quads, then instances.
  // A unit is the whole scene (single level), the top level of an instanced scene (world primitives + instances,
  // primitives owned by groups are skipped through `exclude`) or one group (its sub-ranges of the three arrays).
  int n_prims, n_spheres, n_tris, n_quads, n_inst, n_boxes;
  const rt_sphere *spheres;   // already offset to the unit's first sphere / triangle / quad ...
  const rt_triangle *tris;
  const rt_quad *quads;
  int id_base[3];             // ... whose indices in the scene arrays these are (RT_PRIM_IDs are scene-wide)
  const float4 *inst_lo, *inst_hi; // [n_inst] world boxes of the instances (host-computed from the groups' root boxes)
  const float4 *box_rec;      // the scene's box records {p0, first rect} {p1, -}
  const int *box_ids;         // [n_boxes] the boxes of this unit (gids after the instances), indices into box_rec
  const uint8_t *exclude;     // [geometry gids] or null: 1 = not a leaf of this unit (owned by a group / a side of a box)
  int link_base;              // node index of this unit's root inside an ordering of the packed array
  int leaf_base;              // index of this unit's first entry in leaf_prims
  int end_link;               // link stored where the traversal leaves this unit's tree (finished / back to the top level)
  float thickness;      // flat-box padding (THICKNESS, rtweekend.h:60 of each tree)
  float4 *pbox_lo, *pbox_hi; // [n_prims]
  int *big_flag;             // [n_prims] 0/1
  BuildBounds *bounds;       // [4] one per classification round
  // tree over the n_small non-big primitives
  int n_small;
  const int *small_gid;      // [n_small] ascending gid
  unsigned long long *keys;  // [n_pad] (morton << 32 | gid), padded with ~0
  int n_pad;
  int *left, *right;         // [n_small-1]
  int *parent;               // [2*n_small-1]
  int *flag;                 // [n_small-1]
  int *size;                 // [2*n_small-1] KEPT nodes in subtree (after leaf collapsing)
  int *lcnt;                 // [2*n_small-1] primitives (build leaves) in subtree
  int max_leaf;              // subtrees with <= max_leaf primitives become ONE leaf node (1..8)
  int32_t *leaf_prims;       // [n_small] RT_PRIM_IDs in depth-first leaf order
  int *swapmask;             // [n_small-1] bit o: in octant o the RIGHT child is visited first
  int packed_stride;         // float4 elements between two octant orderings of the packed nodes
  float4 *nbox_lo, *nbox_hi; // [2*n_small-1]
  float4 *packed;            // [2*(2*n_small-1)] output nodes
};

RT_HD int32_t gid_to_prim_id(const BuildArrays &B, int gid) {
  if (gid < B.n_spheres) return RT_PRIM_ID(RT_PRIM_SPHERE, B.id_base[0] + gid);
  if (gid < B.n_spheres + B.n_tris) return RT_PRIM_ID(RT_PRIM_TRIANGLE, B.id_base[1] + gid - B.n_spheres);
  if (gid < B.n_spheres + B.n_tris + B.n_quads) return RT_PRIM_ID(RT_PRIM_QUAD, B.id_base[2] + gid - B.n_spheres - B.n_tris);
  gid -= B.n_spheres + B.n_tris + B.n_quads;
  if (gid < B.n_inst) return RT_PRIM_ID(RT_PRIM_INSTANCE, gid);
  return RT_PRIM_ID(RT_PRIM_BOX, B.box_ids[gid - B.n_inst]);
}

// ---- kernel 1: primitive boxes (sphere.h:79-84, moving_sphere.h:74-82,
// triangle.h:74-100, aarect.h:26-31) + round-0 scene bounds
RT_HD void body_prim_box(const BuildArrays &B, int gid) {
  float lo[3], hi[3];
  const int n_geom = B.n_spheres + B.n_tris + B.n_quads;
  if (gid < n_geom && B.exclude && B.exclude[gid]) { // owned by a group: neither in this tree nor in its bounds
    B.pbox_lo[gid] = make_float4(0.f, 0.f, 0.f, 0.f);
    B.pbox_hi[gid] = make_float4(0.f, 0.f, 0.f, 0.f);
    B.big_flag[gid] = 2;
    return;
  }
  if (gid >= n_geom + B.n_inst) { // a box: its corners
    const float4 *b = B.box_rec + 2 * B.box_ids[gid - n_geom - B.n_inst];
    lo[0] = b[0].x; lo[1] = b[0].y; lo[2] = b[0].z; hi[0] = b[1].x; hi[1] = b[1].y; hi[2] = b[1].z;
  } else if (gid >= n_geom) {
    const float4 l = B.inst_lo[gid - n_geom], h = B.inst_hi[gid - n_geom];
    lo[0] = l.x; lo[1] = l.y; lo[2] = l.z; hi[0] = h.x; hi[1] = h.y; hi[2] = h.z;
  } else if (gid < B.n_spheres) {
    const rt_sphere &s = B.spheres[gid];
    float r = fabsf(s.radius);
    for (int a = 0; a < 3; a++) {
      float c0 = s.center0[a], c1 = s.moving ? s.center1[a] : c0;
      lo[a] = fminf(c0, c1) - r;
      hi[a] = fmaxf(c0, c1) + r;
    }
  } else if (gid < B.n_spheres + B.n_tris) {
    const rt_triangle &t = B.tris[gid - B.n_spheres];
    for (int a = 0; a < 3; a++) {
      lo[a] = fminf(fminf(t.v0[a], t.v1[a]), t.v2[a]);
      hi[a] = fmaxf(fmaxf(t.v0[a], t.v1[a]), t.v2[a]);
      if (fabsf(lo[a] - hi[a]) < B.thickness) { lo[a] -= B.thickness; hi[a] += B.thickness; }
    }
  } else {
    const rt_quad &q = B.quads[gid - B.n_spheres - B.n_tris];
    int ax = q.axis, ia = (ax == 0) ? 1 : 0, ib = (ax == 2) ? 1 : 2;
    lo[ax] = q.k - B.thickness; hi[ax] = q.k + B.thickness;
    lo[ia] = q.a0; hi[ia] = q.a1;
    lo[ib] = q.b0; hi[ib] = q.b1;
  }
  B.pbox_lo[gid] = make_float4(lo[0], lo[1], lo[2], 0.f);
  B.pbox_hi[gid] = make_float4(hi[0], hi[1], hi[2], 0.f);
  B.big_flag[gid] = 0;
  for (int a = 0; a < 3; a++) {
    RT_ATOMIC_MIN(&B.bounds[0].lo[a], f2ord(lo[a]));
    RT_ATOMIC_MAX(&B.bounds[0].hi[a], f2ord(hi[a]));
  }
}

RT_HD float box_half_area(float ex, float ey, float ez) { return RT_FMA(ex, ey, RT_FMA(ey, ez, ez * ex)); }

// ---- kernel 2 (x rounds): split off primitives whose box covers more than
// `frac` of the current bounds' surface (the r=1000 ground sphere, room walls):
// inside an LBVH they bloat every ancestor (SURVEY.md §7 hard part 5). They go to
// an always-tested list. Bounds of the remainder are accumulated for the next round.
RT_HD void body_classify(const BuildArrays &B, int gid, int round, float frac) {
  if (B.big_flag[gid]) return;
  const BuildBounds &bb = B.bounds[round];
  float sx = ord2f(bb.hi[0]) - ord2f(bb.lo[0]), sy = ord2f(bb.hi[1]) - ord2f(bb.lo[1]),
        sz = ord2f(bb.hi[2]) - ord2f(bb.lo[2]);
  float4 lo = B.pbox_lo[gid], hi = B.pbox_hi[gid];
  float area = box_half_area(hi.x - lo.x, hi.y - lo.y, hi.z - lo.z);
  // (an instance is never split off: the always-tested list holds primitives only)
  if (gid < B.n_spheres + B.n_tris + B.n_quads && area > frac * box_half_area(sx, sy, sz)) { B.big_flag[gid] = 1; return; }
  BuildBounds &nb = B.bounds[round + 1];
  RT_ATOMIC_MIN(&nb.lo[0], f2ord(lo.x)); RT_ATOMIC_MIN(&nb.lo[1], f2ord(lo.y)); RT_ATOMIC_MIN(&nb.lo[2], f2ord(lo.z));
  RT_ATOMIC_MAX(&nb.hi[0], f2ord(hi.x)); RT_ATOMIC_MAX(&nb.hi[1], f2ord(hi.y)); RT_ATOMIC_MAX(&nb.hi[2], f2ord(hi.z));
}

RT_HD unsigned expand_bits10(unsigned v) { // 10 bits -> every third bit
  v = (v * 0x00010001u) & 0xFF0000FFu;
  v = (v * 0x00000101u) & 0x0F00F00Fu;
  v = (v * 0x00000011u) & 0xC30C30C3u;
  v = (v * 0x00000005u) & 0x49249249u;
  return v;
}

// ---- kernel 3: Morton keys of box centroids inside the final bounds
RT_HD void body_morton(const BuildArrays &B, int j, int final_round) {
  if (j >= B.n_small) { B.keys[j] = ~0ull; return; }
  int gid = B.small_gid[j];
  const BuildBounds &bb = B.bounds[final_round];
  float4 lo = B.pbox_lo[gid], hi = B.pbox_hi[gid];
  float c[3] = {0.5f * (lo.x + hi.x), 0.5f * (lo.y + hi.y), 0.5f * (lo.z + hi.z)};
  unsigned q[3];
  for (int a = 0; a < 3; a++) {
    float l = ord2f(bb.lo[a]), h = ord2f(bb.hi[a]);
    float ext = h - l;
    float f = ext > 0.f ? (c[a] - l) / ext : 0.5f;
    f = fminf(fmaxf(f * 1024.0f, 0.0f), 1023.0f);
    q[a] = (unsigned)f;
  }
  unsigned code = (expand_bits10(q[0]) << 2) | (expand_bits10(q[1]) << 1) | expand_bits10(q[2]);
  B.keys[j] = ((unsigned long long)code << 32) | (unsigned)gid;
}

// ---- kernel 4: bitonic compare-exchange step (keys are unique -> no stability needed)
RT_HD void body_bitonic(unsigned long long *a, int i, int j, int k) {
  int ixj = i ^ j;
  if (ixj > i) {
    unsigned long long x = a[i], y = a[ixj];
    bool up = (i & k) == 0;
    if ((x > y) == up) { a[i] = y; a[ixj] = x; }
  }
}

// ---- kernel 5: Karras 2012 — internal node i of the radix tree over sorted unique keys
RT_HD int key_delta(const BuildArrays &B, int i, int j) {
  if (j < 0 || j >= B.n_small) return -1;
  return RT_CLZ64(B.keys[i] ^ B.keys[j]);
}
RT_HD void body_karras(const BuildArrays &B, int i) {
  const int n = B.n_small;
  int d = (key_delta(B, i, i + 1) - key_delta(B, i, i - 1)) >= 0 ? 1 : -1;
  int dmin = key_delta(B, i, i - d);
  int lmax = 2;
  while (key_delta(B, i, i + lmax * d) > dmin) lmax *= 2;
  int l = 0;
  for (int t = lmax / 2; t >= 1; t /= 2)
    if (key_delta(B, i, i + (l + t) * d) > dmin) l += t;
  int j = i + l * d;
  int dnode = key_delta(B, i, j);
  int s = 0;
  for (int t = (l + 1) / 2;; t = (t + 1) / 2) { // ceil halving
    if (key_delta(B, i, i + (s + t) * d) > dnode) s += t;
    if (t == 1) break;
  }
  int gamma = i + s * d + (d < 0 ? -1 : 0);
  int lo = i < j ? i : j, hi = i < j ? j : i;
  int lc = (lo == gamma) ? (n - 1 + gamma) : gamma;           // leaf ids are n-1+k
  int rc = (hi == gamma + 1) ? (n - 1 + gamma + 1) : (gamma + 1);
  B.left[i] = lc; B.right[i] = rc;
  B.parent[lc] = i; B.parent[rc] = i;
  if (i == 0) B.parent[0] = -1;
  B.flag[i] = 0;
}

RT_HD float4 f4min(float4 a, float4 b) { return make_float4(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z), 0.f); }
RT_HD float4 f4max(float4 a, float4 b) { return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), 0.f); }
RT_HD float union_area(float4 alo, float4 ahi, float4 blo, float4 bhi) {
  float4 l = f4min(alo, blo), h = f4max(ahi, bhi);
  return box_half_area(h.x - l.x, h.y - l.y, h.z - l.z);
}

// ---- kernel 6: bottom-up fit. The second thread to reach an internal node owns its
// (finished) subtree: it unions the child boxes, counts subtree nodes and, when
// `rotate` is set, applies the best of the four child<->grandchild swaps if it
// lowers the surface area of the affected internal child (SAH tree rotation).
RT_HD void body_fit(const BuildArrays &B, int leaf, int rotate) {
  const int n = B.n_small;
  int node = n - 1 + leaf;
  int gid = (int)(B.keys[leaf] & 0xffffffffull);
  B.nbox_lo[node] = B.pbox_lo[gid];
  B.nbox_hi[node] = B.pbox_hi[gid];
  B.size[node] = 1;
  B.lcnt[node] = 1;
  if (n == 1) { B.parent[node] = -1; return; }
  int p = B.parent[node];
  while (p >= 0) {
    RT_FENCE();
    if (RT_ATOMIC_ADD(&B.flag[p], 1) == 0) return;
    RT_FENCE();
    int l = RT_LDCG_I(&B.left[p]), r = RT_LDCG_I(&B.right[p]);
    float4 llo = RT_LDCG4(&B.nbox_lo[l]), lhi = RT_LDCG4(&B.nbox_hi[l]);
    float4 rlo = RT_LDCG4(&B.nbox_lo[r]), rhi = RT_LDCG4(&B.nbox_hi[r]);
    if (rotate) {
      float best = 0.f;
      int which = -1;
      int la_ = -1, lb_ = -1, ra_ = -1, rb_ = -1;
      float la = box_half_area(lhi.x - llo.x, lhi.y - llo.y, lhi.z - llo.z);
      float ra = box_half_area(rhi.x - rlo.x, rhi.y - rlo.y, rhi.z - rlo.z);
      if (l < n - 1) { // l internal: swap r with one of l's children
        la_ = RT_LDCG_I(&B.left[l]); lb_ = RT_LDCG_I(&B.right[l]);
        float g0 = la - union_area(rlo, rhi, RT_LDCG4(&B.nbox_lo[lb_]), RT_LDCG4(&B.nbox_hi[lb_])); // r <-> la_
        float g1 = la - union_area(RT_LDCG4(&B.nbox_lo[la_]), RT_LDCG4(&B.nbox_hi[la_]), rlo, rhi); // r <-> lb_
        if (g0 > best) { best = g0; which = 0; }
        if (g1 > best) { best = g1; which = 1; }
      }
      if (r < n - 1) {
        ra_ = RT_LDCG_I(&B.left[r]); rb_ = RT_LDCG_I(&B.right[r]);
        float g2 = ra - union_area(llo, lhi, RT_LDCG4(&B.nbox_lo[rb_]), RT_LDCG4(&B.nbox_hi[rb_])); // l <-> ra_
        float g3 = ra - union_area(RT_LDCG4(&B.nbox_lo[ra_]), RT_LDCG4(&B.nbox_hi[ra_]), llo, lhi); // l <-> rb_
        if (g2 > best) { best = g2; which = 2; }
        if (g3 > best) { best = g3; which = 3; }
      }
      if (which >= 0) {
        int c = (which < 2) ? l : r;      // internal child that is rebuilt
        int o = (which < 2) ? r : l;      // the other child, pushed down
        int ca = (which < 2) ? la_ : ra_, cb = (which < 2) ? lb_ : rb_;
        int up = (which & 1) ? cb : ca;   // grandchild pulled up
        int keep = (which & 1) ? ca : cb;
        B.left[c] = keep; B.right[c] = o;
        B.parent[o] = c; B.parent[up] = p;
        if (which < 2) B.right[p] = up; else B.left[p] = up;
        float4 klo = RT_LDCG4(&B.nbox_lo[keep]), khi = RT_LDCG4(&B.nbox_hi[keep]);
        float4 olo = (which < 2) ? rlo : llo, ohi = (which < 2) ? rhi : lhi;
        B.nbox_lo[c] = f4min(klo, olo);
        B.nbox_hi[c] = f4max(khi, ohi);
        const int lc_c = RT_LDCG_I(&B.lcnt[keep]) + RT_LDCG_I(&B.lcnt[o]);
        B.lcnt[c] = lc_c;
        B.size[c] = lc_c <= B.max_leaf ? 1 : 1 + RT_LDCG_I(&B.size[keep]) + RT_LDCG_I(&B.size[o]);
        RT_FENCE();
        if (which < 2) { r = up; rlo = RT_LDCG4(&B.nbox_lo[r]); rhi = RT_LDCG4(&B.nbox_hi[r]);
                         llo = f4min(klo, olo); lhi = f4max(khi, ohi); }
        else { l = up; llo = RT_LDCG4(&B.nbox_lo[l]); lhi = RT_LDCG4(&B.nbox_hi[l]);
               rlo = f4min(klo, olo); rhi = f4max(khi, ohi); }
      }
    }
    B.nbox_lo[p] = f4min(llo, rlo);
    B.nbox_hi[p] = f4max(lhi, rhi);
    const int lc_p = RT_LDCG_I(&B.lcnt[l]) + RT_LDCG_I(&B.lcnt[r]);
    B.lcnt[p] = lc_p;
    B.size[p] = lc_p <= B.max_leaf ? 1 : 1 + RT_LDCG_I(&B.size[l]) + RT_LDCG_I(&B.size[r]);
    p = RT_LDCG_I(&B.parent[p]);
  }
}

// ---- kernel 6b: front-to-back child order per ray-direction OCTANT. Octant o = sign bits of
// (d.x, d.y, d.z) (bit 0: d.x < 0, bit 1: d.y < 0, bit 2: d.z < 0). Children are ordered along the
// axis on which their box centres differ most: the child the ray meets first comes first, so
// hits shrink t_max early and later subtrees are culled by the slab test.
#define RT_N_ORDERINGS 8
RT_HD void body_order(const BuildArrays &B, int p) {
  const int l = B.left[p], r = B.right[p];
  const float4 llo = B.nbox_lo[l], lhi = B.nbox_hi[l], rlo = B.nbox_lo[r], rhi = B.nbox_hi[r];
  const float d[3] = {(rlo.x + rhi.x) - (llo.x + lhi.x), (rlo.y + rhi.y) - (llo.y + lhi.y), (rlo.z + rhi.z) - (llo.z + lhi.z)};
  int ax = 0;
  if (fabsf(d[1]) > fabsf(d[ax])) ax = 1;
  if (fabsf(d[2]) > fabsf(d[ax])) ax = 2;
  const bool left_low = d[ax] >= 0.f; // left is the low child on that axis
  int mask = 0;
  for (int o = 0; o < RT_N_ORDERINGS; o++) { const bool neg = (o >> ax) & 1; if (left_low == neg) mask |= 1 << o; }
  B.swapmask[p] = mask;
}

// ---- kernel 7: depth-first threaded layout + 32-byte packing. A build node `v` is
// emitted unless it lies strictly inside a collapsed subtree (an ancestor holds
// <= max_leaf primitives and became a multi-primitive leaf). Its position = number of
// emitted nodes before it in DFS preorder (walk to the root); escape = position +
// kept subtree size. Leaves reference leaf_prims[first .. first+count): payload =
// ~(first << 3 | count-1). Boxes are padded by a few ulps so that the FP32 slab test
// stays conservative.
RT_HD void body_pack(const BuildArrays &B, int v, int quadrant) {
  const int n = B.n_small;
  int pos = 0, lpos = 0, cur = v;
  for (int p = B.parent[cur]; p >= 0; p = B.parent[cur]) {
    pos += 1;
    const bool is_right = B.right[p] == cur;
    const bool right_first = (B.swapmask[p] >> quadrant) & 1;
    if (is_right != right_first) pos += B.size[is_right ? B.left[p] : B.right[p]]; // visited second in this quadrant
    if (is_right) lpos += B.lcnt[B.left[p]];                                       // leaf order stays canonical
    cur = p;
  }
  const bool build_leaf = v >= n - 1;
  if (build_leaf && quadrant == 0) B.leaf_prims[B.leaf_base + lpos] = gid_to_prim_id(B, (int)(B.keys[v - (n - 1)] & 0xffffffffull));
  const int par = B.parent[v];
  if (par >= 0 && B.lcnt[par] <= B.max_leaf) return; // inside a collapsed subtree
  float4 lo = B.nbox_lo[v], hi = B.nbox_hi[v];
  float e[3] = {fmaxf(fabsf(lo.x), fabsf(hi.x)), fmaxf(fabsf(lo.y), fabsf(hi.y)), fmaxf(fabsf(lo.z), fabsf(hi.z))};
  lo.x -= RT_FMA(e[0], 4e-7f, 1e-9f); lo.y -= RT_FMA(e[1], 4e-7f, 1e-9f); lo.z -= RT_FMA(e[2], 4e-7f, 1e-9f);
  hi.x += RT_FMA(e[0], 4e-7f, 1e-9f); hi.y += RT_FMA(e[1], 4e-7f, 1e-9f); hi.z += RT_FMA(e[2], 4e-7f, 1e-9f);
  const int escape = pos + B.size[v];
  int payload;
  if (build_leaf || B.lcnt[v] <= B.max_leaf) payload = ~(((B.leaf_base + lpos) << 3) | (B.lcnt[v] - 1));
  else payload = B.link_base + pos + 1;
  // links are stored as BYTE offsets (index * 32) so that traversal needs no address
  // arithmetic; rt_accel_download converts them back to indices. The link that leaves this unit's tree
  // (escape == its kept node count) is the unit's end link.
  // (end_link < 0: a single-level scene, where "finished" is simply the node count)
  lo.w = RT_I2F((escape == B.size[0] && B.end_link >= 0) ? B.end_link : ((B.link_base + escape) << RT_NODE_SHIFT));
  hi.w = RT_I2F(payload >= 0 ? (payload << RT_NODE_SHIFT) : payload);
  float4 *dst = B.packed + (size_t)quadrant * B.packed_stride + 2 * (size_t)B.link_base;
  dst[2 * pos] = lo;
  dst[2 * pos + 1] = hi;
}
