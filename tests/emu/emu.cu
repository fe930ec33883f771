// tests/emu/emu.cu — TEST-ONLY host emulation of the device code.
//
// Compiles the very same `__host__ __device__` per-thread bodies the kernels run
// (csrc/bvh_build.cuh, intersect.cuh, shade.cuh, philox.cuh) for the CPU and steps
// them serially, so that their logic can be checked against the oracle in a container
// without a GPU before GPU time is spent. It is NOT part of the product: libb200rt.so
// never contains or calls this, and no result reported anywhere comes from it.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../a_dive_into_ray_tracing_b200/csrc/bvh_build.cuh"
#include "../../a_dive_into_ray_tracing_b200/csrc/scene_flatten.h"
#include "../../a_dive_into_ray_tracing_b200/csrc/render_kernels.cuh"

struct EmuScene {
  HostFlat F;
  std::vector<float4> nodes;
  std::vector<int32_t> big, leaf_prims;
  std::vector<uint8_t> sph_is_big;
  std::vector<DevImage> images;
  DevScene S;
  DevCamera cam;
  ShadeParams sp;
  int profile;
  // two-level scenes
  bool two_level = false;
  int n_nodes_total = 0;
  std::vector<float4> inst_lo, inst_hi;
};

static void bind(EmuScene &E, const rt_scene_desc *sc) {
  DevScene &S = E.S;
  memset(&S, 0, sizeof S);
  S.nodes = E.nodes.data(); S.sph = E.F.sph.data(); S.sph_mv = E.F.sph_mv.data(); S.sph_t0 = E.F.sph_t0.data(); S.sph_k = E.F.sph_k.data();
  S.tri = E.F.tri.data(); S.tri_n = E.F.tri_n.data(); S.quad = E.F.quad.data();
  S.sph_mat = E.F.sph_mat.data(); S.tri_mat = E.F.tri_mat.data(); S.quad_mat = E.F.quad_mat.data();
  S.mats = E.F.mats.data(); S.big = E.big.data(); S.leaf_prims = E.leaf_prims.data();
  S.n_nodes = E.n_nodes_total; S.n_big = (int)E.big.size();
  S.inst = E.F.inst.data(); S.n_inst = sc->n_instances; S.groups = (const int32_t *)sc->groups;
  S.box = E.F.box.data(); S.n_boxes = (int)E.F.box.size() / 2;
  S.n_spheres = sc->n_spheres; S.n_tris = sc->n_triangles; S.n_quads = sc->n_quads; S.n_mats = sc->n_materials;
  S.any_moving = E.F.any_moving;
  E.images.resize((size_t)sc->n_images);
  for (int i = 0; i < sc->n_images; i++) {
    E.images[i].rgb = E.F.image_bytes.data() + E.F.image_offset[i];
    E.images[i].width = sc->images[i].width; E.images[i].height = sc->images[i].height;
  }
  S.media = E.F.media.data(); S.perlin_vec = E.F.perlin_vec.data(); S.perlin_perm = E.F.perlin_perm.data();
  S.images = E.images.data();
  S.n_media = sc->n_media; S.n_perlin = sc->n_perlin; S.n_images = sc->n_images;
  const rt_camera &c = sc->camera;
  E.cam.origin = v3_from(c.origin); E.cam.llc = v3_from(c.lower_left_corner);
  E.cam.horizontal = v3_from(c.horizontal); E.cam.vertical = v3_from(c.vertical);
  E.cam.u = v3_from(c.u); E.cam.v = v3_from(c.v);
  E.cam.lens_radius = c.lens_radius; E.cam.time0 = c.time0; E.cam.time1 = c.time1;
  E.sp.background = v3_from(sc->background); E.sp.sky_gradient = sc->sky_gradient; E.sp.flags = sc->flags;
  E.sp.t_min = sc->t_min; E.sp.max_depth = sc->max_depth;
}

// serial replay of rt_accel_build (csrc/b200rt.cu) with the same kernel bodies: one unit = one tree (the whole
// scene, the top level of a two-level scene, or one group in object space)
struct EmuUnit {
  int first[3] = {0, 0, 0}, count[3] = {0, 0, 0};
  int n_inst = 0;
  const uint8_t *exclude = nullptr;
  bool classify = true;
  int link_base = 0, leaf_base = 0, end_link = -1;
  bool own_arrays = true; // single-level: the unit sizes E.nodes / E.leaf_prims itself
  int n_small = 0, kept = 0;
  float root_box[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};

static void build_unit(EmuScene &E, const rt_scene_desc *sc, EmuUnit &U, int quality, float big_frac, int big_rounds, int shuffle, int max_leaf) {
  const int ns = U.count[0], nt = U.count[1], nq = U.count[2], n_geom = ns + nt + nq;
  int n = n_geom + U.n_inst;
  if (n == 0) return;
  BuildArrays B;
  memset(&B, 0, sizeof B);
  B.n_prims = n; B.n_spheres = ns; B.n_tris = nt; B.n_quads = nq; B.n_inst = U.n_inst;
  B.spheres = sc->spheres + U.first[0]; B.tris = sc->triangles + U.first[1]; B.quads = sc->quads + U.first[2];
  for (int k = 0; k < 3; k++) B.id_base[k] = U.first[k];
  B.inst_lo = E.inst_lo.data(); B.inst_hi = E.inst_hi.data();
  // geometry that is not a leaf of this unit (owned by groups / sides of the unit's boxes): build_unit of csrc/b200rt.cu
  std::vector<uint8_t> mask((size_t)std::max(n_geom, 1), 0);
  std::vector<int> unit_boxes;
  if (U.exclude) for (int i = 0; i < n_geom; i++) mask[i] = U.exclude[i];
  for (int b = 0; b < (int)E.F.box.size() / 2; b++) {
    const int fq = RT_F2I(E.F.box[2 * (size_t)b].w);
    if (fq < U.first[2] || fq + 6 > U.first[2] + nq) continue;
    if (U.exclude && U.exclude[(size_t)sc->n_spheres + sc->n_triangles + fq]) continue;
    unit_boxes.push_back(b);
    for (int k = 0; k < 6; k++) mask[(size_t)ns + nt + (fq - U.first[2]) + k] = 1;
  }
  n += (int)unit_boxes.size();
  B.n_prims = n; B.n_boxes = (int)unit_boxes.size();
  B.box_rec = E.F.box.data(); B.box_ids = unit_boxes.data();
  B.exclude = mask.data();
  B.link_base = U.link_base; B.leaf_base = U.leaf_base; B.end_link = U.end_link;
  B.thickness = (sc->flags & RT_FLAG_FLIP_NORMALS) ? 0.01f : 0.1f;
  std::vector<float4> lo(n), hi(n);
  std::vector<int> flag(n);
  BuildBounds bounds[4];
  for (int r = 0; r < 4; r++) for (int a = 0; a < 3; a++) { bounds[r].lo[a] = 0x7fffffff; bounds[r].hi[a] = (int)0x80000000; }
  B.pbox_lo = lo.data(); B.pbox_hi = hi.data(); B.big_flag = flag.data(); B.bounds = bounds;
  for (int i = 0; i < n; i++) body_prim_box(B, i);
  int final_round = 0, n_big = 0;
  if (U.classify) {
    for (int r = 0; r < big_rounds; r++) for (int i = 0; i < n; i++) body_classify(B, i, r, big_frac);
    for (int i = 0; i < n; i++) n_big += flag[i] == 1;
    final_round = big_rounds;
    if (n_big > 32) { for (int &f : flag) if (f == 1) f = 0; n_big = 0; final_round = 0; }
  }
  std::vector<int> small;
  for (int i = 0; i < n; i++) {
    if (flag[i] == 1) {
      E.big.push_back(i < ns ? RT_PRIM_ID(RT_PRIM_SPHERE, U.first[0] + i)
                             : (i < ns + nt ? RT_PRIM_ID(RT_PRIM_TRIANGLE, U.first[1] + i - ns) : RT_PRIM_ID(RT_PRIM_QUAD, U.first[2] + i - ns - nt)));
      if (i < ns) E.sph_is_big[U.first[0] + i] = 1;
    } else if (flag[i] == 0) small.push_back(i);
  }
  const int nsm = (int)small.size();
  U.n_small = nsm;
  if (nsm == 0) return;
  const int n_nodes = 2 * nsm - 1;
  int n_pad = 2;
  while (n_pad < nsm) n_pad <<= 1;
  B.n_small = nsm; B.n_pad = n_pad; B.small_gid = small.data();
  std::vector<unsigned long long> keys(n_pad);
  std::vector<int> left(nsm), right(nsm), parent(n_nodes), nflag(nsm), size(n_nodes), lcnt(n_nodes);
  std::vector<float4> nlo(n_nodes), nhi(n_nodes);
  if (U.own_arrays) {
    E.leaf_prims.assign(nsm, 0);
    E.nodes.assign(2 * (size_t)n_nodes, make_float4(0, 0, 0, 0));
  }
  B.keys = keys.data(); B.left = left.data(); B.right = right.data(); B.parent = parent.data(); B.flag = nflag.data();
  B.size = size.data(); B.lcnt = lcnt.data(); B.leaf_prims = E.leaf_prims.data();
  B.max_leaf = max_leaf < 1 ? 1 : (max_leaf > 8 ? 8 : max_leaf);
  if (U.n_inst) B.max_leaf = 1;
  B.nbox_lo = nlo.data(); B.nbox_hi = nhi.data(); B.packed = E.nodes.data();
  for (int i = 0; i < n_pad; i++) body_morton(B, i, final_round);
  for (int k = 2; k <= n_pad; k <<= 1) for (int j = k >> 1; j > 0; j >>= 1) for (int i = 0; i < n_pad; i++) body_bitonic(B.keys, i, j, k);
  for (int i = 0; i < nsm - 1; i++) body_karras(B, i);
  const int rounds = quality > 0 ? 1 + 2 * std::min(quality, 4) : 1;
  for (int r = 0; r < rounds; r++) {
    if (r > 0) std::fill(nflag.begin(), nflag.end(), 0);
    // thread order is arbitrary on the GPU: optionally visit leaves in a scrambled order
    for (int t = 0; t < nsm; t++) {
      int i = shuffle ? (int)(((long long)t * 7919 + 13) % nsm) : t;
      body_fit(B, i, (quality > 0 && r < rounds - 1) ? 1 : 0);
    }
    if (shuffle) { // 7919 may share a factor with nsm: make sure every leaf ran
      std::vector<char> seen(nsm, 0);
      for (int t = 0; t < nsm; t++) seen[(int)(((long long)t * 7919 + 13) % nsm)] = 1;
      for (int i = 0; i < nsm; i++) if (!seen[i]) body_fit(B, i, (quality > 0 && r < rounds - 1) ? 1 : 0);
    }
  }
  std::vector<int> swapmask(std::max(nsm - 1, 1), 0);
  B.swapmask = swapmask.data();
  B.packed_stride = 2 * n_nodes;
  for (int i = 0; i < nsm - 1; i++) body_order(B, i);
  for (int v = 0; v < n_nodes; v++) body_pack(B, v, 0);
  U.kept = size[0];
  memcpy(U.root_box, &E.nodes[2 * (size_t)U.link_base], sizeof U.root_box);
  if (U.own_arrays) E.nodes.resize(2 * (size_t)size[0]); // kept nodes after leaf collapsing
}

static void build(EmuScene &E, const rt_scene_desc *sc, int quality, float big_frac, int big_rounds, int shuffle, int max_leaf) {
  const int ns = sc->n_spheres, nt = sc->n_triangles, nq = sc->n_quads, n = ns + nt + nq;
  E.nodes.clear(); E.big.clear(); E.leaf_prims.clear(); E.sph_is_big.assign(std::max(ns, 1), 0);
  E.two_level = sc->n_groups > 0 || sc->n_instances > 0;
  E.n_nodes_total = 0;
  if (n == 0) return;
  if (!E.two_level) {
    EmuUnit U;
    U.count[0] = ns; U.count[1] = nt; U.count[2] = nq;
    build_unit(E, sc, U, quality, big_frac, big_rounds, shuffle, max_leaf);
    E.n_nodes_total = (int)E.nodes.size() / 2;
    return;
  }
  // the layout of rt_accel_build: [top level, worst case][group 0][group 1]...
  const int ng = sc->n_groups, ni = sc->n_instances;
  int n_world = 0;
  for (int i = 0; i < n; i++) n_world += E.F.grouped[i] ? 0 : 1;
  int node_at = std::max(2 * (n_world + ni) - 1, 1), leaf_at = std::max(n_world + ni, 1);
  std::vector<EmuUnit> G(ng);
  for (int g = 0; g < ng; g++) {
    const rt_group &R = sc->groups[g];
    G[g].first[0] = R.first_sphere; G[g].count[0] = R.n_spheres;
    G[g].first[1] = R.first_triangle; G[g].count[1] = R.n_triangles;
    G[g].first[2] = R.first_quad; G[g].count[2] = R.n_quads;
    G[g].classify = false; G[g].own_arrays = false;
    G[g].link_base = node_at; G[g].leaf_base = leaf_at;
    const int m = R.n_spheres + R.n_triangles + R.n_quads;
    node_at += std::max(2 * m - 1, 0);
    leaf_at += m;
  }
  E.nodes.assign(2 * (size_t)node_at, make_float4(0, 0, 0, 0));
  E.leaf_prims.assign(leaf_at, 0);
  std::vector<int> root_off(ng, 0);
  for (int g = 0; g < ng; g++) {
    G[g].end_link = RT_POP_LINK(node_at);
    build_unit(E, sc, G[g], quality, big_frac, big_rounds, shuffle, max_leaf);
    root_off[g] = G[g].link_base << RT_NODE_SHIFT;
  }
  std::string err;
  make_instance_records(sc->instances, ni, ng, root_off.data(), E.F.inst, err);
  E.inst_lo.assign(std::max(ni, 1), make_float4(0, 0, 0, 0)); E.inst_hi = E.inst_lo;
  for (int i = 0; i < ni; i++) { // upload_instance_boxes of csrc/b200rt.cu
    const rt_instance &I = sc->instances[i];
    const float *rb = G[I.group].root_box;
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int c = 0; c < 8; c++) {
      const double p[3] = {(c & 1) ? rb[4] : rb[0], (c & 2) ? rb[5] : rb[1], (c & 4) ? rb[6] : rb[2]};
      for (int a = 0; a < 3; a++) {
        const double w = (double)I.m[4 * a] * p[0] + (double)I.m[4 * a + 1] * p[1] + (double)I.m[4 * a + 2] * p[2] + (double)I.m[4 * a + 3];
        lo[a] = std::min(lo[a], w); hi[a] = std::max(hi[a], w);
      }
    }
    float l[3], h[3];
    for (int a = 0; a < 3; a++) {
      const double e = 4e-7 * std::max(std::fabs(lo[a]), std::fabs(hi[a])) + 1e-9;
      l[a] = nextafterf((float)(lo[a] - e), -INFINITY);
      h[a] = nextafterf((float)(hi[a] + e), INFINITY);
    }
    E.inst_lo[i] = make_float4(l[0], l[1], l[2], 0.f);
    E.inst_hi[i] = make_float4(h[0], h[1], h[2], 0.f);
  }
  EmuUnit T;
  T.count[0] = ns; T.count[1] = nt; T.count[2] = nq;
  T.n_inst = ni; T.exclude = E.F.grouped.data();
  T.own_arrays = false;
  T.end_link = node_at << RT_NODE_SHIFT;
  build_unit(E, sc, T, quality, big_frac, big_rounds, shuffle, max_leaf);
  E.n_nodes_total = T.n_small > 0 ? node_at : 0;
}

template <int PROFILE, bool GENERAL>
static void render_t(EmuScene *E, int W, int H, int spp_begin, int spp_count, uint64_t seed, int j0, int j1, double *sum, double *sumsq,
                     unsigned long long *stats) {
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  TraceCounters cnt; cnt.box_tests = cnt.prim_tests = 0;
  unsigned long long nseg = 0, npath = 0;
  for (int j = j0; j < j1; j++)
    for (int i = 0; i < W; i++) {
      const int pixel_index = j * W + i;
      double s[3] = {0, 0, 0}, s2[3] = {0, 0, 0};
      for (int smp = spp_begin; smp < spp_begin + spp_count; smp++) {
        Philox4 q = philox4x32_10((uint32_t)pixel_index, (uint32_t)smp, 0u, 0u, k0, k1);
        float x5 = 0.f;
        if (PROFILE == 2 && E->cam.time1 != E->cam.time0) x5 = u01(philox4x32_10((uint32_t)pixel_index, (uint32_t)smp, 0u, 1u, k0, k1).x);
        Ray r = gen_camera_ray<PROFILE>(E->cam, W, H, i, j, u01(q.x), u01(q.y), u01(q.z), u01(q.w), x5);
        V3f beta = v3(1, 1, 1), L = v3(0, 0, 0);
        int bounce = 0;
        uint32_t media_seed = q.w; // the fourth word of the Philox block that produced the current ray (k_render)
        npath++;
        for (;;) {
          HitAcc hm;
          hm.t = INFINITY; hm.id = -1;
          if (GENERAL && E->S.n_media)
            hm = apply_media(E->S.media, E->S.n_media, r.o, r.d, media_seed, hm);
          int inst = -1;
          HitAcc h = trace_closest<PROFILE, GENERAL, true, GENERAL>(E->S, r, E->sp.t_min, hm.t, &cnt, &inst);
          if (h.id < 0) { h = hm; inst = -1; }
          nseg++;
          if (h.id < 0) { L = L + beta * miss_radiance(E->sp, r.d); break; }
          Philox4 qq = philox4x32_10((uint32_t)pixel_index, (uint32_t)smp, (uint32_t)(1 + bounce), 0u, k0, k1);
          media_seed = qq.w;
          bool cont = shade_hit<PROFILE, GENERAL, GENERAL, GENERAL>(E->S, E->sp, r, h, beta, L, qq, inst);
          bounce++;
          if (!cont) break;
          if (bounce >= E->sp.max_depth) {
            if (PROFILE == 2) { if (E->sp.flags & RT_FLAG_DEPTH_BACKGROUND) L = L + beta * E->sp.background; else L = E->sp.background; }
            else L = v3(0, 0, 0);
            break;
          }
        }
        double c[3] = {L.x, L.y, L.z};
        for (int a = 0; a < 3; a++) { s[a] += c[a]; s2[a] += c[a] * c[a]; }
      }
      for (int a = 0; a < 3; a++) { sum[3 * (size_t)pixel_index + a] = s[a]; if (sumsq) sumsq[3 * (size_t)pixel_index + a] = s2[a]; }
    }
  if (stats) { stats[0] = npath; stats[1] = nseg; stats[2] = cnt.box_tests; stats[3] = cnt.prim_tests; }
}


extern "C" {

void *emu_create(const rt_scene_desc *sc, int profile, int quality, float big_frac, int big_rounds, int shuffle, int max_leaf, char *err, int errcap) {
  EmuScene *E = new EmuScene();
  std::string e;
  E->profile = profile;
  if (flatten_scene(sc, profile, E->F, e)) { snprintf(err, errcap, "%s", e.c_str()); delete E; return nullptr; }
  build(*E, sc, quality, big_frac, big_rounds, shuffle, max_leaf);
  bind(*E, sc);
  return E;
}
void emu_destroy(void *p) { delete (EmuScene *)p; }
int emu_counts(void *p, int *n_nodes, int *n_leaf, int *n_big) { EmuScene *E = (EmuScene *)p; *n_nodes = E->S.n_nodes; *n_leaf = (int)E->leaf_prims.size(); *n_big = E->S.n_big; return 0; }
int emu_download(void *p, rt_bvh_node *nodes, int32_t *leaf, int32_t *big) {
  EmuScene *E = (EmuScene *)p;
  memcpy(nodes, E->nodes.data(), sizeof(float4) * E->nodes.size());
  for (size_t i = 0; i < E->nodes.size() / 2; i++) { // byte offsets -> indices (public rt_bvh_node contract)
    nodes[i].escape >>= RT_NODE_SHIFT;
    if (nodes[i].payload >= 0) nodes[i].payload >>= RT_NODE_SHIFT;
  }
  memcpy(leaf, E->leaf_prims.data(), sizeof(int32_t) * E->leaf_prims.size());
  memcpy(big, E->big.data(), sizeof(int32_t) * E->big.size());
  return 0;
}

int emu_trace_inst(void *p, const float *rays, int n, float t_min, float t_max, int use_accel, int32_t *ids, int32_t *insts, float *ts);
int emu_trace(void *p, const float *rays, int n, float t_min, float t_max, int use_accel, int32_t *ids, float *ts, unsigned long long *counters) {
  if (((EmuScene *)p)->two_level) { if (counters) counters[0] = counters[1] = 0; return emu_trace_inst(p, rays, n, t_min, t_max, use_accel, ids, nullptr, ts); }
  EmuScene *E = (EmuScene *)p;
  TraceCounters cnt; cnt.box_tests = cnt.prim_tests = 0;
  unsigned long long cb = 0, cp = 0;
  for (int k = 0; k < n; k++) {
    Ray r;
    r.o = v3(rays[8 * k], rays[8 * k + 1], rays[8 * k + 2]); r.tm = rays[8 * k + 3];
    r.d = v3(rays[8 * k + 4], rays[8 * k + 5], rays[8 * k + 6]);
    HitAcc h;
    cnt.box_tests = cnt.prim_tests = 0;
    if (use_accel) {
      if (E->profile == 0) h = trace_closest<0, false, true>(E->S, r, t_min, t_max, &cnt);
      else if (E->profile == 1) h = trace_closest<1, false, true>(E->S, r, t_min, t_max, &cnt);
      else h = trace_closest<2, true, true>(E->S, r, t_min, t_max, &cnt);
    } else {
      if (E->profile == 0) h = trace_brute<0, false>(E->S, E->sph_is_big.data(), r, t_min, t_max);
      else if (E->profile == 1) h = trace_brute<1, false>(E->S, E->sph_is_big.data(), r, t_min, t_max);
      else h = trace_brute<2, true>(E->S, E->sph_is_big.data(), r, t_min, t_max);
    }
    cb += cnt.box_tests; cp += cnt.prim_tests;
    ids[k] = h.id; ts[k] = h.id >= 0 ? h.t : 0.f;
  }
  if (counters) { counters[0] = cb; counters[1] = cp; }
  return 0;
}

// two-level scenes: also the instance of the hit primitive
int emu_trace_inst(void *p, const float *rays, int n, float t_min, float t_max, int use_accel, int32_t *ids, int32_t *insts, float *ts) {
  EmuScene *E = (EmuScene *)p;
  for (int k = 0; k < n; k++) {
    Ray r;
    r.o = v3(rays[8 * k], rays[8 * k + 1], rays[8 * k + 2]); r.tm = rays[8 * k + 3];
    r.d = v3(rays[8 * k + 4], rays[8 * k + 5], rays[8 * k + 6]);
    int inst = -1;
    HitAcc h;
    if (use_accel) h = trace_closest<2, true, false, true>(E->S, r, t_min, t_max, nullptr, &inst);
    else h = trace_brute<2, true, true>(E->S, E->sph_is_big.data(), r, t_min, t_max, E->F.grouped.data(), &inst);
    ids[k] = h.id; ts[k] = h.id >= 0 ? h.t : 0.f;
    if (insts) insts[k] = h.id >= 0 ? inst : -1;
  }
  return 0;
}

int emu_render(void *p, int W, int H, int spp_begin, int spp_count, uint64_t seed, int j0, int j1, double *sum, double *sumsq,
               unsigned long long *stats) {
  EmuScene *E = (EmuScene *)p;
  if (E->profile == 0) render_t<0, false>(E, W, H, spp_begin, spp_count, seed, j0, j1, sum, sumsq, stats);
  else if (E->profile == 1) render_t<1, false>(E, W, H, spp_begin, spp_count, seed, j0, j1, sum, sumsq, stats);
  else render_t<2, true>(E, W, H, spp_begin, spp_count, seed, j0, j1, sum, sumsq, stats);
  return 0;
}
}

// material colour at a hit point through the device code (material_color -> texture_ext)
extern "C" void emu_texture(void *p_, int material, int n, const int32_t *prim, const float *p /*[n][3]*/,
                            const float *outward /*[n][3]*/, float *rgb /*[n][3]*/) {
  EmuScene *E = (EmuScene *)p_;
  const float4 m0 = E->S.mats[2 * material], m1 = E->S.mats[2 * material + 1];
  for (int k = 0; k < n; k++) {
    V3f c = material_color<true>(E->S, m0, m1, v3_from(p + 3 * k), v3_from(p + 3 * k), v3_from(outward + 3 * k), prim[k]);
    rgb[3 * k] = c.x; rgb[3 * k + 1] = c.y; rgb[3 * k + 2] = c.z;
  }
}

// rt_render_aov through the same per-pixel device code (render_kernels.cuh aov_pixel)
extern "C" int emu_aov(void *p_, int W, int H, int spp, uint64_t seed, float *out) {
  EmuScene *E = (EmuScene *)p_;
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  for (int p = 0; p < W * H; p++) {
    if (E->profile == 0) aov_pixel<0, false>(E->S, E->cam, E->sp, W, H, spp, k0, k1, p, out);
    else if (E->profile == 1) aov_pixel<1, false>(E->S, E->cam, E->sp, W, H, spp, k0, k1, p, out);
    else aov_pixel<2, true, true>(E->S, E->cam, E->sp, W, H, spp, k0, k1, p, out);
  }
  return 0;
}

#include "warpsim.inc"

extern "C" void emu_philox(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1, unsigned *out) {
  Philox4 q = philox4x32_10(c0, c1, c2, c3, k0, k1);
  out[0] = q.x; out[1] = q.y; out[2] = q.z; out[3] = q.w;
}
extern "C" void emu_samplers(float u1, float u2, float u3, float *out9) {
  float x, y;
  sample_unit_disk(u1, u2, &x, &y);
  V3f a = sample_unit_vector(u1, u2), b = sample_unit_ball(u1, u2, u3);
  out9[0] = x; out9[1] = y; out9[2] = a.x; out9[3] = a.y; out9[4] = a.z; out9[5] = b.x; out9[6] = b.y; out9[7] = b.z; out9[8] = 0;
}
