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
};

static void bind(EmuScene &E, const rt_scene_desc *sc) {
  DevScene &S = E.S;
  memset(&S, 0, sizeof S);
  S.nodes = E.nodes.data(); S.sph = E.F.sph.data(); S.sph_mv = E.F.sph_mv.data(); S.sph_t0 = E.F.sph_t0.data(); S.sph_k = E.F.sph_k.data();
  S.tri = E.F.tri.data(); S.tri_n = E.F.tri_n.data(); S.quad = E.F.quad.data();
  S.sph_mat = E.F.sph_mat.data(); S.tri_mat = E.F.tri_mat.data(); S.quad_mat = E.F.quad_mat.data();
  S.mats = E.F.mats.data(); S.big = E.big.data(); S.leaf_prims = E.leaf_prims.data();
  S.n_nodes = (int)E.nodes.size() / 2; S.n_big = (int)E.big.size();
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

// serial replay of rt_accel_build (csrc/b200rt.cu) with the same kernel bodies
static void build(EmuScene &E, const rt_scene_desc *sc, int quality, float big_frac, int big_rounds, int shuffle, int max_leaf, int quadrant = 0) {
  const int ns = sc->n_spheres, nt = sc->n_triangles, nq = sc->n_quads, n = ns + nt + nq;
  E.nodes.clear(); E.big.clear(); E.leaf_prims.clear(); E.sph_is_big.assign(std::max(ns, 1), 0);
  if (n == 0) return;
  BuildArrays B;
  memset(&B, 0, sizeof B);
  B.n_prims = n; B.n_spheres = ns; B.n_tris = nt; B.n_quads = nq;
  B.spheres = sc->spheres; B.tris = sc->triangles; B.quads = sc->quads;
  B.thickness = (sc->flags & RT_FLAG_FLIP_NORMALS) ? 0.01f : 0.1f;
  std::vector<float4> lo(n), hi(n);
  std::vector<int> flag(n);
  BuildBounds bounds[4];
  for (int r = 0; r < 4; r++) for (int a = 0; a < 3; a++) { bounds[r].lo[a] = 0x7fffffff; bounds[r].hi[a] = (int)0x80000000; }
  B.pbox_lo = lo.data(); B.pbox_hi = hi.data(); B.big_flag = flag.data(); B.bounds = bounds;
  for (int i = 0; i < n; i++) body_prim_box(B, i);
  for (int r = 0; r < big_rounds; r++) for (int i = 0; i < n; i++) body_classify(B, i, r, big_frac);
  int n_big = 0;
  for (int i = 0; i < n; i++) n_big += flag[i];
  int final_round = big_rounds;
  if (n_big > 32) { std::fill(flag.begin(), flag.end(), 0); n_big = 0; final_round = 0; }
  std::vector<int> small;
  for (int i = 0; i < n; i++) {
    if (flag[i]) {
      E.big.push_back(i < ns ? RT_PRIM_ID(RT_PRIM_SPHERE, i) : (i < ns + nt ? RT_PRIM_ID(RT_PRIM_TRIANGLE, i - ns) : RT_PRIM_ID(RT_PRIM_QUAD, i - ns - nt)));
      if (i < ns) E.sph_is_big[i] = 1;
    } else small.push_back(i);
  }
  const int nsm = (int)small.size();
  if (nsm == 0) return;
  const int n_nodes = 2 * nsm - 1;
  int n_pad = 2;
  while (n_pad < nsm) n_pad <<= 1;
  B.n_small = nsm; B.n_pad = n_pad; B.small_gid = small.data();
  std::vector<unsigned long long> keys(n_pad);
  std::vector<int> left(nsm), right(nsm), parent(n_nodes), nflag(nsm), size(n_nodes), lcnt(n_nodes);
  E.leaf_prims.assign(nsm, 0);
  std::vector<float4> nlo(n_nodes), nhi(n_nodes);
  E.nodes.assign(2 * (size_t)n_nodes, make_float4(0, 0, 0, 0));
  B.keys = keys.data(); B.left = left.data(); B.right = right.data(); B.parent = parent.data(); B.flag = nflag.data();
  B.size = size.data(); B.lcnt = lcnt.data(); B.leaf_prims = E.leaf_prims.data();
  B.max_leaf = max_leaf < 1 ? 1 : (max_leaf > 8 ? 8 : max_leaf);
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
  for (int v = 0; v < n_nodes; v++) body_pack(B, v, quadrant);
  E.nodes.resize(2 * (size_t)size[0]); // kept nodes after leaf collapsing
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
        npath++;
        for (;;) {
          HitAcc hm;
          hm.t = INFINITY; hm.id = -1;
          if (GENERAL && E->S.n_media)
            hm = apply_media(E->S.media, E->S.n_media, r.o, r.d, (uint32_t)pixel_index, (uint32_t)smp, (uint32_t)bounce, k0, k1, hm);
          HitAcc h = trace_closest<PROFILE, GENERAL, true>(E->S, r, E->sp.t_min, hm.t, &cnt);
          if (h.id < 0) h = hm;
          nseg++;
          if (h.id < 0) { L = L + beta * miss_radiance(E->sp, r.d); break; }
          Philox4 qq = philox4x32_10((uint32_t)pixel_index, (uint32_t)smp, (uint32_t)(1 + bounce), 0u, k0, k1);
          bool cont = shade_hit<PROFILE, GENERAL, GENERAL>(E->S, E->sp, r, h, beta, L, qq);
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

int emu_trace(void *p, const float *rays, int n, float t_min, float t_max, int use_accel, int32_t *ids, float *ts, unsigned long long *counters) {
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
    V3f c = material_color<true>(E->S, m0, m1, v3_from(p + 3 * k), v3_from(outward + 3 * k), prim[k]);
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
    else aov_pixel<2, true>(E->S, E->cam, E->sp, W, H, spp, k0, k1, p, out);
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
