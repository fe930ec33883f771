// scene_flatten.h — host-side flattening of an rt_scene_desc into the SoA arrays the
// kernels read (layout documented in rt_common.cuh DevScene). Replaces the device-side
// `new sphere(...)/new lambertian(...)` object graph of create_world<<<1,1>>>
// (accelerated-rt-cuda/final.cu:100-143, rt_next_week/cuda/main.cu:386-467).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "rt_common.cuh"

struct HostFlat {
  std::vector<float4> sph, sph_mv, tri, tri_n, quad, mats;
  std::vector<float> sph_t0, sph_k;
  std::vector<int32_t> sph_mat, tri_mat, quad_mat;
  std::vector<float4> media, perlin_vec;
  std::vector<uint8_t> perlin_perm, image_bytes;
  std::vector<size_t> image_offset; // into image_bytes, per image
  bool any_moving = false;
  // two-level scenes: device instance records (rt_common.cuh DevScene::inst), and per primitive in unified order
  // (spheres, triangles, quads) whether some group owns it (= not part of the world level)
  std::vector<float4> inst;
  std::vector<uint8_t> grouped;
  // rects that form boxes (six consecutive rects in box.h's order): device records + per rect "inside a box"
  std::vector<float4> box;
  std::vector<uint8_t> quad_in_box;
};

static inline int flat_fail(std::string &err, const char *fmt, ...) {
  char buf[256];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  err = buf;
  return RT_ERR_INVALID;
}

// Frame of the 16-bit quantised node records of a tree with this root box {bmin.xyz, -, bmax.xyz, -}: plane =
// qbase + q * qscale, q = 0 .. q_max - 8 over the box (room for the outward padding steps of the staging loop)
static inline void quant_frame(const float *root_box, int q_max, float qbase[3], float qscale[3], float qinv[3]) {
  for (int a = 0; a < 3; a++) {
    const float lo = root_box[a], hi = root_box[4 + a];
    const float ext = std::max(hi - lo, 1e-6f * std::max(std::max(std::fabs(lo), std::fabs(hi)), 1e-30f));
    qbase[a] = lo;
    qscale[a] = ext * (1.0f / (float)(q_max - 8));
    qinv[a] = 1.0f / qscale[a];
  }
}

// Device records of rt_instance[n]: the inverse of the rigid map M = [R | T] in double (R^T, -R^T T), rounded
// once. `root_offset_of_group` (byte offsets of the groups' root nodes) may be null before the build.
// group_frames: per group {qbase[3], qscale[3]} of its tree (null before the build).
static inline int make_instance_records(const rt_instance *inst, int n, int n_groups, const int *root_offset_of_group,
                                        std::vector<float4> &out, std::string &err, const float *group_frames = nullptr) {
  out.assign(RT_INST_STRIDE * (size_t)n, make_float4(0, 0, 0, 0));
  for (int i = 0; i < n; i++) {
    const float *m = inst[i].m;
    for (int k = 0; k < 12; k++)
      if (!std::isfinite(m[k])) return flat_fail(err, "instance %d: non-finite matrix", i);
    if (inst[i].group < 0 || inst[i].group >= n_groups) return flat_fail(err, "instance %d: group index", i);
    const double R[3][3] = {{m[0], m[1], m[2]}, {m[4], m[5], m[6]}, {m[8], m[9], m[10]}}, T[3] = {m[3], m[7], m[11]};
    // rigid: R R^T = I (1e-4) and det R = +1 - the ray parameter t must mean the same in both spaces
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) {
        const double d = R[a][0] * R[b][0] + R[a][1] * R[b][1] + R[a][2] * R[b][2];
        if (std::fabs(d - (a == b ? 1.0 : 0.0)) > 1e-4) return flat_fail(err, "instance %d: matrix is not rigid (rotation + translation only)", i);
      }
    const double det = R[0][0] * (R[1][1] * R[2][2] - R[1][2] * R[2][1]) - R[0][1] * (R[1][0] * R[2][2] - R[1][2] * R[2][0]) +
                       R[0][2] * (R[1][0] * R[2][1] - R[1][1] * R[2][0]);
    if (det < 0.0) return flat_fail(err, "instance %d: mirrored matrix (det < 0)", i);
    for (int k = 0; k < 3; k++) { // row k of R^T = column k of R
      const double ti = -(R[0][k] * T[0] + R[1][k] * T[1] + R[2][k] * T[2]);
      out[RT_INST_STRIDE * (size_t)i + k] = make_float4((float)R[0][k], (float)R[1][k], (float)R[2][k], (float)ti);
    }
    out[RT_INST_STRIDE * (size_t)i + 3] = make_float4(RT_I2F(root_offset_of_group ? root_offset_of_group[inst[i].group] : 0), RT_I2F(inst[i].group), 0.f, 0.f);
    if (group_frames) {
      const float *f = group_frames + 6 * (size_t)inst[i].group;
      out[RT_INST_STRIDE * (size_t)i + 4] = make_float4(f[0], f[1], f[2], 0.f);
      out[RT_INST_STRIDE * (size_t)i + 5] = make_float4(f[3], f[4], f[5], 0.f);
    }
  }
  return RT_OK;
}

// 2 * RT_MAX_PRIMS nodes * 32 bytes * 8 orderings = 2^31 - one node
#define RT_MAX_PRIMS ((1 << 22) - 1)

static inline int flatten_scene(const rt_scene_desc *sc, int profile, HostFlat &F, std::string &err) {
#define FAIL(...) return flat_fail(err, __VA_ARGS__)
  if (sc->n_spheres < 0 || sc->n_triangles < 0 || sc->n_quads < 0 || sc->n_materials < 0)
    FAIL("negative count");
  if ((sc->n_spheres && !sc->spheres) || (sc->n_triangles && !sc->triangles) || (sc->n_quads && !sc->quads) ||
      (sc->n_materials && !sc->materials))
    FAIL("null array with non-zero count");
  // Node links and array offsets on the device are signed 32-bit BYTE offsets: 2 n - 1 nodes of 32 bytes, times
  // the eight octant orderings addressed from one base (node_stride * 7 + n_nodes * 32 < 2^31), and 64 bytes per
  // flattened triangle. RT_MAX_PRIMS keeps every one of them inside 31 bits.
  if ((long long)sc->n_spheres + sc->n_triangles + sc->n_quads > RT_MAX_PRIMS)
    FAIL("too many primitives: %lld > %d (32-bit byte offsets of the node array)",
         (long long)sc->n_spheres + sc->n_triangles + sc->n_quads, RT_MAX_PRIMS);
  if (sc->max_depth < 1) FAIL("max_depth must be >= 1");
  // non-finite geometry would poison the bounds / Morton codes of the builder: reject it here
  auto finite3 = [](const float *p) { return std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2]); };
  {
    const rt_camera &c = sc->camera;
    if (!finite3(c.origin) || !finite3(c.lower_left_corner) || !finite3(c.horizontal) || !finite3(c.vertical) ||
        !finite3(c.u) || !finite3(c.v) || !std::isfinite(c.lens_radius) || !std::isfinite(c.time0) || !std::isfinite(c.time1))
      FAIL("camera: non-finite field");
    if (!finite3(sc->background) || !std::isfinite(sc->t_min)) FAIL("non-finite background / t_min");
  }
  bool any_moving = false;
  for (int i = 0; i < sc->n_spheres; i++) {
    const rt_sphere &s = sc->spheres[i];
    if (s.material < 0 || s.material >= sc->n_materials) FAIL("sphere %d: material index", i);
    if (!finite3(s.center0) || !finite3(s.center1) || !std::isfinite(s.radius) || !std::isfinite(s.time0) ||
        !std::isfinite(s.time1))
      FAIL("sphere %d: non-finite field", i);
    if (s.radius == 0.0f) FAIL("sphere %d: zero radius", i);
    if (s.moving) {
      any_moving = true;
      if (s.time1 == s.time0) FAIL("moving sphere %d: time0 == time1", i);
    }
  }
  for (int i = 0; i < sc->n_triangles; i++) {
    const rt_triangle &t = sc->triangles[i];
    if (t.material < 0 || t.material >= sc->n_materials) FAIL("triangle %d: material index", i);
    if (!finite3(t.v0) || !finite3(t.v1) || !finite3(t.v2) || !finite3(t.normal)) FAIL("triangle %d: non-finite field", i);
  }
  for (int i = 0; i < sc->n_quads; i++) {
    const rt_quad &q = sc->quads[i];
    if (q.material < 0 || q.material >= sc->n_materials) FAIL("quad %d: material index", i);
    if (q.axis < 0 || q.axis > 2) FAIL("quad %d: axis", i);
    if (!std::isfinite(q.a0) || !std::isfinite(q.a1) || !std::isfinite(q.b0) || !std::isfinite(q.b1) || !std::isfinite(q.k))
      FAIL("quad %d: non-finite field", i);
  }
  for (int i = 0; i < sc->n_materials; i++) {
    const rt_material &m = sc->materials[i];
    if (m.type < 0 || m.type > RT_MAT_ISOTROPIC) FAIL("material %d: type", i);
    if (m.texture < 0 || m.texture > RT_TEX_IMAGE) FAIL("material %d: texture", i);
    if (profile != RT_PROFILE_NEXT_WEEK && (m.type >= RT_MAT_DIFFUSE_LIGHT || m.texture != RT_TEX_SOLID))
      FAIL("material %d: lights/textures/phase functions need profile 2", i);
    if (m.texture == RT_TEX_NOISE && !(m.albedo2[1] >= 0.f && (int)m.albedo2[1] < sc->n_perlin))
      FAIL("material %d: perlin table index", i);
    if (m.texture == RT_TEX_IMAGE && !(m.albedo2[0] >= 0.f && (int)m.albedo2[0] < sc->n_images))
      FAIL("material %d: image index", i);
  }
  if (sc->n_media < 0 || sc->n_perlin < 0 || sc->n_images < 0) FAIL("negative count");
  if ((sc->n_media && !sc->media) || (sc->n_perlin && !sc->perlin) || (sc->n_images && !sc->images))
    FAIL("null array with non-zero count");
  if (sc->n_media && profile != RT_PROFILE_NEXT_WEEK) FAIL("media need profile 2");
  for (int i = 0; i < sc->n_media; i++) {
    const rt_medium &m = sc->media[i];
    if (m.shape < 0 || m.shape > 1) FAIL("medium %d: shape", i);
    if (!(m.density > 0.f)) FAIL("medium %d: density must be positive", i);
    if (m.material < 0 || m.material >= sc->n_materials || sc->materials[m.material].type != RT_MAT_ISOTROPIC)
      FAIL("medium %d: material must be an isotropic phase function", i);
    if (m.shape == 0 && !(m.p1[0] > 0.f)) FAIL("medium %d: radius", i);
    if (!finite3(m.p0) || !finite3(m.p1) || !finite3(m.offset) || !std::isfinite(m.sin_y) || !std::isfinite(m.cos_y) ||
        !std::isfinite(m.density))
      FAIL("medium %d: non-finite field", i);
  }
  for (int i = 0; i < sc->n_perlin; i++)
    for (int k = 0; k < 256; k++) {
      const rt_perlin &p = sc->perlin[i];
      if ((unsigned)p.perm_x[k] > 255u || (unsigned)p.perm_y[k] > 255u || (unsigned)p.perm_z[k] > 255u)
        FAIL("perlin table %d: permutation entry out of range", i);
    }
  for (int i = 0; i < sc->n_images; i++)
    if (sc->images[i].width < 1 || sc->images[i].height < 1 || !sc->images[i].rgb) FAIL("image %d: empty", i);
  if (profile != RT_PROFILE_NEXT_WEEK && (sc->n_triangles || sc->n_quads || any_moving))
    FAIL("triangles, rects and moving spheres need profile 2 (next-week / triangles trees)");

  // instancing
  if (sc->n_groups < 0 || sc->n_instances < 0) FAIL("negative count");
  if ((sc->n_groups && !sc->groups) || (sc->n_instances && !sc->instances)) FAIL("null array with non-zero count");
  if ((sc->n_groups || sc->n_instances) && profile != RT_PROFILE_NEXT_WEEK) FAIL("instances need profile 2");
  // every instance adds a top-level leaf, every group its own tree: the same 32-bit offset budget
  if ((long long)sc->n_spheres + sc->n_triangles + sc->n_quads + sc->n_instances + sc->n_groups > RT_MAX_PRIMS)
    FAIL("too many primitives + instances");
  F.grouped.assign((size_t)sc->n_spheres + sc->n_triangles + sc->n_quads, 0);
  for (int g = 0; g < sc->n_groups; g++) {
    const rt_group &G = sc->groups[g];
    if (G.first_sphere < 0 || G.n_spheres < 0 || (long long)G.first_sphere + G.n_spheres > sc->n_spheres ||
        G.first_triangle < 0 || G.n_triangles < 0 || (long long)G.first_triangle + G.n_triangles > sc->n_triangles ||
        G.first_quad < 0 || G.n_quads < 0 || (long long)G.first_quad + G.n_quads > sc->n_quads)
      FAIL("group %d: primitive range", g);
    for (int i = 0; i < G.n_spheres; i++) F.grouped[(size_t)G.first_sphere + i] = 1;
    for (int i = 0; i < G.n_triangles; i++) F.grouped[(size_t)sc->n_spheres + G.first_triangle + i] = 1;
    for (int i = 0; i < G.n_quads; i++) F.grouped[(size_t)sc->n_spheres + sc->n_triangles + G.first_quad + i] = 1;
  }
  {
    int rc = make_instance_records(sc->instances, sc->n_instances, sc->n_groups, nullptr, F.inst, err);
    if (rc) return rc;
    for (int i = 0; i < sc->n_instances; i++) { // (the traversal enters an instance at its object's root node)
      const rt_group &G = sc->groups[sc->instances[i].group];
      if (G.n_spheres + G.n_triangles + G.n_quads == 0) FAIL("instance %d places an empty group", i);
    }
  }

  const int ns = sc->n_spheres, nt = sc->n_triangles, nq = sc->n_quads, nm = sc->n_materials;
  std::vector<float4> &sph = F.sph, &sph_mv = F.sph_mv, &tri = F.tri, &tri_n = F.tri_n, &quad = F.quad, &mats = F.mats;
  std::vector<float> &sph_t0 = F.sph_t0;
  std::vector<int32_t> &sph_mat = F.sph_mat, &tri_mat = F.tri_mat, &quad_mat = F.quad_mat;
  sph.assign(ns, make_float4(0, 0, 0, 0)); sph_mv.assign(ns, make_float4(0, 0, 0, 0));
  tri.assign(4 * (size_t)nt, make_float4(0, 0, 0, 0)); tri_n.assign(nt, make_float4(0, 0, 0, 0));
  quad.assign(2 * (size_t)nq, make_float4(0, 0, 0, 0)); mats.assign(2 * (size_t)nm, make_float4(0, 0, 0, 0));
  sph_t0.assign(ns, 0.f); F.sph_k.assign(ns, 0.f); sph_mat.assign(ns, 0); tri_mat.assign(nt, 0); quad_mat.assign(nq, 0);
  for (int i = 0; i < ns; i++) {
    const rt_sphere &s = sc->spheres[i];
    sph[i] = make_float4(s.center0[0], s.center0[1], s.center0[2], s.radius);
    if (s.moving)
      sph_mv[i] = make_float4(s.center1[0] - s.center0[0], s.center1[1] - s.center0[1], s.center1[2] - s.center0[2],
                              1.0f / (s.time1 - s.time0));
    else
      sph_mv[i] = make_float4(0, 0, 0, 0);
    sph_t0[i] = s.time0;
    F.sph_k[i] = (float)((double)s.center0[0] * s.center0[0] + (double)s.center0[1] * s.center0[1] +
                         (double)s.center0[2] * s.center0[2] - (double)s.radius * s.radius);
    sph_mat[i] = s.material;
  }
  // The origin-expanded quadratic of hit_sphere_big (intersect.cuh) is the accurate one only for a sphere much
  // larger than the region the rays live in (the r = 1000 ground under a 22-unit scene): its error grows like
  // (R^2 + 2 R |c|) / 2r with R = how far ray origins lie from the WORLD origin, the ordinary form's like |c| + R.
  // R = reach of the primitives that are small against the sphere (< r / 4); K = NaN marks "use the ordinary form".
  {
    std::vector<std::pair<float, float>> reach; // (size, distance of the farthest point from the world origin)
    auto len3 = [](const float *p) { return std::sqrt((double)p[0] * p[0] + (double)p[1] * p[1] + (double)p[2] * p[2]); };
    for (int i = 0; i < ns; i++) reach.push_back({std::fabs(sc->spheres[i].radius), (float)(len3(sc->spheres[i].center0) + std::fabs(sc->spheres[i].radius))});
    for (int i = 0; i < nt; i++) {
      const rt_triangle &t = sc->triangles[i];
      const float e0[3] = {t.v1[0] - t.v0[0], t.v1[1] - t.v0[1], t.v1[2] - t.v0[2]}, e1[3] = {t.v2[0] - t.v0[0], t.v2[1] - t.v0[1], t.v2[2] - t.v0[2]};
      reach.push_back({(float)std::max(len3(e0), len3(e1)), (float)std::max(len3(t.v0), std::max(len3(t.v1), len3(t.v2)))});
    }
    for (int i = 0; i < nq; i++) {
      const rt_quad &q = sc->quads[i];
      const float c[3] = {std::max(std::fabs(q.a0), std::fabs(q.a1)), std::max(std::fabs(q.b0), std::fabs(q.b1)), q.k};
      reach.push_back({std::max(q.a1 - q.a0, q.b1 - q.b0), (float)len3(c)});
    }
    std::vector<std::pair<float, float>> by_size = reach;
    std::sort(by_size.begin(), by_size.end());
    std::vector<float> prefix_reach(by_size.size()); // max reach among the k smallest primitives
    for (size_t k = 0; k < by_size.size(); k++) prefix_reach[k] = std::max(by_size[k].second, k ? prefix_reach[k - 1] : 0.f);
    for (int i = 0; i < ns; i++) {
      const double r = std::fabs(sc->spheres[i].radius), c = len3(sc->spheres[i].center0);
      const size_t k = std::lower_bound(by_size.begin(), by_size.end(), std::make_pair((float)(0.25 * r), -1.0f)) - by_size.begin();
      // camera rays start at the camera: its distance from the world origin counts as reach too (a lone giant
      // sphere seen from next to its surface keeps the expanded form; a finite camera origin is all that is asked)
      const double cam = len3(sc->camera.origin);
      const bool have_cam = std::isfinite(cam) && len3(sc->camera.horizontal) > 0.0; // a zeroed camera says nothing
      const double R = std::max(k > 0 ? (double)prefix_reach[k - 1] : 0.0, have_cam ? cam : 0.0);
      // no evidence at all about where rays start (no small primitives, no camera): the ordinary form is the safe one
      const bool expand = (k > 0 || have_cam) && 2.0 * (R * R + 2.0 * R * c) / (2.0 * r) < c + R;
      if (!expand) F.sph_k[i] = std::nanf("");
    }
  }
  for (int i = 0; i < nt; i++) {
    const rt_triangle &t = sc->triangles[i];
    V3f v0 = v3_from(t.v0), v1 = v3_from(t.v1), v2 = v3_from(t.v2), N = v3_from(t.normal);
    float len = sqrtf(N.x * N.x + N.y * N.y + N.z * N.z);
    if (!(len > 0.f)) FAIL("triangle %d: degenerate normal", i);
    V3f Nu = (1.0f / len) * N; // unit_vector(face_normal), triangle.h:49
    V3f e[3] = {v1 - v0, v2 - v1, v0 - v2};
    V3f vert[3] = {v0, v1, v2};
    tri[4 * (size_t)i] = make_float4(N.x, N.y, N.z, dot(v0, N));
    for (int k = 0; k < 3; k++) {
      V3f m = cross(Nu, e[k]);
      tri[4 * (size_t)i + 1 + k] = make_float4(m.x, m.y, m.z, dot(m, vert[k]));
    }
    tri_n[i] = make_float4(Nu.x, Nu.y, Nu.z, 0.f);
    tri_mat[i] = t.material;
  }
  for (int i = 0; i < nq; i++) {
    const rt_quad &q = sc->quads[i];
    quad[2 * (size_t)i] = make_float4(q.k, q.a0, q.a1, RT_I2F(q.axis));
    quad[2 * (size_t)i + 1] = make_float4(q.b0, q.b1, 0.f, 0.f);
    quad_mat[i] = q.material;
  }
  // `box` objects (rt_next_week/cuda/box.h:41-58): six rects xy@z1, xy@z0, xz@y1, xz@y0, yz@x1, yz@x0 over the same
  // corner pair. Recognised by exact equality of the numbers, so that testing the six sides from the two corners
  // is the same arithmetic as testing the six rects. (B200RT_BOXES=0 keeps six leaves: measurement / verification.)
  F.box.clear();
  F.quad_in_box.assign((size_t)nq, 0);
  {
    const char *e = getenv("B200RT_BOXES");
    const bool on = profile == RT_PROFILE_NEXT_WEEK && !(e && atoi(e) == 0);
    for (int i = 0; on && i + 6 <= nq;) {
      const rt_quad *q = sc->quads + i;
      const float x0 = q[0].a0, x1 = q[0].a1, y0 = q[0].b0, y1 = q[0].b1, z1 = q[0].k, z0 = q[1].k;
      bool is_box = q[0].axis == 2 && q[1].axis == 2 && q[2].axis == 1 && q[3].axis == 1 && q[4].axis == 0 && q[5].axis == 0 &&
                    x0 < x1 && y0 < y1 && z0 < z1 &&
                    q[1].a0 == x0 && q[1].a1 == x1 && q[1].b0 == y0 && q[1].b1 == y1 &&
                    q[2].a0 == x0 && q[2].a1 == x1 && q[2].b0 == z0 && q[2].b1 == z1 && q[2].k == y1 &&
                    q[3].a0 == x0 && q[3].a1 == x1 && q[3].b0 == z0 && q[3].b1 == z1 && q[3].k == y0 &&
                    q[4].a0 == y0 && q[4].a1 == y1 && q[4].b0 == z0 && q[4].b1 == z1 && q[4].k == x1 &&
                    q[5].a0 == y0 && q[5].a1 == y1 && q[5].b0 == z0 && q[5].b1 == z1 && q[5].k == x0;
      // all six rects must belong to the same trees: none or all owned by groups, and no group boundary inside
      if (is_box && sc->n_groups) {
        for (int g = 0; g < sc->n_groups && is_box; g++) {
          const int lo = sc->groups[g].first_quad, hi = lo + sc->groups[g].n_quads;
          const int inside = std::max(0, std::min(hi, i + 6) - std::max(lo, i)); // rects of this box the group owns
          if (inside != 0 && inside != 6) is_box = false;
        }
      }
      if (!is_box) { i++; continue; }
      F.box.push_back(make_float4(x0, y0, z0, RT_I2F(i)));
      F.box.push_back(make_float4(x1, y1, z1, 0.f));
      for (int k = 0; k < 6; k++) F.quad_in_box[(size_t)i + k] = 1;
      i += 6;
    }
  }
  for (int i = 0; i < nm; i++) {
    const rt_material &m = sc->materials[i];
    mats[2 * (size_t)i] = make_float4(m.albedo[0], m.albedo[1], m.albedo[2], RT_I2F(m.type | (m.texture << 8)));
    mats[2 * (size_t)i + 1] = make_float4(m.albedo2[0], m.albedo2[1], m.albedo2[2], m.param);
  }
  F.media.clear(); F.perlin_vec.clear(); F.perlin_perm.clear(); F.image_bytes.clear(); F.image_offset.clear();
  for (int i = 0; i < sc->n_media; i++) {
    const rt_medium &m = sc->media[i];
    F.media.push_back(make_float4(m.p0[0], m.p0[1], m.p0[2], RT_I2F(m.shape)));
    F.media.push_back(make_float4(m.p1[0], m.p1[1], m.p1[2], -1.0f / m.density));
    F.media.push_back(make_float4(m.offset[0], m.offset[1], m.offset[2], RT_I2F(m.material)));
    F.media.push_back(make_float4(m.shape == 1 ? m.sin_y : 0.f, m.shape == 1 ? m.cos_y : 1.f, 0.f, 0.f));
  }
  for (int i = 0; i < sc->n_perlin; i++) {
    const rt_perlin &p = sc->perlin[i];
    for (int k = 0; k < 256; k++) F.perlin_vec.push_back(make_float4(p.ranvec[k][0], p.ranvec[k][1], p.ranvec[k][2], 0.f));
    for (int k = 0; k < 256; k++) F.perlin_perm.push_back((uint8_t)p.perm_x[k]);
    for (int k = 0; k < 256; k++) F.perlin_perm.push_back((uint8_t)p.perm_y[k]);
    for (int k = 0; k < 256; k++) F.perlin_perm.push_back((uint8_t)p.perm_z[k]);
  }
  for (int i = 0; i < sc->n_images; i++) {
    const rt_image &im = sc->images[i];
    F.image_offset.push_back(F.image_bytes.size());
    const size_t bytes = (size_t)im.width * im.height * 3;
    F.image_bytes.insert(F.image_bytes.end(), im.rgb, im.rgb + bytes);
    while (F.image_bytes.size() & 15) F.image_bytes.push_back(0);
  }
  F.any_moving = any_moving;
  return RT_OK;
#undef FAIL
}
