// oracle/ref_harness.cpp — ORACLE L0 (test infrastructure, NOT product code).
//
// Drives the UNMODIFIED reference CPU renderer (rt_in_one_weekend/main.cpp and
// its headers) where it lies under /root/reference: the file is pulled in with
// `#include`, its `main` renamed, and its own functions are called
// (`random_scene` main.cpp:86-131, `worker` main.cpp:267-290, `ray_color`
// main.cpp:57-83, `hittable_list::hit` hittable_list.h:20-34, `camera`
// camera.h:8-62, `write_color` color.h:14-28). No reference source is copied
// into this repository; the built library goes to oracle/_ref/ (git-ignored).
//
// Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl
// reference) may load the library built from this file.
//
// Build: see oracle/Makefile (g++ — scene determinism depends on GCC's
// argument-evaluation order).

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <sstream>
#include <thread>
#include <vector>

#define main reference_main
#include "rt_in_one_weekend/main.cpp"
#undef main

namespace {

hittable_list g_world;
bool g_built = false;

// tagging proxy over the reference's public virtual interface: records which
// list index produced the accepted hit.
struct tag_proxy : public hittable {
  shared_ptr<hittable> inner;
  int index;
  int *last;
  tag_proxy(shared_ptr<hittable> h, int i, int *l) : inner(h), index(i), last(l) {}
  bool hit(const ray &r, double t_min, double t_max, hit_record &rec) const override {
    if (inner->hit(r, t_min, t_max, rec)) {
      *last = index;
      return true;
    }
    return false;
  }
};

// counting wrapper around a whole world: one call == one ray segment.
struct count_proxy : public hittable {
  const hittable *inner;
  std::atomic<uint64_t> *n;
  count_proxy(const hittable *h, std::atomic<uint64_t> *c) : inner(h), n(c) {}
  bool hit(const ray &r, double t_min, double t_max, hit_record &rec) const override {
    n->fetch_add(1, std::memory_order_relaxed);
    return inner->hit(r, t_min, t_max, rec);
  }
};

struct null_buf : public std::streambuf {
  int overflow(int c) override { return c == EOF ? 0 : c; }
  std::streamsize xsputn(const char *, std::streamsize n) override { return n; }
};

bool g_cam_override = false;
double g_cam22[22];

camera make_camera(const double *c) {
  // c = lookfrom[3], lookat[3], vup[3], vfov, aspect, aperture, focus_dist
  camera cam(point3(c[0], c[1], c[2]), point3(c[3], c[4], c[5]), vec3(c[6], c[7], c[8]),
             c[9], c[10], c[11], c[12]);
  if (g_cam_override) {
    // the camera's fields are public (camera.h:64-70): overwrite them with the
    // float-rounded values every other implementation is given, so that all
    // sides see bit-identical inputs.
    vec3 *f[7] = {&cam.origin, &cam.lower_left_corner, &cam.horizontal, &cam.vertical, &cam.u, &cam.v, &cam.w};
    for (int k = 0; k < 7; k++)
      for (int a = 0; a < 3; a++) (*f[k])[a] = g_cam22[3 * k + a];
    cam.lens_radius = g_cam22[21];
  }
  return cam;
}

} // namespace

extern "C" {

// (Re)build the reference scene exactly as a fresh process would: glibc's
// default seed is srand(1). Returns the number of spheres.
int l0_scene_build(unsigned seed) {
  srand(seed);
  g_world = random_scene();
  g_built = true;
  return (int)g_world.objects.size();
}

// Replace the world by spheres built through the reference's own constructors from
// caller-supplied rows (same layout as l0_scene_get) — used to give the reference
// the float-rounded scene that the GPU core and the L1 oracle consume.
int l0_scene_set(const double *rows, int n) {
  hittable_list w;
  for (int i = 0; i < n; i++) {
    const double *o = rows + 12 * i;
    shared_ptr<material> m;
    int kind = (int)o[4];
    if (kind == 0) m = make_shared<lambertian>(color(o[5], o[6], o[7]));
    else if (kind == 1) m = make_shared<metal>(color(o[5], o[6], o[7]), o[8]);
    else if (kind == 2) m = make_shared<dielectric>(o[8]);
    else return -1;
    w.add(make_shared<sphere>(point3(o[0], o[1], o[2]), o[3], m));
  }
  g_world = w;
  g_built = true;
  return n;
}

// cam22 = NULL switches the override off
void l0_camera_override(const double *cam22) {
  g_cam_override = cam22 != nullptr;
  if (cam22) std::memcpy(g_cam22, cam22, sizeof(g_cam22));
}

// rows of 12 doubles: cx cy cz r kind a0 a1 a2 param 0 0 0
// kind 0 = lambertian (a = albedo), 1 = metal (a = albedo, param = fuzz),
// 2 = dielectric (param = ir)
int l0_scene_get(double *out, int cap) {
  if (!g_built) return -1;
  int n = (int)g_world.objects.size();
  if (cap < n) return -2;
  for (int i = 0; i < n; i++) {
    auto sp = std::dynamic_pointer_cast<sphere>(g_world.objects[i]);
    if (!sp) return -3;
    double *o = out + 12 * i;
    std::memset(o, 0, 12 * sizeof(double));
    o[0] = sp->center.x(); o[1] = sp->center.y(); o[2] = sp->center.z(); o[3] = sp->radius;
    if (auto l = std::dynamic_pointer_cast<lambertian>(sp->mat_ptr)) {
      o[4] = 0; o[5] = l->albedo.x(); o[6] = l->albedo.y(); o[7] = l->albedo.z();
    } else if (auto m = std::dynamic_pointer_cast<metal>(sp->mat_ptr)) {
      o[4] = 1; o[5] = m->albedo.x(); o[6] = m->albedo.y(); o[7] = m->albedo.z(); o[8] = m->fuzz;
    } else if (auto d = std::dynamic_pointer_cast<dielectric>(sp->mat_ptr)) {
      o[4] = 2; o[8] = d->ir;
    } else {
      return -4;
    }
  }
  return n;
}

// Camera as the reference constructs it (camera.h:8-45): 22 doubles
// origin, lower_left_corner, horizontal, vertical, u, v, w, lens_radius
void l0_camera_get(const double *cam13, double *out22) {
  camera c = make_camera(cam13);
  const vec3 *f[7] = {&c.origin, &c.lower_left_corner, &c.horizontal, &c.vertical, &c.u, &c.v, &c.w};
  for (int k = 0; k < 7; k++)
    for (int a = 0; a < 3; a++) out22[3 * k + a] = (*f[k])[a];
  out22[21] = c.lens_radius;
}

// Closest hit through the reference's own hittable_list::hit. rays [n][6] = o, d.
// id = list index or -1; t = rec.t (undefined -> 0 on miss).
int l0_closest_hit(const double *rays, int n, double t_min, double t_max, int32_t *id, double *t) {
  if (!g_built) return -1;
  int last = -1;
  hittable_list tagged;
  for (size_t i = 0; i < g_world.objects.size(); i++)
    tagged.add(make_shared<tag_proxy>(g_world.objects[i], (int)i, &last));
  for (int k = 0; k < n; k++) {
    const double *q = rays + 6 * k;
    ray r(point3(q[0], q[1], q[2]), vec3(q[3], q[4], q[5]));
    hit_record rec;
    last = -1;
    if (tagged.hit(r, t_min, t_max, rec)) {
      id[k] = last;
      t[k] = rec.t;
    } else {
      id[k] = -1;
      t[k] = 0.0;
    }
  }
  return 0;
}

// Second-closest information for robustness classification of a ray:
// gap[k] = (t2 - t1)/t1 where t2 = closest hit when the winning sphere is
// removed (inf if none); disc_rel[k] = |disc|/hb^2 of the winning sphere
// (sphere.h:24-29). Rays with small gap or small disc_rel sit on silhouettes.
int l0_hit_robustness(const double *rays, int n, double t_min, double t_max, double *gap, double *disc_rel) {
  if (!g_built) return -1;
  int nobj = (int)g_world.objects.size();
  for (int k = 0; k < n; k++) {
    const double *q = rays + 6 * k;
    ray r(point3(q[0], q[1], q[2]), vec3(q[3], q[4], q[5]));
    double best = infinity, second = infinity;
    int best_i = -1;
    double min_abs_disc_rel = infinity;
    for (int i = 0; i < nobj; i++) {
      hit_record rec;
      auto sp = std::static_pointer_cast<sphere>(g_world.objects[i]);
      vec3 oc = r.origin() - sp->center;
      double a = r.direction().length_squared();
      double hb = dot(oc, r.direction());
      double c = oc.length_squared() - sp->radius * sp->radius;
      double disc = hb * hb - a * c;
      double rel = fabs(disc) / (hb * hb + 1e-300);
      if (rel < min_abs_disc_rel) min_abs_disc_rel = rel;
      if (sp->hit(r, t_min, t_max, rec)) {
        if (rec.t < best) { second = best; best = rec.t; best_i = i; }
        else if (rec.t < second) second = rec.t;
      }
    }
    (void)best_i;
    gap[k] = (best < infinity) ? ((second < infinity) ? (second - best) / best : infinity) : infinity;
    disc_rel[k] = min_abs_disc_rel; // the closest-to-tangent sphere decides hit-vs-miss stability
  }
  return 0;
}

// Render rows [j0, j1) x all columns with the reference's own sample loop body
// (main.cpp:277-283), single thread, srand(seed) first. Writes per-pixel sum and
// sum of squares (linear, pre-gamma), index = j*W + i with j=0 the BOTTOM row
// as in the reference. Returns the number of ray segments traced.
uint64_t l0_render(int W, int H, int spp, int depth, const double *cam13, unsigned seed, int j0, int j1,
                   double *sum, double *sumsq) {
  if (!g_built) return 0;
  camera cam = make_camera(cam13);
  std::atomic<uint64_t> nseg{0};
  count_proxy counted(&g_world, &nseg);
  srand(seed);
  for (int j = j0; j < j1; j++) {
    for (int i = 0; i < W; i++) {
      double s[3] = {0, 0, 0}, s2[3] = {0, 0, 0};
      for (int k = 0; k < spp; k++) {
        auto u = (i + random_double()) / (W - 1);
        auto v = (j + random_double()) / (H - 1);
        ray r = cam.get_ray(u, v);
        color c = ray_color(r, counted, depth);
        for (int a = 0; a < 3; a++) { s[a] += c[a]; s2[a] += c[a] * c[a]; }
      }
      size_t idx = (size_t)j * W + i;
      for (int a = 0; a < 3; a++) { sum[3 * idx + a] = s[a]; if (sumsq) sumsq[3 * idx + a] = s2[a]; }
    }
  }
  return nseg.load();
}

// The reference's worker() itself (main.cpp:267-290) over pixel index range
// [start, end), `threads` std::threads splitting that range into contiguous
// batches exactly as parallel_render() does (main.cpp:320-334). Output: per
// pixel colour sums (index j*W+i) for pixels in range; returns wall seconds.
// segments_out (optional) receives the segment count (counting wrapper inside
// a one-element hittable_list so that worker() can take it by value).
double l0_worker_timed(int W, int H, int spp, int depth, const double *cam13, unsigned seed, int start, int end,
                       int threads, double *sum, uint64_t *segments_out) {
  if (!g_built) return -1.0;
  camera cam = make_camera(cam13);
  std::atomic<uint64_t> nseg{0};
  hittable_list world;
  if (segments_out)
    world.add(make_shared<count_proxy>(&g_world, &nseg));
  else
    world = g_world;
  std::vector<shared_ptr<color>> img((size_t)W * H);
  srand(seed);
  std::streambuf *old = std::cerr.rdbuf();
  static null_buf sink;          // stateless: safe to share between the worker threads
  std::cerr.rdbuf(&sink);        // worker() prints its range on stderr (main.cpp:271)
  int size = end - start;
  int batch = (size + threads - 1) / threads;
  auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> pool;
  for (int k = 0; k < threads; k++) {
    int s = start + batch * k, e = std::min(start + batch * (k + 1), end);
    if (s >= e) break;
    pool.emplace_back(worker, s, e, std::ref(img), W, H, world, cam, spp, depth);
  }
  for (auto &t : pool) t.join();
  auto t1 = std::chrono::steady_clock::now();
  std::cerr.rdbuf(old);
  if (sum)
    for (int idx = start; idx < end; idx++)
      for (int a = 0; a < 3; a++) sum[3 * (size_t)idx + a] = (*img[idx])[a];
  if (segments_out) *segments_out = nseg.load();
  return std::chrono::duration<double>(t1 - t0).count();
}

// write_color (color.h:14-28) for one pixel sum -> 3 ints
void l0_write_color(const double *sum3, int spp, int *rgb) {
  std::ostringstream os;
  write_color(os, color(sum3[0], sum3[1], sum3[2]), spp);
  std::istringstream is(os.str());
  is >> rgb[0] >> rgb[1] >> rgb[2];
}

// Scatter one hit through the reference's material classes, for distribution
// tests of the three BSDFs: kind/albedo/param as in l0_scene_get; in: ray dir,
// outward unit normal; the hit point is the origin. Output dir + attenuation.
int l0_scatter(int kind, const double *albedo, double param, const double *d_in, const double *outward_n,
               double *dir_out, double *atten_out) {
  shared_ptr<material> m;
  if (kind == 0) m = make_shared<lambertian>(color(albedo[0], albedo[1], albedo[2]));
  else if (kind == 1) m = make_shared<metal>(color(albedo[0], albedo[1], albedo[2]), param);
  else m = make_shared<dielectric>(param);
  ray rin(point3(0, 0, 0) - vec3(d_in[0], d_in[1], d_in[2]), vec3(d_in[0], d_in[1], d_in[2]));
  hit_record rec;
  rec.p = point3(0, 0, 0);
  rec.t = 1.0;
  rec.mat_ptr = m;
  rec.set_face_nromal(rin, vec3(outward_n[0], outward_n[1], outward_n[2]));
  ray sc;
  color att;
  bool ok = m->scatter(rin, rec, att, sc);
  for (int a = 0; a < 3; a++) { dir_out[a] = sc.direction()[a]; atten_out[a] = att[a]; }
  return ok ? 1 : 0;
}

void l0_srand(unsigned seed) { srand(seed); }

} // extern "C"
