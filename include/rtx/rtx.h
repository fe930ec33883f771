// rtx.h — host-side scene-description surface of the reference, kept source-compatible
// so that existing scenes drop in: the same class names and constructor
// signatures as
//   vec3/point3/color   rt_in_one_weekend/vec3.h:10-101
//   camera              rt_in_one_weekend/camera.h:8-45, rt_next_week/cuda/camera.h:25-61
//   hittable, hittable_list   hittable.h:29-33, hittable_list.h:6-18,
//                             accelerated-rt-cuda/hittable_list.h:9-12 (raw-pointer form)
//   sphere, moving_sphere, triangle, xy/xz/yz_rect   sphere.h:10-11, moving_sphere.h:12-15,
//                             triangles/cuda/include/triangle.h:17-20, aarect.h:15-17,70-72,126-128
//   translate, rotate_y, bvh_node   rt_next_week/cuda/hittable.h:49-190, bvh.h:45-46
//   lambertian, metal, dielectric, diffuse_light, isotropic, solid_color, checker_texture,
//   noise_texture (+ perlin), image_texture
//                             material.h:9-97, rt_next_week/cuda/material.h:27-195, texture.h:13-124,
//                             perlin.h:6-125
//   constant_medium           rt_next_week/cuda/constant_medium.h:10-33 (sphere or box boundary)
// These are DESCRIPTION objects only: nothing here intersects rays. `flatten()` turns an
// object graph into the plain-C rt_scene_desc (include/rt_capi.h) that the GPU core
// consumes; instances (translate / rotate_y) are baked into the vertices, and a
// bvh_node is a request ("build an acceleration structure over these") honoured by
// rt_accel_build on the device.
#ifndef RTX_RTX_H
#define RTX_RTX_H

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <limits>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../rt_capi.h"
#include "image_io.h"

namespace rtx {

using std::make_shared;
using std::shared_ptr;

// The CUDA trees of the reference pass raw `new`-ed pointers around and share them
// (one material on several rects) without ever freeing them; raw-pointer overloads
// therefore BORROW (no ownership), exactly as leaky as the original.
template <class T> inline shared_ptr<T> borrow(T *p) { return shared_ptr<T>(p, [](T *) {}); }

const double infinity = std::numeric_limits<double>::infinity();
const double pi = 3.1415926535897932385;
inline double degrees_to_radians(double d) { return d * pi / 180.0; }
// rtweekend.h:21-29 — the reference's generator (glibc rand), for scenes that call it
inline double random_double() { return rand() / (RAND_MAX + 1.0); }
inline double random_double(double lo, double hi) { return lo + (hi - lo) * random_double(); }

class vec3 {
public:
  vec3() : e{0, 0, 0} {}
  vec3(double e0, double e1, double e2) : e{e0, e1, e2} {}
  double x() const { return e[0]; }
  double y() const { return e[1]; }
  double z() const { return e[2]; }
  vec3 operator-() const { return vec3(-e[0], -e[1], -e[2]); }
  double operator[](int i) const { return e[i]; }
  double &operator[](int i) { return e[i]; }
  vec3 &operator+=(const vec3 &v) { e[0] += v.e[0]; e[1] += v.e[1]; e[2] += v.e[2]; return *this; }
  vec3 &operator*=(double t) { e[0] *= t; e[1] *= t; e[2] *= t; return *this; }
  vec3 &operator/=(double t) { return *this *= 1 / t; }
  double length_squared() const { return e[0] * e[0] + e[1] * e[1] + e[2] * e[2]; }
  double length() const { return std::sqrt(length_squared()); }
  // vec3.h:46-49: g++ evaluates the three arguments right to left; spelled out here
  static vec3 random() { double c = random_double(), b = random_double(), a = random_double(); return vec3(a, b, c); }
  static vec3 random(double lo, double hi) {
    double c = random_double(lo, hi), b = random_double(lo, hi), a = random_double(lo, hi);
    return vec3(a, b, c);
  }
  double e[3];
};
using point3 = vec3;
using color = vec3;
inline vec3 operator+(const vec3 &u, const vec3 &v) { return vec3(u.e[0] + v.e[0], u.e[1] + v.e[1], u.e[2] + v.e[2]); }
inline vec3 operator-(const vec3 &u, const vec3 &v) { return vec3(u.e[0] - v.e[0], u.e[1] - v.e[1], u.e[2] - v.e[2]); }
inline vec3 operator*(const vec3 &u, const vec3 &v) { return vec3(u.e[0] * v.e[0], u.e[1] * v.e[1], u.e[2] * v.e[2]); }
inline vec3 operator*(double t, const vec3 &v) { return vec3(t * v.e[0], t * v.e[1], t * v.e[2]); }
inline vec3 operator*(const vec3 &v, double t) { return t * v; }
inline vec3 operator/(vec3 v, double t) { return (1 / t) * v; }
inline double dot(const vec3 &u, const vec3 &v) { return u.e[0] * v.e[0] + u.e[1] * v.e[1] + u.e[2] * v.e[2]; }
inline vec3 cross(const vec3 &u, const vec3 &v) {
  return vec3(u.e[1] * v.e[2] - u.e[2] * v.e[1], u.e[2] * v.e[0] - u.e[0] * v.e[2], u.e[0] * v.e[1] - u.e[1] * v.e[0]);
}
inline vec3 unit_vector(vec3 v) { return v / v.length(); }
inline std::ostream &operator<<(std::ostream &o, const vec3 &v) { return o << v.e[0] << ' ' << v.e[1] << ' ' << v.e[2]; }

// ---------------------------------------------------------------- textures, materials
class flat_scene;
class abstract_texture {
public:
  virtual ~abstract_texture() {}
  // fills the texture fields of `m`; `second` = the checker's odd colour slot. Tables and images
  // are registered in `fs`.
  virtual bool describe(rt_material &m, bool second, flat_scene &fs) const = 0;
};
class solid_color : public abstract_texture {
public:
  solid_color() {}
  solid_color(color c) : color_value(c) {}
  solid_color(double r, double g, double b) : color_value(r, g, b) {}
  bool describe(rt_material &m, bool second, flat_scene &) const override {
    float *dst = second ? m.albedo2 : m.albedo;
    for (int a = 0; a < 3; a++) dst[a] = (float)color_value[a];
    return true;
  }
  color color_value;
};
class checker_texture : public abstract_texture {
public:
  checker_texture(color c1, color c2) : even(make_shared<solid_color>(c1)), odd(make_shared<solid_color>(c2)) {}
  checker_texture(shared_ptr<abstract_texture> e, shared_ptr<abstract_texture> o) : even(e), odd(o) {}
  checker_texture(abstract_texture *e, abstract_texture *o) : even(borrow(e)), odd(borrow(o)) {}
  bool describe(rt_material &m, bool, flat_scene &fs) const override {
    m.texture = RT_TEX_CHECKER;
    return even->describe(m, false, fs) && odd->describe(m, true, fs);
  }
  shared_ptr<abstract_texture> even, odd;
};
// perlin.h:9-19,76-100: 256 gradient vectors random_vec3(-1,1) and three permutation tables
// (for i = n-1..1: swap p[i], p[random_int(n)]), drawn from the host generator.
class perlin {
public:
  perlin() { fill([]() { return random_double(); }); }
  // tables from a private generator (SplitMix64): independent of the global rand() stream
  explicit perlin(uint64_t seed) {
    uint64_t state = seed;
    fill([&state]() {
      uint64_t z = (state += 0x9E3779B97F4A7C15ull);
      z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
      z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
      z ^= z >> 31;
      return (double)(z >> 11) * (1.0 / 9007199254740992.0);
    });
  }
  rt_perlin table;

private:
  template <class Rnd> void fill(Rnd rnd) {
    for (int i = 0; i < 256; i++)
      for (int a = 0; a < 3; a++) table.ranvec[i][a] = (float)(-1.0 + 2.0 * rnd());
    int32_t *perms[3] = {table.perm_x, table.perm_y, table.perm_z};
    for (int32_t *p : perms) {
      for (int i = 0; i < 256; i++) p[i] = i;
      for (int i = 255; i > 0; i--) {
        int target = (int)(rnd() * 256.0);
        if (target > 255) target = 255;
        int32_t tmp = p[i]; p[i] = p[target]; p[target] = tmp;
      }
    }
  }
};
class noise_texture : public abstract_texture {
public:
  noise_texture() {}
  noise_texture(double sc, void * /*curandState*, unused*/ = nullptr) : scale(sc) {}
  static shared_ptr<noise_texture> seeded(double sc, uint64_t seed) {
    auto t = make_shared<noise_texture>(sc);
    t->noise = perlin(seed);
    return t;
  }
  bool describe(rt_material &m, bool second, flat_scene &fs) const override;
  perlin noise;
  double scale = 1;
};
// image_texture(data, w, h) as in rt_next_week/cuda/texture.h:83-88 (the bytes are copied), or
// from a binary/ASCII PPM file. A missing image gives the reference's debug cyan (:94-96).
class image_texture : public abstract_texture {
public:
  image_texture() {}
  image_texture(const unsigned char *d, int w, int h) : width(w), height(h) {
    if (d && w > 0 && h > 0) data.assign(d, d + (size_t)w * h * 3);
  }
  explicit image_texture(const char *filename) {
    std::ifstream in(filename, std::ios::binary);
    std::string magic;
    int maxv = 0;
    if (!(in >> magic) || (magic != "P6" && magic != "P3")) return;
    auto skip = [&]() { while (in >> std::ws && in.peek() == '#') { std::string l; std::getline(in, l); } };
    skip(); in >> width; skip(); in >> height; skip(); in >> maxv;
    // dimensions from an untrusted header: cap at 32768 x 32768 before sizing the buffer
    if (!in || width < 1 || height < 1 || width > 32768 || height > 32768 || maxv < 1 || maxv > 255) { width = height = 0; return; }
    data.resize((size_t)width * height * 3);
    if (magic == "P6") {
      in.get();
      in.read((char *)data.data(), (std::streamsize)data.size());
    } else {
      for (auto &b : data) { int v = 0; in >> v; b = (unsigned char)v; }
    }
    if (!in) { data.clear(); width = height = 0; }
    else if (maxv != 255) for (auto &b : data) b = (unsigned char)((int)b * 255 / maxv);
  }
  bool describe(rt_material &m, bool second, flat_scene &fs) const override;
  std::vector<unsigned char> data;
  int width = 0, height = 0;
};

class material {
public:
  virtual ~material() {}
  virtual rt_material describe(flat_scene &fs) const = 0;
};
class lambertian : public material {
public:
  lambertian(const color &a) : albedo(make_shared<solid_color>(a)) {}
  lambertian(shared_ptr<abstract_texture> a) : albedo(a) {}
  lambertian(abstract_texture *a) : albedo(borrow(a)) {}
  rt_material describe(flat_scene &fs) const override {
    rt_material m = {};
    m.type = RT_MAT_LAMBERTIAN;
    albedo->describe(m, false, fs);
    return m;
  }
  shared_ptr<abstract_texture> albedo;
};
class metal : public material {
public:
  metal(const color &a, double f) : albedo(a), fuzz(f < 1 ? f : 1) {}
  rt_material describe(flat_scene &) const override {
    rt_material m = {};
    m.type = RT_MAT_METAL;
    for (int a = 0; a < 3; a++) m.albedo[a] = (float)albedo[a];
    m.param = (float)fuzz;
    return m;
  }
  color albedo;
  double fuzz;
};
class dielectric : public material {
public:
  dielectric(double index_of_refraction) : ir(index_of_refraction) {}
  rt_material describe(flat_scene &) const override {
    rt_material m = {};
    m.type = RT_MAT_DIELECTRIC;
    m.param = (float)ir;
    return m;
  }
  double ir;
};
class diffuse_light : public material {
public:
  diffuse_light(color c) : emit(make_shared<solid_color>(c)) {}
  diffuse_light(shared_ptr<abstract_texture> a) : emit(a) {}
  diffuse_light(abstract_texture *a) : emit(borrow(a)) {}
  rt_material describe(flat_scene &fs) const override {
    rt_material m = {};
    m.type = RT_MAT_DIFFUSE_LIGHT;
    emit->describe(m, false, fs);
    return m;
  }
  shared_ptr<abstract_texture> emit;
};
// phase function of a constant_medium (rt_next_week/cuda/material.h:178-195)
class isotropic : public material {
public:
  isotropic(color c) : albedo(make_shared<solid_color>(c)) {}
  isotropic(shared_ptr<abstract_texture> a) : albedo(a) {}
  isotropic(abstract_texture *a) : albedo(borrow(a)) {}
  rt_material describe(flat_scene &fs) const override {
    rt_material m = {};
    m.type = RT_MAT_ISOTROPIC;
    albedo->describe(m, false, fs);
    return m;
  }
  shared_ptr<abstract_texture> albedo;
};

// ---------------------------------------------------------------- flattening
// rotate about Y then translate (the only instance transforms the reference has)
struct transform {
  double c = 1, s = 0; // x' = c x + s z ; z' = -s x + c z   (hittable.h:176-180)
  vec3 offset;
  double scale = 1;
  bool identity() const { return c == 1 && s == 0 && offset[0] == 0 && offset[1] == 0 && offset[2] == 0 && scale == 1; }
  // float arithmetic, as the device-side wrappers of the reference compute it
  void apply(const float in[3], float out[3]) const {
    float x = in[0] * (float)scale, y = in[1] * (float)scale, z = in[2] * (float)scale;
    float fc = (float)c, fs = (float)s;
    out[0] = fc * x + fs * z + (float)offset[0];
    out[1] = y + (float)offset[1];
    out[2] = -fs * x + fc * z + (float)offset[2];
  }
  void rotate(const float in[3], float out[3]) const {
    float fc = (float)c, fs = (float)s;
    out[0] = fc * in[0] + fs * in[2];
    out[1] = in[1];
    out[2] = -fs * in[0] + fc * in[2];
  }
  transform then_rotate_y(double deg) const { // inner rotation applied before *this
    transform t = *this;
    double r = degrees_to_radians(deg), ci = std::cos(r), si = std::sin(r);
    t.c = c * ci - s * si;
    t.s = c * si + s * ci;
    return t;
  }
  transform then_translate(const vec3 &d) const { // inner translation applied before *this
    transform t = *this;
    t.offset = vec3(c * d[0] + s * d[2] + offset[0], d[1] + offset[1], -s * d[0] + c * d[2] + offset[2]);
    return t;
  }
};

class hittable;
class flat_scene {
public:
  // First-class instancing (rt_group / rt_instance, the two-level BVH of the GPU core): an object under
  // translate / rotate_y keeps its OBJECT-space primitives as a group and is placed by a rigid matrix, instead of
  // having the transform baked into its vertices. Off: everything is baked (one single-level tree).
  bool instancing = false;
  std::vector<rt_group> groups;
  std::vector<rt_instance> instances;
  // `obj` under the accumulated transform `t` (called by the translate / rotate_y wrappers)
  inline void place(const hittable &obj, const transform &t);
  std::vector<rt_sphere> spheres;
  std::vector<rt_triangle> triangles;
  std::vector<rt_quad> quads;
  std::vector<rt_material> materials;
  std::vector<rt_medium> media;
  std::vector<rt_perlin> perlin_tables;
  std::vector<rt_image> images; // rgb pointers stay owned by the image_texture objects
  bool wants_accel = false;
  int add_material(const material *m) {
    if (!m) throw std::invalid_argument("primitive without material");
    for (size_t i = 0; i < seen.size(); i++)
      if (seen[i] == m) return (int)i;
    seen.push_back(m);
    rt_material d = m->describe(*this); // may register tables / images
    materials.push_back(d);
    return (int)materials.size() - 1;
  }
  int add_perlin(const perlin *p) {
    for (size_t i = 0; i < seen_perlin.size(); i++)
      if (seen_perlin[i] == p) return (int)i;
    seen_perlin.push_back(p);
    perlin_tables.push_back(p->table);
    return (int)perlin_tables.size() - 1;
  }
  int add_image(const image_texture *t) {
    for (size_t i = 0; i < seen_images.size(); i++)
      if (seen_images[i] == t) return (int)i;
    seen_images.push_back(t);
    rt_image im = {t->width, t->height, t->data.data()};
    images.push_back(im);
    return (int)images.size() - 1;
  }

private:
  bool in_group = false;
  std::vector<const material *> seen;
  std::vector<const perlin *> seen_perlin;
  std::vector<const image_texture *> seen_images;
};

inline bool noise_texture::describe(rt_material &m, bool second, flat_scene &fs) const {
  if (second) return false; // only solid colours nest inside a checker
  m.texture = RT_TEX_NOISE;
  m.albedo[0] = m.albedo[1] = m.albedo[2] = 1.0f; // color(1,1,1) * ... texture.h:68
  m.albedo2[0] = (float)scale;
  m.albedo2[1] = (float)fs.add_perlin(&noise);
  return true;
}
inline bool image_texture::describe(rt_material &m, bool second, flat_scene &fs) const {
  if (second) return false;
  if (data.empty()) { // texture.h:94-96
    m.texture = RT_TEX_SOLID;
    m.albedo[0] = 0.0f; m.albedo[1] = 1.0f; m.albedo[2] = 1.0f;
    return true;
  }
  m.texture = RT_TEX_IMAGE;
  m.albedo[0] = m.albedo[1] = m.albedo[2] = 1.0f;
  m.albedo2[0] = (float)fs.add_image(this);
  return true;
}

class hittable {
public:
  virtual ~hittable() {}
  virtual void flatten(flat_scene &out, const transform &xf) const = 0;
  // shape of this object as the boundary of a constant_medium (sphere or box, possibly under
  // rotate_y / translate); false = unsupported
  virtual bool as_medium_boundary(rt_medium &, const transform &) const { return false; }
  virtual bool is_wrapper() const { return false; }   // translate / rotate_y: keep accumulating the transform
  virtual bool instanceable() const { return true; }  // false: holds something that is not a BVH primitive (a medium)
};

inline void flat_scene::place(const hittable &obj, const transform &t) {
  if (obj.is_wrapper() || !instancing || in_group || t.identity() || t.scale != 1 || !obj.instanceable()) {
    obj.flatten(*this, t); // bake (also: a transform nested inside a group is baked relative to that group)
    return;
  }
  const int s0 = (int)spheres.size(), t0 = (int)triangles.size(), q0 = (int)quads.size();
  in_group = true;
  obj.flatten(*this, transform());
  in_group = false;
  rt_group g = {};
  g.first_sphere = s0; g.n_spheres = (int)spheres.size() - s0;
  g.first_triangle = t0; g.n_triangles = (int)triangles.size() - t0;
  g.first_quad = q0; g.n_quads = (int)quads.size() - q0;
  if (g.n_spheres + g.n_triangles + g.n_quads == 0) return;
  rt_instance I = {};
  const float m[12] = {(float)t.c, 0.f, (float)t.s, (float)t.offset[0], 0.f, 1.f, 0.f, (float)t.offset[1],
                       -(float)t.s, 0.f, (float)t.c, (float)t.offset[2]};
  for (int k = 0; k < 12; k++) I.m[k] = m[k];
  // The reference wraps EVERY triangle of a mesh in its own translate(rotate_y()) (obj_render.cu:498-511):
  // adjacent objects under the same transform form one group under one instance.
  if (!instances.empty()) {
    rt_group &b = groups[instances.back().group];
    bool same = instances.back().group == (int)groups.size() - 1;
    for (int k = 0; k < 12; k++) same = same && instances.back().m[k] == I.m[k];
    same = same && (g.n_spheres == 0 || b.n_spheres == 0 || b.first_sphere + b.n_spheres == g.first_sphere) &&
           (g.n_triangles == 0 || b.n_triangles == 0 || b.first_triangle + b.n_triangles == g.first_triangle) &&
           (g.n_quads == 0 || b.n_quads == 0 || b.first_quad + b.n_quads == g.first_quad);
    if (same) {
      if (g.n_spheres) { if (!b.n_spheres) b.first_sphere = g.first_sphere; b.n_spheres += g.n_spheres; }
      if (g.n_triangles) { if (!b.n_triangles) b.first_triangle = g.first_triangle; b.n_triangles += g.n_triangles; }
      if (g.n_quads) { if (!b.n_quads) b.first_quad = g.first_quad; b.n_quads += g.n_quads; }
      return;
    }
  }
  I.group = (int)groups.size();
  groups.push_back(g);
  instances.push_back(I);
}

// triangle ctor — triangles/cuda/include/triangle.h:17-53 in float arithmetic
inline void triangle_face_normal(const float v0[3], const float v1[3], const float v2[3], const float vn0[3],
                                 const float vn1[3], const float vn2[3], float n[3]) {
  float avg[3], ab[3], ac[3];
  for (int a = 0; a < 3; a++) {
    avg[a] = -((vn0[a] + vn1[a]) + vn2[a]) * (1.0f / 3.0f);
    ab[a] = v1[a] - v0[a];
    ac[a] = v2[a] - v0[a];
  }
  float c[3] = {ab[1] * ac[2] - ab[2] * ac[1], ab[2] * ac[0] - ab[0] * ac[2], ab[0] * ac[1] - ab[1] * ac[0]};
  float d = c[0] * avg[0] + c[1] * avg[1] + c[2] * avg[2];
  for (int a = 0; a < 3; a++) n[a] = d > 0.0f ? c[a] : -c[a];
}

class sphere : public hittable {
public:
  sphere() {}
  sphere(point3 cen, double r, shared_ptr<material> m) : center(cen), radius(r), mat_ptr(m) {}
  sphere(point3 cen, double r, material *m) : center(cen), radius(r), mat_ptr(borrow(m)) {}
  void flatten(flat_scene &out, const transform &xf) const override {
    rt_sphere s = {};
    float c[3] = {(float)center[0], (float)center[1], (float)center[2]};
    if (xf.identity()) for (int a = 0; a < 3; a++) s.center0[a] = c[a];
    else xf.apply(c, s.center0);
    for (int a = 0; a < 3; a++) s.center1[a] = s.center0[a];
    s.radius = (float)(radius * xf.scale);
    s.material = out.add_material(mat_ptr.get());
    s.time0 = 0; s.time1 = 1; s.moving = 0;
    out.spheres.push_back(s);
  }
  bool as_medium_boundary(rt_medium &m, const transform &xf) const override {
    m.shape = 0;
    float c[3] = {(float)center[0], (float)center[1], (float)center[2]};
    xf.apply(c, m.p0);
    m.p1[0] = (float)(std::fabs(radius) * xf.scale); m.p1[1] = m.p1[2] = 0;
    m.sin_y = 0; m.cos_y = 1;
    m.offset[0] = m.offset[1] = m.offset[2] = 0;
    return true;
  }
  point3 center;
  double radius = 0;
  shared_ptr<material> mat_ptr;
};

class moving_sphere : public hittable {
public:
  moving_sphere(point3 c0, point3 c1, double t0, double t1, double r, shared_ptr<material> m)
      : center0(c0), center1(c1), time0(t0), time1(t1), radius(r), mat_ptr(m) {}
  moving_sphere(point3 c0, point3 c1, double t0, double t1, double r, material *m)
      : center0(c0), center1(c1), time0(t0), time1(t1), radius(r), mat_ptr(borrow(m)) {}
  void flatten(flat_scene &out, const transform &xf) const override {
    rt_sphere s = {};
    float a0[3] = {(float)center0[0], (float)center0[1], (float)center0[2]};
    float a1[3] = {(float)center1[0], (float)center1[1], (float)center1[2]};
    if (xf.identity()) for (int a = 0; a < 3; a++) { s.center0[a] = a0[a]; s.center1[a] = a1[a]; }
    else { xf.apply(a0, s.center0); xf.apply(a1, s.center1); }
    s.radius = (float)(radius * xf.scale);
    s.material = out.add_material(mat_ptr.get());
    s.time0 = (float)time0; s.time1 = (float)time1; s.moving = 1;
    out.spheres.push_back(s);
  }
  point3 center0, center1;
  double time0, time1, radius;
  shared_ptr<material> mat_ptr;
};

class triangle : public hittable {
public:
  triangle(vec3 _v0, vec3 _v1, vec3 _v2, vec3 _vn0, vec3 _vn1, vec3 _vn2, shared_ptr<material> m)
      : v0(_v0), v1(_v1), v2(_v2), vn0(_vn0), vn1(_vn1), vn2(_vn2), mat_ptr(m) {}
  triangle(vec3 _v0, vec3 _v1, vec3 _v2, vec3 _vn0, vec3 _vn1, vec3 _vn2, material *m)
      : v0(_v0), v1(_v1), v2(_v2), vn0(_vn0), vn1(_vn1), vn2(_vn2), mat_ptr(borrow(m)) {}
  void flatten(flat_scene &out, const transform &xf) const override {
    rt_triangle t = {};
    const vec3 *v[3] = {&v0, &v1, &v2}, *vn[3] = {&vn0, &vn1, &vn2};
    float fv[3][3], fn[3][3];
    for (int k = 0; k < 3; k++) {
      float a[3] = {(float)(*v[k])[0], (float)(*v[k])[1], (float)(*v[k])[2]};
      float b[3] = {(float)(*vn[k])[0], (float)(*vn[k])[1], (float)(*vn[k])[2]};
      if (xf.identity()) { for (int q = 0; q < 3; q++) { fv[k][q] = a[q]; fn[k][q] = b[q]; } }
      else { xf.apply(a, fv[k]); xf.rotate(b, fn[k]); }
    }
    for (int q = 0; q < 3; q++) { t.v0[q] = fv[0][q]; t.v1[q] = fv[1][q]; t.v2[q] = fv[2][q]; }
    triangle_face_normal(fv[0], fv[1], fv[2], fn[0], fn[1], fn[2], t.normal);
    t.material = out.add_material(mat_ptr.get());
    out.triangles.push_back(t);
  }
  vec3 v0, v1, v2, vn0, vn1, vn2;
  shared_ptr<material> mat_ptr;
};

// aarect.h: axis-aligned rectangles; only translation-free use is supported when flattened
// under a rotation (the reference never rotates its rects in the in-scope scenes).
class aa_rect : public hittable {
public:
  aa_rect(int axis_, double a0_, double a1_, double b0_, double b1_, double k_, shared_ptr<material> m)
      : axis(axis_), a0(a0_), a1(a1_), b0(b0_), b1(b1_), k(k_), mp(m) {}
  void flatten(flat_scene &out, const transform &xf) const override {
    const int ia = (axis == 0) ? 1 : 0, ib = (axis == 2) ? 1 : 2;
    if (xf.s != 0 || xf.c != 1 || xf.scale != 1) {
      // a rotated rect (rotate_y(box(...)) in the Cornell box, main.cu:269-275) is no longer
      // axis aligned: emit it as two triangles whose face normal is the rect's nominal
      // (+axis, never flipped: rt_next_week/cuda/hittable.h:29) normal, rotated
      float c[4][3], w[4][3], n[3] = {0, 0, 0}, nw[3];
      // corner order = counter-clockwise seen from the +axis side, so that the triangle's edge tests (which follow
      // the vertex order, triangle.h:172-202) agree with the +axis face normal: (a, b) = (x, z) is left-handed
      // about +y, so the xz rect runs the other way round
      const double aa1[4] = {a0, a0, a1, a1}, bb1[4] = {b0, b1, b1, b0};
      const double aa0[4] = {a0, a1, a1, a0}, bb0[4] = {b0, b0, b1, b1};
      const double *aa = axis == 1 ? aa1 : aa0, *bb = axis == 1 ? bb1 : bb0;
      for (int q = 0; q < 4; q++) { c[q][axis] = (float)k; c[q][ia] = (float)aa[q]; c[q][ib] = (float)bb[q]; xf.apply(c[q], w[q]); }
      n[axis] = 1.0f;
      xf.rotate(n, nw);
      const int m = out.add_material(mp.get());
      const int tri[2][3] = {{0, 1, 2}, {0, 2, 3}};
      for (int t = 0; t < 2; t++) {
        rt_triangle r = {};
        float vn[3] = {-nw[0], -nw[1], -nw[2]}; // triangle.h:24: the face normal agrees with MINUS the vertex normals
        for (int q = 0; q < 3; q++) { r.v0[q] = w[tri[t][0]][q]; r.v1[q] = w[tri[t][1]][q]; r.v2[q] = w[tri[t][2]][q]; }
        triangle_face_normal(r.v0, r.v1, r.v2, vn, vn, vn, r.normal);
        r.material = m;
        out.triangles.push_back(r);
      }
      return;
    }
    rt_quad q = {};
    q.axis = axis;
    q.a0 = (float)(a0 + xf.offset[ia]); q.a1 = (float)(a1 + xf.offset[ia]);
    q.b0 = (float)(b0 + xf.offset[ib]); q.b1 = (float)(b1 + xf.offset[ib]);
    q.k = (float)(k + xf.offset[axis]);
    q.material = out.add_material(mp.get());
    out.quads.push_back(q);
  }
  int axis;
  double a0, a1, b0, b1, k;
  shared_ptr<material> mp;
};
class xy_rect : public aa_rect {
public:
  xy_rect(double x0, double x1, double y0, double y1, double k, shared_ptr<material> m) : aa_rect(2, x0, x1, y0, y1, k, m) {}
  xy_rect(double x0, double x1, double y0, double y1, double k, material *m) : aa_rect(2, x0, x1, y0, y1, k, borrow(m)) {}
};
class xz_rect : public aa_rect {
public:
  xz_rect(double x0, double x1, double z0, double z1, double k, shared_ptr<material> m) : aa_rect(1, x0, x1, z0, z1, k, m) {}
  xz_rect(double x0, double x1, double z0, double z1, double k, material *m) : aa_rect(1, x0, x1, z0, z1, k, borrow(m)) {}
};
class yz_rect : public aa_rect {
public:
  yz_rect(double y0, double y1, double z0, double z1, double k, shared_ptr<material> m) : aa_rect(0, y0, y1, z0, z1, k, m) {}
  yz_rect(double y0, double y1, double z0, double z1, double k, material *m) : aa_rect(0, y0, y1, z0, z1, k, borrow(m)) {}
};

// box(p0, p1, material) — rt_next_week/cuda/box.h:41-58: six rects in the reference's order
class box : public hittable {
public:
  box(const point3 &p0, const point3 &p1, shared_ptr<material> m) : box_min(p0), box_max(p1) { build(m); }
  box(const point3 &p0, const point3 &p1, material *m) : box_min(p0), box_max(p1) { build(borrow(m)); }
  void flatten(flat_scene &out, const transform &xf) const override {
    for (const auto &s : sides) s->flatten(out, xf);
  }
  bool as_medium_boundary(rt_medium &m, const transform &xf) const override {
    m.shape = 1;
    for (int a = 0; a < 3; a++) {
      m.p0[a] = (float)(box_min[a] * xf.scale);
      m.p1[a] = (float)(box_max[a] * xf.scale);
      m.offset[a] = (float)xf.offset[a];
    }
    m.sin_y = (float)xf.s; m.cos_y = (float)xf.c;
    return true;
  }
  point3 box_min, box_max;
  std::vector<shared_ptr<hittable>> sides;

private:
  void build(shared_ptr<material> m) {
    const point3 &p0 = box_min, &p1 = box_max;
    sides.push_back(make_shared<xy_rect>(p0.x(), p1.x(), p0.y(), p1.y(), p1.z(), m));
    sides.push_back(make_shared<xy_rect>(p0.x(), p1.x(), p0.y(), p1.y(), p0.z(), m));
    sides.push_back(make_shared<xz_rect>(p0.x(), p1.x(), p0.z(), p1.z(), p1.y(), m));
    sides.push_back(make_shared<xz_rect>(p0.x(), p1.x(), p0.z(), p1.z(), p0.y(), m));
    sides.push_back(make_shared<yz_rect>(p0.y(), p1.y(), p0.z(), p1.z(), p1.x(), m));
    sides.push_back(make_shared<yz_rect>(p0.y(), p1.y(), p0.z(), p1.z(), p0.x(), m));
  }
};

class hittable_list : public hittable {
public:
  hittable_list() {}
  hittable_list(shared_ptr<hittable> object) { add(object); }
  hittable_list(hittable **l, int n) { for (int i = 0; i < n; i++) objects.push_back(borrow(l[i])); }
  void clear() { objects.clear(); }
  void add(shared_ptr<hittable> object) { objects.push_back(object); }
  void flatten(flat_scene &out, const transform &xf) const override {
    for (const auto &o : objects) o->flatten(out, xf);
  }
  bool instanceable() const override {
    for (const auto &o : objects) if (!o->instanceable()) return false;
    return true;
  }
  std::vector<shared_ptr<hittable>> objects;
};

class translate : public hittable {
public:
  translate(shared_ptr<hittable> p, const vec3 &d) : ptr(p), offset(d) {}
  translate(hittable *p, const vec3 &d) : ptr(borrow(p)), offset(d) {}
  void flatten(flat_scene &out, const transform &xf) const override { out.place(*ptr, xf.then_translate(offset)); }
  bool is_wrapper() const override { return true; }
  bool instanceable() const override { return ptr->instanceable(); }
  bool as_medium_boundary(rt_medium &m, const transform &xf) const override {
    return ptr->as_medium_boundary(m, xf.then_translate(offset));
  }
  shared_ptr<hittable> ptr;
  vec3 offset;
};
class rotate_y : public hittable {
public:
  rotate_y(shared_ptr<hittable> p, double angle_deg) : ptr(p), angle(angle_deg) {}
  rotate_y(hittable *p, double angle_deg) : ptr(borrow(p)), angle(angle_deg) {}
  void flatten(flat_scene &out, const transform &xf) const override { out.place(*ptr, xf.then_rotate_y(angle)); }
  bool is_wrapper() const override { return true; }
  bool instanceable() const override { return ptr->instanceable(); }
  bool as_medium_boundary(rt_medium &m, const transform &xf) const override {
    return ptr->as_medium_boundary(m, xf.then_rotate_y(angle));
  }
  shared_ptr<hittable> ptr;
  double angle;
};
// constant_medium(boundary, density, colour | texture) — constant_medium.h:10-33. The boundary
// must be convex: a sphere or a box, directly or under rotate_y / translate.
class constant_medium : public hittable {
public:
  constant_medium(shared_ptr<hittable> b, double d, shared_ptr<abstract_texture> a)
      : boundary(b), density(d), phase_function(make_shared<isotropic>(a)) {}
  constant_medium(shared_ptr<hittable> b, double d, color c)
      : boundary(b), density(d), phase_function(make_shared<isotropic>(c)) {}
  constant_medium(hittable *b, double d, abstract_texture *a)
      : boundary(borrow(b)), density(d), phase_function(make_shared<isotropic>(a)) {}
  constant_medium(hittable *b, double d, color c)
      : boundary(borrow(b)), density(d), phase_function(make_shared<isotropic>(c)) {}
  void flatten(flat_scene &out, const transform &xf) const override {
    rt_medium m = {};
    if (!boundary || !boundary->as_medium_boundary(m, xf))
      throw std::invalid_argument("constant_medium: the boundary must be a sphere or a box");
    m.density = (float)density;
    m.material = out.add_material(phase_function.get());
    out.media.push_back(m);
  }
  bool instanceable() const override { return false; } // media are not BVH primitives: their transform is baked
  shared_ptr<hittable> boundary;
  double density;
  shared_ptr<material> phase_function;
};
// bvh_node(list, start, end, time0, time1[, rng]) — bvh.h:45-46: a request for an
// acceleration structure over the range; the device builds one LBVH over everything.
class bvh_node : public hittable {
public:
  bvh_node(const hittable_list &l, double = 0, double = 1) { for (const auto &o : l.objects) objects.push_back(o); }
  bvh_node(hittable **l, size_t start, size_t end, double = 0, double = 1, void * = nullptr) {
    for (size_t i = start; i < end; i++) objects.push_back(borrow(l[i]));
  }
  void flatten(flat_scene &out, const transform &xf) const override {
    out.wants_accel = true;
    for (const auto &o : objects) o->flatten(out, xf);
  }
  bool instanceable() const override {
    for (const auto &o : objects) if (!o->instanceable()) return false;
    return true;
  }
  std::vector<shared_ptr<hittable>> objects;
};

// ---------------------------------------------------------------- camera
class camera {
public:
  // double arithmetic = camera.h:8-45 (CPU); use_float = accelerated-rt-cuda/camera.h:20-53
  camera(point3 lookfrom, point3 lookat, vec3 vup, double vfov, double aspect_ratio, double aperture, double focus_dist,
         double _time0 = 0, double _time1 = 0, bool use_float = false) {
    time0 = _time0; time1 = _time1;
    if (!use_float) {
      double theta = degrees_to_radians(vfov), h = std::tan(theta / 2);
      double viewport_height = 2.0 * h, viewport_width = aspect_ratio * viewport_height;
      w = unit_vector(lookfrom - lookat);
      u = unit_vector(cross(vup, w));
      v = cross(w, u);
      origin = lookfrom;
      horizontal = focus_dist * viewport_width * u;
      vertical = focus_dist * viewport_height * v;
      lower_left_corner = origin - horizontal / 2 - vertical / 2 - focus_dist * w;
      lens_radius = aperture / 2;
    } else {
      auto F = [](double x) { return (double)(float)x; };
      float theta = (float)((double)(float)vfov * M_PI / 180);
      float half_height = (float)std::tan((double)(theta / 2)), half_width = (float)aspect_ratio * half_height;
      float fd = (float)focus_dist;
      float lf[3] = {(float)lookfrom[0], (float)lookfrom[1], (float)lookfrom[2]};
      float wv[3], uv[3], vv[3];
      float d[3] = {lf[0] - (float)lookat[0], lf[1] - (float)lookat[1], lf[2] - (float)lookat[2]};
      float il = 1.0f / std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
      for (int a = 0; a < 3; a++) wv[a] = il * d[a];
      float up[3] = {(float)vup[0], (float)vup[1], (float)vup[2]};
      float c[3] = {up[1] * wv[2] - up[2] * wv[1], up[2] * wv[0] - up[0] * wv[2], up[0] * wv[1] - up[1] * wv[0]};
      il = 1.0f / std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
      for (int a = 0; a < 3; a++) uv[a] = il * c[a];
      vv[0] = wv[1] * uv[2] - wv[2] * uv[1]; vv[1] = wv[2] * uv[0] - wv[0] * uv[2]; vv[2] = wv[0] * uv[1] - wv[1] * uv[0];
      float hs = fd * 2.0f * half_width, vs = fd * 2.0f * half_height;
      for (int a = 0; a < 3; a++) {
        float hz = hs * uv[a], vt = vs * vv[a];
        horizontal[a] = F(hz); vertical[a] = F(vt);
        lower_left_corner[a] = F(((lf[a] - hz * 0.5f) - vt * 0.5f) - fd * wv[a]);
        origin[a] = F(lf[a]); u[a] = F(uv[a]); v[a] = F(vv[a]); w[a] = F(wv[a]);
      }
      lens_radius = F((float)aperture / 2.0f);
    }
  }
  rt_camera describe() const {
    rt_camera c = {};
    for (int a = 0; a < 3; a++) {
      c.origin[a] = (float)origin[a]; c.lower_left_corner[a] = (float)lower_left_corner[a];
      c.horizontal[a] = (float)horizontal[a]; c.vertical[a] = (float)vertical[a];
      c.u[a] = (float)u[a]; c.v[a] = (float)v[a]; c.w[a] = (float)w[a];
    }
    c.lens_radius = (float)lens_radius; c.time0 = (float)time0; c.time1 = (float)time1;
    return c;
  }
  point3 origin, lower_left_corner;
  vec3 horizontal, vertical, u, v, w;
  double lens_radius, time0, time1;
};

// read_triangles — triangles/cuda/include/triangle.h:217-300: `v`, `vn`, `f a/b/c` x3; the
// face is emitted with vertex indices [6],[3],[0] and normal indices [8],[5],[2] of the
// flattened index list (reversed winding). Beyond the reference's parser (which only
// accepts triangles written as a/b/c or a//c): `v`, `v/vt`, `v//vn`, `v/vt/vn` corners,
// negative (relative) indices, polygons (fan triangulation, each triangle then reversed
// like the reference does) and faces without normals (the geometric normal of the
// REVERSED triangle stands in for the vertex normals, so orientation is well defined).
inline void read_triangles(const std::string &filename, std::vector<shared_ptr<hittable>> &out, shared_ptr<material> mat,
                           double scale = 1.0) {
  std::ifstream in(filename);
  if (!in.is_open()) throw std::runtime_error("read failed: " + filename);
  std::vector<vec3> vs, vns;
  std::string line;
  while (std::getline(in, line)) {
    std::istringstream ls(line);
    std::string type;
    ls >> type;
    double x, y, z;
    if (type == "vn") { ls >> x >> y >> z; vns.push_back(vec3(x, y, z)); }
    else if (type == "v") { ls >> x >> y >> z; vs.push_back(vec3(x, y, z)); }
    else if (type == "f") {
      std::vector<int> vi, ni;
      std::string sec;
      while (ls >> sec) {
        int idx[3] = {0, 0, 0}, k = 0;
        std::istringstream ss(sec);
        std::string num;
        while (k < 3 && std::getline(ss, num, '/')) { idx[k++] = num.empty() ? 0 : std::stoi(num); }
        auto fix = [](int i, size_t n) { return i > 0 ? i - 1 : (i < 0 ? (int)n + i : -1); };
        vi.push_back(fix(idx[0], vs.size()));
        ni.push_back(fix(idx[2], vns.size()));
      }
      for (size_t k = 1; k + 1 < vi.size(); k++) { // fan: (0, k, k+1)
        const int a = 0, b = (int)k, c = (int)k + 1;
        const vec3 &p0 = vs.at(vi[c]), &p1 = vs.at(vi[b]), &p2 = vs.at(vi[a]); // reversed winding
        vec3 n0, n1, n2;
        if (ni[a] >= 0 && ni[b] >= 0 && ni[c] >= 0) { n0 = vns.at(ni[c]); n1 = vns.at(ni[b]); n2 = vns.at(ni[a]); }
        else { n0 = n1 = n2 = -cross(p1 - p0, p2 - p0); } // keeps the reversed triangle's own orientation
        out.push_back(make_shared<triangle>(p0 * scale, p1 * scale, p2 * scale, n0, n1, n2, mat));
      }
    }
  }
}

// ---------------------------------------------------------------- render entry
struct render_options {
  int profile = RT_PROFILE_WEEKEND_CPU;
  int device = 0;
  uint64_t seed = 1984;
  color background = color(0, 0, 0);
  bool sky_gradient = true;
  double t_min = 0.001;
  int max_depth = 50;
  uint32_t flags = 0;
  // translate / rotate_y as rt_group + rt_instance (two-level BVH: the object keeps its own tree, a moved object costs
  // rt_instances_update = 0.1-0.5 ms instead of a re-flatten + upload + full build) or baked into the vertices (one
  // tree; a static frame renders 25-35 % faster because no ray ever changes space). Default: baked.
  bool instancing = false;
};

struct image8 {
  int width = 0, height = 0;
  std::vector<uint8_t> rgb; // top row first
  // P3 text on a stream, as the reference writes it (main.cpp:344-355)
  void write_ppm(std::ostream &out) const {
    out << "P3\n" << width << ' ' << height << "\n255\n";
    for (size_t i = 0; i + 2 < rgb.size(); i += 3) out << (int)rgb[i] << ' ' << (int)rgb[i + 1] << ' ' << (int)rgb[i + 2] << '\n';
  }
  void write_png(std::ostream &out) const { rtx::write_png(out, rgb.data(), width, height); }
  void write_ppm_binary(std::ostream &out) const {
    out << "P6\n" << width << ' ' << height << "\n255\n";
    out.write((const char *)rgb.data(), (std::streamsize)rgb.size());
  }
};

// The drop-in for worker()/render<<<>>>: flatten -> rt_scene_upload -> rt_accel_build ->
// rt_render -> rt_resolve. Throws std::runtime_error with rt_last_error() on failure.
class renderer {
public:
  renderer(const render_options &o = render_options()) : opt(o) {
    rt_config cfg = {};
    cfg.device = o.device; cfg.profile = o.profile; cfg.flags = 0; cfg.seed = o.seed;
    int rc = rt_create(&ctx, &cfg);
    if (rc) throw std::runtime_error(rc == RT_ERR_NODEVICE ? "no CUDA device (there is no CPU fallback)" : "rt_create failed");
  }
  ~renderer() { if (ctx) rt_destroy(ctx); }
  renderer(const renderer &) = delete;
  renderer &operator=(const renderer &) = delete;

  static rt_scene_desc describe(const flat_scene &fs, const camera &cam, const render_options &o) {
    rt_scene_desc d = {};
    d.n_spheres = (int)fs.spheres.size(); d.spheres = fs.spheres.data();
    d.n_triangles = (int)fs.triangles.size(); d.triangles = fs.triangles.data();
    d.n_quads = (int)fs.quads.size(); d.quads = fs.quads.data();
    d.n_materials = (int)fs.materials.size(); d.materials = fs.materials.data();
    d.n_media = (int)fs.media.size(); d.media = fs.media.data();
    d.n_perlin = (int)fs.perlin_tables.size(); d.perlin = fs.perlin_tables.data();
    d.n_images = (int)fs.images.size(); d.images = fs.images.data();
    d.n_groups = (int)fs.groups.size(); d.groups = fs.groups.data();
    d.n_instances = (int)fs.instances.size(); d.instances = fs.instances.data();
    d.camera = cam.describe();
    for (int a = 0; a < 3; a++) d.background[a] = (float)o.background[a];
    d.sky_gradient = o.sky_gradient ? 1 : 0;
    d.t_min = (float)o.t_min; d.max_depth = o.max_depth; d.flags = o.flags;
    return d;
  }
  void set_scene(const hittable &world, const camera &cam) {
    flat = flat_scene();
    flat.instancing = opt.instancing;
    world.flatten(flat, transform());
    rt_scene_desc d = describe(flat, cam, opt);
    check(rt_scene_upload(ctx, &d));
    check(rt_accel_build(ctx, 1));
  }
  void render(int width, int height, int samples_per_pixel, int spp_begin = 0) {
    W = width; H = height;
    check(rt_render(ctx, width, height, spp_begin, samples_per_pixel));
  }
  image8 resolve() {
    image8 im;
    im.width = W; im.height = H;
    im.rgb.resize((size_t)W * H * 3);
    check(rt_resolve(ctx, nullptr, im.rgb.data()));
    return im;
  }
  // denoiser feature buffers [H][W][8]: albedo rgb, normal xyz, t, hit fraction (rt_render_aov)
  std::vector<float> render_aov(int width, int height, int samples_per_pixel = 1) {
    std::vector<float> f((size_t)width * height * 8);
    check(rt_render_aov(ctx, width, height, samples_per_pixel, f.data()));
    return f;
  }
  // accumulation frame (float4 per pixel: sum R, G, B and the sample count) — checkpoints, and the
  // host-side combine of a multi-GPU sample split
  std::vector<float> accum_download() {
    std::vector<float> f((size_t)W * H * 4);
    check(rt_accum_download(ctx, f.data(), f.size()));
    return f;
  }
  void accum_upload(int width, int height, const std::vector<float> &f) {
    W = width; H = height;
    check(rt_accum_upload(ctx, width, height, f.data(), f.size()));
  }
  // multi-GPU sample split inside ONE process: group() gives every renderer (one per device) its NCCL
  // communicator, reduce() - called concurrently from one host thread per renderer - sums the frames onto
  // renderer `root` of the group (rt_reduce: R,G,B through ncclReduce)
  static void group(const std::vector<renderer *> &rs) {
    std::vector<rt_ctx *> h;
    for (renderer *r : rs) h.push_back(r->ctx);
    if (!h.empty()) rs[0]->check(rt_comm_init_all(h.data(), (int)h.size()));
  }
  void reduce(int root = 0) {
    check(rt_reduce(ctx, W, H, nullptr, root, RT_REDUCE_UNIFORM_COUNT, nullptr));
    check(rt_sync(ctx));
  }
  // Transform-only update (opt.instancing): edit flat.instances[k].m (or use place_instance) and call this - only
  // the top level of the two-level BVH is rebuilt, the objects' trees and all primitive arrays stay on the device.
  void place_instance(int k, double angle_deg, const vec3 &offset) {
    const transform t = transform().then_translate(offset).then_rotate_y(angle_deg);
    const float m[12] = {(float)t.c, 0.f, (float)t.s, (float)t.offset[0], 0.f, 1.f, 0.f, (float)t.offset[1],
                         -(float)t.s, 0.f, (float)t.c, (float)t.offset[2]};
    for (int q = 0; q < 12; q++) flat.instances.at((size_t)k).m[q] = m[q];
  }
  void update_instances() { check(rt_instances_update(ctx, flat.instances.data(), (int)flat.instances.size())); }
  rt_stats_t stats() { rt_stats_t s; check(rt_stats(ctx, &s)); return s; }
  rt_ctx *handle() { return ctx; }
  flat_scene flat;

private:
  void check(int rc) { if (rc) throw std::runtime_error(std::string("b200rt: ") + rt_last_error(ctx)); }
  render_options opt;
  rt_ctx *ctx = nullptr;
  int W = 0, H = 0;
};

} // namespace rtx
#endif
