// scene_file.h — text scene files for the host programs, so that a scene no longer has to be
// compiled into a main() the way every renderer of the reference does it
// (rt_in_one_weekend/main.cpp:86-131, rt_next_week/cuda/main.cu:153-383,
// triangles/cuda/obj_render.cu:384-524). One statement per line, `#` starts a comment, the
// objects are the reference's own classes (rtx.h), constructed with the same arguments:
//
//   profile weekend | final_cu | next_week
//   image W H SPP                                   default frame of the host program
//   camera lookfrom x y z lookat x y z [vup x y z] vfov deg [aperture a] [focus d] [shutter t0 t1]
//   sky | background r g b                          miss shader (gradient or constant)
//   tmin t | depth n | flags [flip_normals] [depth_background] [reference_medium] [instancing] | seed s
//                 (instancing: objects under an XFORM become instances of a two-level BVH instead of being baked)
//   material NAME lambertian r g b | lambertian checker r g b r g b | lambertian noise scale [table-seed]
//                 | lambertian image FILE.ppm | metal r g b fuzz | dielectric index | light r g b
//   sphere cx cy cz r MAT [XFORM]
//   moving_sphere x0 y0 z0 x1 y1 z1 t0 t1 r MAT [XFORM]
//   xy_rect x0 x1 y0 y1 k MAT | xz_rect x0 x1 z0 z1 k MAT | yz_rect y0 y1 z0 z1 k MAT   [XFORM]
//   box x0 y0 z0 x1 y1 z1 MAT [XFORM]
//   triangle x0 y0 z0 x1 y1 z1 x2 y2 z2 MAT [XFORM]      (geometric normal)
//   obj FILE.obj MAT [scale s] [XFORM]
//   medium sphere cx cy cz r DENSITY r g b | medium box x0 y0 z0 x1 y1 z1 DENSITY r g b [XFORM]
//   XFORM = [rotate_y deg] [translate x y z]        applied as translate(rotate_y(object))
// File names are relative to the scene file. Errors throw std::runtime_error("file:line: what").
#ifndef RTX_SCENE_FILE_H
#define RTX_SCENE_FILE_H

#include <map>

#include "rtx.h"

namespace rtx {

struct scene_file {
  hittable_list world;
  render_options opt;
  int width = 600, height = 400, spp = 100;
  // camera statement (the camera needs the final aspect ratio, so it is built on demand)
  point3 lookfrom = point3(0, 0, 1), lookat = point3(0, 0, 0);
  vec3 vup = vec3(0, 1, 0);
  double vfov = 40, aperture = 0, focus = -1, time0 = 0, time1 = 0;
  std::map<std::string, shared_ptr<material>> materials;

  camera make_camera(int w, int h) const {
    const double fd = focus > 0 ? focus : (lookfrom - lookat).length();
    return camera(lookfrom, lookat, vup, vfov, double(w) / h, aperture, fd, time0, time1,
                  opt.profile != RT_PROFILE_WEEKEND_CPU);
  }
};

namespace detail {
struct line_reader {
  std::string file;
  int line_no = 0;
  std::vector<std::string> tok;
  size_t pos = 0;
  [[noreturn]] void fail(const std::string &what) const {
    throw std::runtime_error(file + ":" + std::to_string(line_no) + ": " + what);
  }
  bool more() const { return pos < tok.size(); }
  const std::string &word(const char *what) {
    if (!more()) fail(std::string("missing ") + what);
    return tok[pos++];
  }
  bool accept(const char *kw) {
    if (more() && tok[pos] == kw) { pos++; return true; }
    return false;
  }
  double num(const char *what) {
    const std::string &w = word(what);
    char *end = nullptr;
    double v = std::strtod(w.c_str(), &end);
    if (end == w.c_str() || *end) fail(std::string("expected a number for ") + what + ", got '" + w + "'");
    return v;
  }
  vec3 v3(const char *what) { double x = num(what), y = num(what), z = num(what); return vec3(x, y, z); }
  void done() { if (more()) fail("unexpected '" + tok[pos] + "'"); }
};
} // namespace detail

inline scene_file load_scene_file(const std::string &path) {
  std::ifstream in(path);
  if (!in.is_open()) throw std::runtime_error("cannot open scene file " + path);
  const size_t slash = path.find_last_of('/');
  const std::string dir = slash == std::string::npos ? "" : path.substr(0, slash + 1);
  auto resolve = [&](const std::string &f) { return (!f.empty() && f[0] == '/') ? f : dir + f; };
  scene_file sf;
  sf.opt.sky_gradient = true;
  detail::line_reader L;
  L.file = path;
  std::string line;
  bool have_camera = false;
  auto mat = [&](const std::string &name) {
    auto it = sf.materials.find(name);
    if (it == sf.materials.end()) L.fail("unknown material '" + name + "'");
    return it->second;
  };
  // [rotate_y deg] [translate x y z] around an object
  auto xform = [&](shared_ptr<hittable> h) {
    if (L.accept("rotate_y")) h = make_shared<rotate_y>(h, L.num("angle"));
    if (L.accept("translate")) h = make_shared<translate>(h, L.v3("offset"));
    return h;
  };
  while (std::getline(in, line)) {
    L.line_no++;
    const size_t hash = line.find('#');
    if (hash != std::string::npos) line.erase(hash);
    std::istringstream ls(line);
    L.tok.clear();
    L.pos = 0;
    for (std::string w; ls >> w;) L.tok.push_back(w);
    if (L.tok.empty()) continue;
    const std::string kw = L.word("statement");
    if (kw == "profile") {
      const std::string &p = L.word("profile name");
      if (p == "weekend") sf.opt.profile = RT_PROFILE_WEEKEND_CPU;
      else if (p == "final_cu") sf.opt.profile = RT_PROFILE_FINAL_CU;
      else if (p == "next_week") sf.opt.profile = RT_PROFILE_NEXT_WEEK;
      else L.fail("unknown profile '" + p + "'");
    } else if (kw == "image") {
      sf.width = (int)L.num("width"); sf.height = (int)L.num("height"); sf.spp = (int)L.num("samples");
      if (sf.width < 2 || sf.height < 2 || sf.spp < 1) L.fail("image needs width, height >= 2 and samples >= 1");
    } else if (kw == "camera") {
      bool have_from = false, have_at = false, have_fov = false;
      while (L.more()) {
        if (L.accept("lookfrom")) { sf.lookfrom = L.v3("lookfrom"); have_from = true; }
        else if (L.accept("lookat")) { sf.lookat = L.v3("lookat"); have_at = true; }
        else if (L.accept("vup")) sf.vup = L.v3("vup");
        else if (L.accept("vfov")) { sf.vfov = L.num("vfov"); have_fov = true; }
        else if (L.accept("aperture")) sf.aperture = L.num("aperture");
        else if (L.accept("focus")) sf.focus = L.num("focus");
        else if (L.accept("shutter")) { sf.time0 = L.num("time0"); sf.time1 = L.num("time1"); }
        else L.fail("unknown camera field '" + L.tok[L.pos] + "'");
      }
      if (!have_from || !have_at || !have_fov) L.fail("camera needs lookfrom, lookat and vfov");
      have_camera = true;
    } else if (kw == "sky") {
      sf.opt.sky_gradient = true;
    } else if (kw == "background") {
      sf.opt.background = L.v3("background");
      sf.opt.sky_gradient = false;
    } else if (kw == "tmin") {
      sf.opt.t_min = L.num("t_min");
    } else if (kw == "depth") {
      sf.opt.max_depth = (int)L.num("depth");
    } else if (kw == "seed") {
      sf.opt.seed = (uint64_t)L.num("seed");
    } else if (kw == "flags") {
      while (L.more()) {
        if (L.accept("flip_normals")) sf.opt.flags |= RT_FLAG_FLIP_NORMALS;
        else if (L.accept("depth_background")) sf.opt.flags |= RT_FLAG_DEPTH_BACKGROUND;
        else if (L.accept("reference_medium")) sf.opt.flags |= RT_FLAG_REFERENCE_MEDIUM;
        else if (L.accept("instancing")) sf.opt.instancing = true;
        else L.fail("unknown flag '" + L.tok[L.pos] + "'");
      }
    } else if (kw == "material") {
      const std::string name = L.word("material name");
      const std::string type = L.word("material type");
      shared_ptr<material> m;
      if (type == "lambertian") {
        if (L.accept("checker")) { color a = L.v3("even"), b = L.v3("odd"); m = make_shared<lambertian>(make_shared<checker_texture>(a, b)); }
        else if (L.accept("noise")) {
          const double scale = L.num("scale");
          const uint64_t seed = L.more() ? (uint64_t)L.num("table seed") : 1984 + sf.materials.size();
          m = make_shared<lambertian>(noise_texture::seeded(scale, seed));
        }
        else if (L.accept("image")) {
          const std::string f = resolve(L.word("image file"));
          auto tex = make_shared<image_texture>(f.c_str());
          if (tex->data.empty()) L.fail("cannot read image '" + f + "' (binary or ASCII PPM expected)");
          m = make_shared<lambertian>(tex);
        } else m = make_shared<lambertian>(L.v3("albedo"));
      } else if (type == "metal") { color a = L.v3("albedo"); m = make_shared<metal>(a, L.num("fuzz")); }
      else if (type == "dielectric") m = make_shared<dielectric>(L.num("index"));
      else if (type == "light") m = make_shared<diffuse_light>(L.v3("emission"));
      else L.fail("unknown material type '" + type + "'");
      sf.materials[name] = m;
    } else if (kw == "sphere") {
      point3 c = L.v3("center");
      double r = L.num("radius");
      sf.world.add(xform(make_shared<sphere>(c, r, mat(L.word("material")))));
    } else if (kw == "moving_sphere") {
      point3 c0 = L.v3("center0"), c1 = L.v3("center1");
      double t0 = L.num("time0"), t1 = L.num("time1"), r = L.num("radius");
      sf.world.add(xform(make_shared<moving_sphere>(c0, c1, t0, t1, r, mat(L.word("material")))));
    } else if (kw == "xy_rect" || kw == "xz_rect" || kw == "yz_rect") {
      double a0 = L.num("a0"), a1 = L.num("a1"), b0 = L.num("b0"), b1 = L.num("b1"), k = L.num("k");
      auto m = mat(L.word("material"));
      shared_ptr<hittable> h;
      if (kw == "xy_rect") h = make_shared<xy_rect>(a0, a1, b0, b1, k, m);
      else if (kw == "xz_rect") h = make_shared<xz_rect>(a0, a1, b0, b1, k, m);
      else h = make_shared<yz_rect>(a0, a1, b0, b1, k, m);
      sf.world.add(xform(h));
    } else if (kw == "box") {
      point3 p0 = L.v3("min"), p1 = L.v3("max");
      auto m = mat(L.word("material"));
      sf.world.add(xform(make_shared<box>(p0, p1, m)));
    } else if (kw == "triangle") {
      vec3 a = L.v3("v0"), b = L.v3("v1"), c = L.v3("v2");
      auto m = mat(L.word("material"));
      vec3 n = -cross(b - a, c - a); // vertex-normal convention of triangle.h:42-44: the face normal opposes their mean
      sf.world.add(xform(make_shared<triangle>(a, b, c, n, n, n, m)));
    } else if (kw == "obj") {
      const std::string f = resolve(L.word("obj file"));
      auto m = mat(L.word("material"));
      double scale = L.accept("scale") ? L.num("scale") : 1.0;
      std::vector<shared_ptr<hittable>> tris;
      try { read_triangles(f, tris, m, scale); } catch (const std::exception &e) { L.fail(e.what()); }
      // one transform statement for the whole mesh
      double angle = 0;
      vec3 offset(0, 0, 0);
      const bool rot = L.accept("rotate_y");
      if (rot) angle = L.num("angle");
      const bool tr = L.accept("translate");
      if (tr) offset = L.v3("offset");
      for (auto &t : tris) {
        shared_ptr<hittable> h = t;
        if (rot) h = make_shared<rotate_y>(h, angle);
        if (tr) h = make_shared<translate>(h, offset);
        sf.world.add(h);
      }
    } else if (kw == "medium") {
      const std::string shape = L.word("medium shape");
      shared_ptr<hittable> b;
      auto dummy = make_shared<dielectric>(1.5); // boundary material is irrelevant (main.cu:360)
      if (shape == "sphere") { point3 c = L.v3("center"); b = make_shared<sphere>(c, L.num("radius"), dummy); }
      else if (shape == "box") { point3 p0 = L.v3("min"), p1 = L.v3("max"); b = make_shared<box>(p0, p1, dummy); }
      else L.fail("medium shape must be sphere or box");
      double density = L.num("density");
      if (!(density > 0)) L.fail("medium density must be positive");
      color c = L.v3("albedo");
      sf.world.add(make_shared<constant_medium>(xform(b), density, c));
    } else {
      L.fail("unknown statement '" + kw + "'");
    }
    L.done();
  }
  if (!have_camera) throw std::runtime_error(path + ": no camera statement");
  return sf;
}

} // namespace rtx
#endif
