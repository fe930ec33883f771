// apps/scenes.h — the reference's scenes as DATA, built through its own scene-description
// surface (include/rtx compat headers: sphere, moving_sphere, xy/xz/yz_rect, box, translate,
// rotate_y, constant_medium, bvh_node, the materials and textures). The numbers are the
// reference's (file:line cited per scene); the code that assembles them is table-driven and
// shared between scenes. Where the reference draws random numbers the ORDER of the draws is kept,
// so that the Weekend scene reproduces tests/golden bit for bit under glibc's rand().
#ifndef APPS_SCENES_H
#define APPS_SCENES_H
#include <algorithm>

#include "aarect.h"
#include "box.h"
#include "bvh.h"
#include "camera.h"
#include "constant_medium.h"
#include "hittable_list.h"
#include "material.h"
#include "moving_sphere.h"
#include "sphere.h"
#include "triangle.h"

namespace scene_data {

// One axis-aligned rectangle: plane (0 = yz, 1 = xz, 2 = xy), in-plane ranges, offset, material slot.
struct rect_row { int plane; double a0, a1, b0, b1, k; int mat; };

inline hittable *make_rect(const rect_row &r, material *m) {
  if (r.plane == 0) return new yz_rect(r.a0, r.a1, r.b0, r.b1, r.k, m);
  if (r.plane == 1) return new xz_rect(r.a0, r.a1, r.b0, r.b1, r.k, m);
  return new xy_rect(r.a0, r.a1, r.b0, r.b1, r.k, m);
}

// translate(rotate_y(object, degrees), offset) — the only instance wrapper chain the reference uses
inline hittable *placed(hittable *h, double degrees, const vec3 &offset) {
  return new translate(new rotate_y(h, degrees), offset);
}

inline hittable *as_bvh(std::vector<hittable *> &objects) { // the vector must outlive the node's use
  return new bvh_node(objects.data(), 0, objects.size(), 0.0, 1.0);
}

// the three showpiece balls both random scenes end with (main.cpp:123-130, main.cu:189-194)
struct big_ball { double x; int kind; color c; double param; };
static const big_ball kBigBalls[3] = {{0, 2, color(0, 0, 0), 1.5}, {-4, 0, color(0.4, 0.2, 0.1), 0}, {4, 1, color(0.7, 0.6, 0.5), 0.0}};

} // namespace scene_data

// Weekend cover scene — rt_in_one_weekend/main.cpp:86-131. Draw order per grid cell (g++ evaluates
// constructor / operator arguments right to left): material choice, z offset, x offset; then for a
// diffuse ball the SECOND factor of the albedo product before the first.
inline hittable_list random_scene() {
  using namespace scene_data;
  hittable_list world;
  world.add(make_shared<sphere>(point3(0, -1000, 0), 1000, make_shared<lambertian>(color(0.5, 0.5, 0.5))));
  const point3 keep_clear(4, 0.2, 0);
  for (int a = -11; a < 11; a++)
    for (int b = -11; b < 11; b++) {
      const double pick = random_double();
      const double dz = 0.9 * random_double(), dx = 0.9 * random_double();
      const point3 at(a + dx, 0.2, b + dz);
      if (!((at - keep_clear).length() > 0.9)) continue;
      shared_ptr<material> m;
      if (pick < 0.8) {
        const color second = color::random(), first = color::random();
        m = make_shared<lambertian>(first * second);
      } else if (pick < 0.95) {
        const color tint = color::random(0.5, 1);
        m = make_shared<metal>(tint, random_double(0, 0.5));
      } else {
        m = make_shared<dielectric>(1.5);
      }
      world.add(make_shared<sphere>(at, 0.2, m));
    }
  for (const big_ball &bb : kBigBalls) {
    shared_ptr<material> m = bb.kind == 2   ? shared_ptr<material>(make_shared<dielectric>(bb.param))
                             : bb.kind == 0 ? shared_ptr<material>(make_shared<lambertian>(bb.c))
                                            : shared_ptr<material>(make_shared<metal>(bb.c, bb.param));
    world.add(make_shared<sphere>(point3(bb.x, 1, 0), 1.0, m));
  }
  return world;
}

// Motion-blur scene — rt_next_week/cuda/main.cu:153-198 in the raw-pointer style of the CUDA trees
// (cuRAND there, the host generator here: the sequence is not part of the contract). d_list must hold
// 22*22 + 4 pointers, as the reference's create_world expects.
inline hittable *next_week_random_scene(hittable **d_list) {
  using namespace scene_data;
  auto rnd = []() { return (float)random_double(); };
  int n = 0;
  d_list[n++] = new sphere(vec3(0, -1000.0, -1), 1000,
                           new lambertian(new checker_texture(color(0.2, 0.3, 0.1), color(0.9, 0.9, 0.9))));
  for (int a = -11; a < 11; a++)
    for (int b = -11; b < 11; b++) {
      const float pick = rnd();
      const float x = a + rnd();
      const float z = b + rnd();
      const vec3 at(x, 0.2, z);
      if (pick < 0.8f) { // moving diffuse ball: squared-uniform albedo, centre rising by up to 0.5
        const vec3 to = at + vec3(0, rnd() * 0.5f, 0);
        float rgb[3];
        for (float &ch : rgb) { const float u = rnd(), v = rnd(); ch = u * v; }
        d_list[n++] = new moving_sphere(at, to, 0.0, 1.0, 0.2, new lambertian(vec3(rgb[0], rgb[1], rgb[2])));
      } else if (pick < 0.95f) {
        float rgb[3];
        for (float &ch : rgb) ch = 0.5f * (1.0f + rnd());
        d_list[n++] = new sphere(at, 0.2, new metal(vec3(rgb[0], rgb[1], rgb[2]), 0.5f * rnd()));
      } else {
        d_list[n++] = new sphere(at, 0.2, new dielectric(1.5));
      }
    }
  for (const big_ball &bb : kBigBalls) {
    material *m = bb.kind == 2 ? (material *)new dielectric(bb.param)
                  : bb.kind == 0 ? (material *)new lambertian(bb.c) : (material *)new metal(bb.c, bb.param);
    d_list[n++] = new sphere(vec3(bb.x, 1, 0), 1.0, m);
  }
  return new bvh_node(d_list, 0, n, 0.0f, 1.0f);
}

// Mesh in a lit, mirrored room — triangles/cuda/obj_render.cu:384-524: two coloured sphere lights, two
// strip lights, red floor / ceiling, blue side walls, a pale back wall, three mirrors, and the mesh
// scaled 2.5x, turned 30 degrees and lifted 1.5.
inline hittable *obj_model(const std::string &obj_path) {
  using namespace scene_data;
  auto rgb256 = [](double r, double g, double b) { return color(r / 256.0, g / 256.0, b / 256.0); };
  material *mats[] = {new diffuse_light(color(20, 20, 20) * 0.25),          // 0 strip lights
                      new lambertian(rgb256(253.0f, 252.0f, 220.0f)),       // 1 back wall
                      new lambertian(rgb256(240.0f, 113.0f, 103.0f)),       // 2 floor / ceiling
                      new lambertian(rgb256(0, 129.0f, 167.0f)),            // 3 side walls
                      nullptr};                                             // 4 mirrors: one material each
  const double top = 4 + 1;
  const rect_row rows[] = {{1, -4, 4, 3, 4, top - 0.01, 0},   {1, -4, 4, 2, 3, -4 + 0.01, 0},
                           {2, -4, 4, -4, top, -4, 1},        {2, -3, 3, -4, top, -3.999, 4},
                           {1, -40, 40, -40, 40, -4, 2},      {1, -40, 40, -40, 40, top, 2},
                           {0, -4, top, -4, 4, -4, 3},        {0, -4, top, -4, 4, 4, 3},
                           {0, -1, 3 + 1, -4, 4, -3.999, 4},  {0, -1, 3 + 1 - 0.001, -4, 4, 3.999, 4}};
  auto *room = new hittable_list();
  const color bulbs[2] = {rgb256(255.0f, 59.0f, 148.0f) * 2, rgb256(166.0f, 253.0f, 41.0f) * 2};
  for (int k = 0; k < 2; k++)
    room->add(borrow<hittable>(new sphere(point3(k ? 1 : -1, 3.69 + 1, -2.5), 0.3, new diffuse_light(bulbs[k]))));
  for (const rect_row &r : rows)
    room->add(borrow<hittable>(make_rect(r, r.mat == 4 ? new metal(color(0.8, 0.8, 0.9), 0.0) : mats[r.mat])));
  std::vector<shared_ptr<hittable>> mesh;
  read_triangles(obj_path, mesh, make_shared<metal>(rgb256(255.0f, 215.0f, 0.0f), 0.5), 2.5);
  for (auto &t : mesh) room->add(make_shared<translate>(make_shared<rotate_y>(t, 30), vec3(0, 1.5, 0)));
  return new bvh_node(*room, 0.0, 1.0);
}

namespace scene_data {
// The Cornell room shared by scenes 6 and 7 (rt_next_week/cuda/main.cu:252-266, 283-297): material
// slots 0 green, 1 red, 2 light, 3 white; the two blocks are 165-wide boxes turned by +15 / -18 degrees.
struct cornell_parts { std::vector<hittable *> walls; material *white; hittable *block[2]; };
inline cornell_parts cornell_room() {
  material *m[4] = {new lambertian(color(.12, .45, .15)), new lambertian(color(.65, .05, .05)),
                    new diffuse_light(color(15, 15, 15)), new lambertian(color(.73, .73, .73))};
  const rect_row rows[] = {{0, 0, 555, 0, 555, 555, 0}, {0, 0, 555, 0, 555, 0, 1}, {1, 213, 343, 227, 332, 554, 2},
                           {1, 0, 555, 0, 555, 0, 3},   {1, 0, 555, 0, 555, 555, 3}, {2, 0, 555, 0, 555, 555, 3}};
  cornell_parts c;
  for (const rect_row &r : rows) c.walls.push_back(make_rect(r, m[r.mat]));
  c.white = m[3];
  const double height[2] = {330, 165}, turn[2] = {15, -18};
  const vec3 where[2] = {vec3(265, 0, 295), vec3(130, 0, 65)};
  for (int k = 0; k < 2; k++)
    c.block[k] = placed(new box(point3(0, 0, 0), point3(165, height[k], 165), c.white), turn[k], where[k]);
  return c;
}
} // namespace scene_data

// rt_next_week/cuda/main.cu:252-281
inline hittable *cornell_box() {
  static std::vector<hittable *> objects;
  scene_data::cornell_parts c = scene_data::cornell_room();
  objects = c.walls;
  objects.push_back(c.block[0]);
  objects.push_back(c.block[1]);
  return scene_data::as_bvh(objects);
}

// rt_next_week/cuda/main.cu:283-310: the blocks become black and white smoke of density 0.01
inline hittable *cornell_smoke() {
  static std::vector<hittable *> objects;
  scene_data::cornell_parts c = scene_data::cornell_room();
  objects = c.walls;
  objects.push_back(new constant_medium(c.block[0], 0.01, color(0, 0, 0)));
  objects.push_back(new constant_medium(c.block[1], 0.01, color(1, 1, 1)));
  return scene_data::as_bvh(objects);
}

// Stand-in for the reference's earthmap.jpeg (a binary asset that is not redistributed): an
// equirectangular map with oceans, continents and polar caps. [h][w][3] bytes, top row first.
inline std::vector<unsigned char> procedural_earth(int w = 512, int h = 256) {
  std::vector<unsigned char> img((size_t)w * h * 3);
  const double waves[6][4] = {{1.5, 0.7, -0.4, 0.3}, {-0.8, 1.9, 1.1, 1.7}, {2.6, -1.2, 2.2, 4.1},
                              {-3.1, 2.4, -2.9, 0.9}, {5.3, 4.1, -3.7, 2.6}, {-6.2, -5.5, 6.8, 5.2}};
  for (int j = 0; j < h; j++)
    for (int i = 0; i < w; i++) {
      const double lon = (i + 0.5) / w * 2 * pi, lat = (0.5 - (j + 0.5) / h) * pi;
      const double x = std::cos(lat) * std::cos(lon), y = std::sin(lat), z = std::cos(lat) * std::sin(lon);
      double f = 0;
      for (int k = 0; k < 6; k++)
        f += std::sin(waves[k][0] * x + waves[k][1] * y + waves[k][2] * z + waves[k][3]) / (1 + k / 2);
      double c[3] = {0.05, 0.15, 0.45};
      if (f > 0.35) {
        const double t = std::min(1.0, (f - 0.35) / 1.5);
        c[0] = (1 - t) * 0.15 + t * 0.45; c[1] = (1 - t) * 0.45 + t * 0.35; c[2] = (1 - t) * 0.12 + t * 0.2;
      }
      if (std::fabs(lat) > 1.25) { c[0] = 0.92; c[1] = 0.94; c[2] = 0.96; }
      for (int a = 0; a < 3; a++) img[((size_t)j * w + i) * 3 + a] = (unsigned char)(c[a] * 255.0 + 0.5);
    }
  return img;
}

// rt_next_week/cuda/main.cu:212-250: noise_texture(4) on a ground sphere and a ball; scene 5 adds a
// rect light and a sphere light
inline hittable *perlin_pair(bool with_lights) {
  static std::vector<hittable *> objects;
  objects.clear();
  auto marble = new noise_texture(4);
  objects.push_back(new sphere(point3(0, -1000, 0), 1000, new lambertian(marble)));
  objects.push_back(new sphere(point3(0, 2, 0), 2, new lambertian(marble)));
  if (with_lights) {
    objects.push_back(new xy_rect(3, 5, 1, 2, -2, new diffuse_light(color(4, 4, 4))));
    objects.push_back(new sphere(point3(0, 6, 0), 1.5, new diffuse_light(color(6, 4, 4))));
  }
  return scene_data::as_bvh(objects);
}
inline hittable *two_perlin_spheres() { return perlin_pair(false); }
inline hittable *simple_light() { return perlin_pair(true); }

// rt_next_week/cuda/main.cu:224-232: one image-textured globe
inline hittable *earth(unsigned char *data, int w, int h) {
  static std::vector<hittable *> objects;
  objects.assign(1, new sphere(point3(0, 0, 0), 2, new lambertian(new image_texture(data, w, h))));
  return scene_data::as_bvh(objects);
}

// The rt_next_week tree's default scene — main.cu:312-383: a 20 x 20 field of ground boxes of random
// height (1..101), an area light, a motion-blurred ball, glass and brushed-metal balls, a blue
// "subsurface" ball (glass boundary + dense medium), thin fog over everything, the globe, a marble
// ball, and 1000 small white balls in a cube turned by 15 degrees.
inline hittable *rt_next_week_final_scene(unsigned char *data, int w, int h) {
  using namespace scene_data;
  static std::vector<hittable *> objects;
  static std::vector<hittable *> cluster;
  objects.clear();
  cluster.clear();
  auto ground = new lambertian(color(0.48, 0.83, 0.53));
  const int side = 20;
  const float pitch = 100.0f;
  for (int i = 0; i < side; i++)
    for (int j = 0; j < side; j++) {
      const float x = -1000.0f + i * pitch, z = -1000.0f + j * pitch;
      const float top = (float)random_double(1, 101);
      objects.push_back(new box(point3(x, 0.0f, z), point3(x + pitch, top, z + pitch), ground));
    }
  objects.push_back(new xz_rect(123, 423, 147, 412, 554, new diffuse_light(color(7, 7, 7))));
  const point3 from(400, 400, 200);
  objects.push_back(new moving_sphere(from, from + vec3(30, 0, 0), 0, 1, 50, new lambertian(color(0.7, 0.3, 0.1))));
  objects.push_back(new sphere(point3(260, 150, 45), 50, new dielectric(1.5)));
  objects.push_back(new sphere(point3(0, 150, 145), 50, new metal(color(0.8, 0.8, 0.9), 1.0)));
  hittable *skin = new sphere(point3(360, 150, 145), 70, new dielectric(1.5));
  objects.push_back(skin);
  objects.push_back(new constant_medium(skin, 0.2, color(0.2, 0.4, 0.9)));
  objects.push_back(new constant_medium(new sphere(point3(0, 0, 0), 5000, new dielectric(1.5)), 0.0001, color(1, 1, 1)));
  objects.push_back(new sphere(point3(400, 200, 400), 100, new lambertian(new image_texture(data, w, h))));
  objects.push_back(new sphere(point3(220, 280, 300), 80, new lambertian(new noise_texture(0.1))));
  auto white = new lambertian(color(.73, .73, .73));
  for (int k = 0; k < 1000; k++) {
    const double x = random_double(0, 165), y = random_double(0, 165), z = random_double(0, 165);
    cluster.push_back(new sphere(point3(x, y, z), 10, white));
  }
  objects.push_back(placed(as_bvh(cluster), 15, vec3(-100, 270, 395)));
  return as_bvh(objects);
}
#endif
