// apps/scenes.h — the reference's scenes written against its own scene-description
// surface (include/rtx compat headers): this is existing-style scene code, compiled
// unchanged in spirit, that now feeds the B200 core.
#ifndef APPS_SCENES_H
#define APPS_SCENES_H
#include "camera.h"
#include "hittable_list.h"
#include "material.h"
#include "sphere.h"
#include "moving_sphere.h"
#include "aarect.h"
#include "triangle.h"
#include "bvh.h"
#include "box.h"

// rt_in_one_weekend/main.cpp:86-131, statement for statement. g++ evaluates the
// arguments of `point3 center(a + 0.9*random_double(), 0.2, b + 0.9*random_double())`
// and of `color::random() * color::random()` right to left; the order is spelled out so
// that any compiler reproduces the reference's 487-sphere scene (tests/golden).
inline hittable_list random_scene() {
  hittable_list world;
  auto ground_material = make_shared<lambertian>(color(0.5, 0.5, 0.5));
  world.add(make_shared<sphere>(point3(0, -1000, 0), 1000, ground_material));
  for (int a = -11; a < 11; a++) {
    for (int b = -11; b < 11; b++) {
      auto choose_mat = random_double();
      double cz = b + 0.9 * random_double();
      double cx = a + 0.9 * random_double();
      point3 center(cx, 0.2, cz);
      if ((center - point3(4, 0.2, 0)).length() > 0.9) {
        shared_ptr<material> sphere_material;
        if (choose_mat < 0.8) {
          color second = color::random();
          color first = color::random();
          auto albedo = first * second;
          sphere_material = make_shared<lambertian>(albedo);
          world.add(make_shared<sphere>(center, 0.2, sphere_material));
        } else if (choose_mat < 0.95) {
          auto albedo = color::random(0.5, 1);
          auto fuzz = random_double(0, 0.5);
          sphere_material = make_shared<metal>(albedo, fuzz);
          world.add(make_shared<sphere>(center, 0.2, sphere_material));
        } else {
          sphere_material = make_shared<dielectric>(1.5);
          world.add(make_shared<sphere>(center, 0.2, sphere_material));
        }
      }
    }
  }
  auto material1 = make_shared<dielectric>(1.5);
  world.add(make_shared<sphere>(point3(0, 1, 0), 1.0, material1));
  auto material2 = make_shared<lambertian>(color(0.4, 0.2, 0.1));
  world.add(make_shared<sphere>(point3(-4, 1, 0), 1.0, material2));
  auto material3 = make_shared<metal>(color(0.7, 0.6, 0.5), 0.0);
  world.add(make_shared<sphere>(color(4, 1, 0), 1.0, material3));
  return world;
}

// rt_next_week/cuda/main.cu:153-198 in its raw-pointer style (`new`, RND); RND is
// cuRAND there, glibc here (the sequence is not part of the contract).
#define RND ((float)random_double())
inline hittable *next_week_random_scene(hittable **d_list) {
  auto checker = new checker_texture(color(0.2, 0.3, 0.1), color(0.9, 0.9, 0.9));
  d_list[0] = new sphere(vec3(0, -1000.0, -1), 1000, new lambertian(checker));
  int i = 1;
  for (int a = -11; a < 11; a++) {
    for (int b = -11; b < 11; b++) {
      float choose_mat = RND;
      float cx = a + RND, cz = b + RND;
      vec3 center(cx, 0.2, cz);
      if (choose_mat < 0.8f) {
        vec3 center2 = center + vec3(0, RND * 0.5f, 0);
        float r0 = RND * RND, r1 = RND * RND, r2 = RND * RND;
        d_list[i++] = new moving_sphere(center, center2, 0.0, 1.0, 0.2, new lambertian(vec3(r0, r1, r2)));
      } else if (choose_mat < 0.95f) {
        float r0 = 0.5f * (1.0f + RND), r1 = 0.5f * (1.0f + RND), r2 = 0.5f * (1.0f + RND);
        d_list[i++] = new sphere(center, 0.2, new metal(vec3(r0, r1, r2), 0.5f * RND));
      } else {
        d_list[i++] = new sphere(center, 0.2, new dielectric(1.5));
      }
    }
  }
  d_list[i++] = new sphere(vec3(0, 1, 0), 1.0, new dielectric(1.5));
  d_list[i++] = new sphere(vec3(-4, 1, 0), 1.0, new lambertian(vec3(0.4, 0.2, 0.1)));
  d_list[i++] = new sphere(vec3(4, 1, 0), 1.0, new metal(vec3(0.7, 0.6, 0.5), 0.0));
  return new bvh_node(d_list, 0, 22 * 22 + 1 + 3, 0.0f, 1.0f);
}

// triangles/cuda/obj_render.cu:384-524 (obj_model): the mesh inside a lit, mirrored room
inline hittable *obj_model(const std::string &obj_path) {
  std::vector<shared_ptr<hittable>> tris;
  auto blue_1 = new lambertian(color(0, 129.0f / 256.0, 167.0f / 256.0));
  auto red_1 = new lambertian(color(240.0f / 256.0, 113.0f / 256.0, 103.0f / 256.0));
  auto yellow_1 = new lambertian(color(253.0f / 256.0, 252.0f / 256.0, 220.0f / 256.0));
  auto gold = make_shared<metal>(color(255.0f / 256.0, 215.0f / 256.0, 0.0f / 256.0), 0.5);
  auto light = new diffuse_light(color(20, 20, 20) * 0.25);
  auto *ret = new hittable_list();
  ret->add(borrow<hittable>(new sphere(point3(-1, 3.69 + 1, -2.5), 0.3,
                                       new diffuse_light(color(255.0f / 256.0, 59.0f / 256.0, 148.0f / 256.0) * 2))));
  ret->add(borrow<hittable>(new sphere(point3(1, 3.69 + 1, -2.5), 0.3,
                                       new diffuse_light(color(166.0f / 256.0, 253.0f / 256.0, 41.0f / 256.0) * 2))));
  ret->add(borrow<hittable>(new xz_rect(-4, 4, 3, 4, 4 + 1 - 0.01, light)));
  ret->add(borrow<hittable>(new xz_rect(-4, 4, 2, 3, -4 + 0.01, light)));
  ret->add(borrow<hittable>(new xy_rect(-4, 4, -4, 4 + 1, -4, yellow_1)));
  ret->add(borrow<hittable>(new xy_rect(-3, 3, -4, 4 + 1, -3.999, new metal(color(0.8, 0.8, 0.9), 0.0))));
  ret->add(borrow<hittable>(new xz_rect(-40, 40, -40, 40, -4, red_1)));
  ret->add(borrow<hittable>(new xz_rect(-40, 40, -40, 40, 4 + 1, red_1)));
  ret->add(borrow<hittable>(new yz_rect(-4, 4 + 1, -4, 4, -4, blue_1)));
  ret->add(borrow<hittable>(new yz_rect(-4, 4 + 1, -4, 4, 4, blue_1)));
  ret->add(borrow<hittable>(new yz_rect(-1, 3 + 1, -4, 4, -3.999, new metal(color(0.8, 0.8, 0.9), 0.0))));
  ret->add(borrow<hittable>(new yz_rect(-1, 3 + 1 - 0.001, -4, 4, 3.999, new metal(color(0.8, 0.8, 0.9), 0.0))));
  read_triangles(obj_path, tris, gold, 2.5);
  for (auto &t : tris) ret->add(make_shared<translate>(make_shared<rotate_y>(t, 30), vec3(0, 1.5, 0)));
  return new bvh_node(*ret, 0.0, 1.0);
}

// rt_next_week/cuda/main.cu:252-281 (cornell_box), in its raw-pointer style
inline hittable *cornell_box() {
  static hittable *ret[8];
  auto red = new lambertian(color(.65, .05, .05));
  auto white = new lambertian(color(.73, .73, .73));
  auto green = new lambertian(color(.12, .45, .15));
  auto light = new diffuse_light(color(15, 15, 15));
  ret[0] = new yz_rect(0, 555, 0, 555, 555, green);
  ret[1] = new yz_rect(0, 555, 0, 555, 0, red);
  ret[2] = new xz_rect(213, 343, 227, 332, 554, light);
  ret[3] = new xz_rect(0, 555, 0, 555, 0, white);
  ret[4] = new xz_rect(0, 555, 0, 555, 555, white);
  ret[5] = new xy_rect(0, 555, 0, 555, 555, white);
  hittable *box1 = new box(point3(0, 0, 0), point3(165, 330, 165), white);
  box1 = new rotate_y(box1, 15);
  box1 = new translate(box1, vec3(265, 0, 295));
  hittable *box2 = new box(point3(0, 0, 0), point3(165, 165, 165), white);
  box2 = new rotate_y(box2, -18);
  box2 = new translate(box2, vec3(130, 0, 65));
  ret[6] = box1;
  ret[7] = box2;
  return new bvh_node(ret, 0, 8, 0.0f, 1.0f);
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

// rt_next_week/cuda/main.cu:212-222
inline hittable *two_perlin_spheres() {
  auto perlin_texture = new noise_texture(4);
  static hittable *ret[2];
  ret[0] = new sphere(point3(0, -1000, 0), 1000, new lambertian(perlin_texture));
  ret[1] = new sphere(point3(0, 2, 0), 2, new lambertian(perlin_texture));
  return new bvh_node(ret, 0, 2, 0.0f, 1.0f);
}

// rt_next_week/cuda/main.cu:224-232
inline hittable *earth(unsigned char *data, int w, int h) {
  auto earth_texture = new image_texture(data, w, h);
  auto earth_surface = new lambertian(earth_texture);
  static hittable *ret[1];
  ret[0] = new sphere(point3(0, 0, 0), 2, earth_surface);
  return new bvh_node(ret, 0, 1, 0.0f, 1.0f);
}

// rt_next_week/cuda/main.cu:234-250
inline hittable *simple_light() {
  auto perlin_texture = new noise_texture(4);
  static hittable *ret[4];
  ret[0] = new sphere(point3(0, -1000, 0), 1000, new lambertian(perlin_texture));
  ret[1] = new sphere(point3(0, 2, 0), 2, new lambertian(perlin_texture));
  auto diff_light = new diffuse_light(color(4, 4, 4));
  ret[2] = new xy_rect(3, 5, 1, 2, -2, diff_light);
  auto diff_light2 = new diffuse_light(color(6, 4, 4));
  ret[3] = new sphere(point3(0, 6, 0), 1.5, diff_light2);
  return new bvh_node(ret, 0, 4, 0.0f, 1.0f);
}

// rt_next_week/cuda/main.cu:283-310
inline hittable *cornell_smoke() {
  static hittable *ret[8];
  auto red = new lambertian(color(.65, .05, .05));
  auto white = new lambertian(color(.73, .73, .73));
  auto green = new lambertian(color(.12, .45, .15));
  auto light = new diffuse_light(color(15, 15, 15));
  ret[0] = new yz_rect(0, 555, 0, 555, 555, green);
  ret[1] = new yz_rect(0, 555, 0, 555, 0, red);
  ret[2] = new xz_rect(213, 343, 227, 332, 554, light);
  ret[3] = new xz_rect(0, 555, 0, 555, 0, white);
  ret[4] = new xz_rect(0, 555, 0, 555, 555, white);
  ret[5] = new xy_rect(0, 555, 0, 555, 555, white);
  hittable *box1 = new box(point3(0, 0, 0), point3(165, 330, 165), white);
  box1 = new rotate_y(box1, 15);
  box1 = new translate(box1, vec3(265, 0, 295));
  box1 = new constant_medium(box1, 0.01, color(0, 0, 0));
  hittable *box2 = new box(point3(0, 0, 0), point3(165, 165, 165), white);
  box2 = new rotate_y(box2, -18);
  box2 = new translate(box2, vec3(130, 0, 65));
  box2 = new constant_medium(box2, 0.01, color(1, 1, 1));
  ret[6] = box1;
  ret[7] = box2;
  return new bvh_node(ret, 0, 8, 0.0f, 1.0f);
}

// rt_next_week/cuda/main.cu:312-383 — the tree's default scene
inline hittable *rt_next_week_final_scene(unsigned char *data, int w, int h) {
  const int boxes_per_side = 20;
  const int num_obj = boxes_per_side * boxes_per_side + 10;
  static std::vector<hittable *> ret;
  ret.assign(num_obj, nullptr);
  auto ground = new lambertian(color(0.48, 0.83, 0.53));
  int index = 0;
  for (int i = 0; i < boxes_per_side; i++) {
    for (int j = 0; j < boxes_per_side; j++) {
      float bw = 100.0;
      float x0 = -1000.0f + i * bw;
      float z0 = -1000.0f + j * bw;
      float y0 = 0.0;
      float x1 = x0 + bw;
      float y1 = (float)random_double(1, 101);
      float z1 = z0 + bw;
      ret[index++] = new box(point3(x0, y0, z0), point3(x1, y1, z1), ground);
    }
  }
  auto light = new diffuse_light(color(7, 7, 7));
  ret[index++] = new xz_rect(123, 423, 147, 412, 554, light);
  auto center1 = point3(400, 400, 200);
  auto center2 = center1 + vec3(30, 0, 0);
  auto moving_sphere_material = new lambertian(color(0.7, 0.3, 0.1));
  ret[index++] = new moving_sphere(center1, center2, 0, 1, 50, moving_sphere_material);
  ret[index++] = new sphere(point3(260, 150, 45), 50, new dielectric(1.5));
  ret[index++] = new sphere(point3(0, 150, 145), 50, new metal(color(0.8, 0.8, 0.9), 1.0));
  auto sphere_dielectric_2 = new sphere(point3(360, 150, 145), 70, new dielectric(1.5));
  ret[index++] = sphere_dielectric_2;
  ret[index++] = new constant_medium(sphere_dielectric_2, 0.2, color(0.2, 0.4, 0.9));
  auto fog = new sphere(point3(0, 0, 0), 5000, new dielectric(1.5));
  ret[index++] = new constant_medium(fog, 0.0001, color(1, 1, 1));
  auto earth_texture = new image_texture(data, w, h);
  auto earth_surface = new lambertian(earth_texture);
  ret[index++] = new sphere(point3(400, 200, 400), 100, earth_surface);
  auto pertext = new noise_texture(0.1);
  ret[index++] = new sphere(point3(220, 280, 300), 80, new lambertian(pertext));
  auto white = new lambertian(color(.73, .73, .73));
  const int ns = 1000;
  static hittable *cluster[ns];
  for (int j = 0; j < ns; j++)
    cluster[j] = new sphere(point3(random_double(0, 165), random_double(0, 165), random_double(0, 165)), 10, white);
  ret[index++] = new translate(new rotate_y(new bvh_node(cluster, 0, ns, 0, 1), 15), vec3(-100, 270, 395));
  return new bvh_node(ret.data(), 0, index, 0.0, 1.0);
}
#endif
