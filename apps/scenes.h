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
#endif
