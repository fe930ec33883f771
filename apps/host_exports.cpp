// apps/host_exports.cpp — C exports of the C++ host layer for the Python tests: build a
// scene with the rtx classes, flatten it, hand the flattened arrays back. No CUDA here.
#include <cstring>
#include "scenes.h"
#include "scene_file.h"

static flat_scene g_flat;
static rt_camera g_cam;
static bool g_instancing = false; // rtx_host_set_instancing: translate / rotate_y as rt_instance instead of baked
static flat_scene fresh_flat() { flat_scene f; f.instancing = g_instancing; return f; }

extern "C" {
// which: 0 = weekend random_scene() under srand(seed) (seed 1 = glibc default), 1 = next_week
int rtx_host_build(int which, unsigned seed, double aspect, const char *obj_path) {
  srand(seed);
  g_flat = fresh_flat();
  if (which == 0) {
    hittable_list w = random_scene();
    w.flatten(g_flat, transform());
    g_cam = camera(point3(13, 2, 3), point3(0, 0, 0), vec3(0, 1, 0), 20, aspect, 0.1, 10.0).describe();
  } else if (which == 1) {
    std::vector<hittable *> d_list(22 * 22 + 1 + 3 + 1);
    hittable *root = next_week_random_scene(d_list.data());
    root->flatten(g_flat, transform());
    g_cam = camera(point3(13, 2, 3), point3(0, 0, 0), vec3(0, 1, 0), 20, aspect, 0.05,
                   (point3(13, 2, 3) - point3(0, 0, 0)).length(), 0.0, 1.0, true).describe();
  } else if (which == 3) { // main.cu:436-443
    hittable *root = cornell_box();
    root->flatten(g_flat, transform());
    g_cam = camera(point3(278, 278, -800), point3(278, 278, 0), vec3(0, 1, 0), 40, aspect, 0.0,
                   (point3(278, 278, -800) - point3(278, 278, 0)).length(), 0.0, 1.0, true).describe();
  } else if (which >= 4 && which <= 8) { // main.cu:417-459: perlin, earth, simple_light, cornell_smoke, final
    static std::vector<unsigned char> map;
    map = procedural_earth(64, 32);
    hittable *root = which == 4 ? two_perlin_spheres()
                     : which == 5 ? earth(map.data(), 64, 32)
                     : which == 6 ? simple_light()
                     : which == 7 ? cornell_smoke()
                                  : rt_next_week_final_scene(map.data(), 64, 32);
    root->flatten(g_flat, transform());
    const point3 from = which == 6 ? point3(26, 3, 6) : which == 7 ? point3(278, 278, -800)
                        : which == 8 ? point3(478, 278, -600) : point3(13, 2, 3);
    const point3 at = which == 6 ? point3(0, 2, 0) : which >= 7 ? point3(278, 278, 0) : point3(0, 0, 0);
    g_cam = camera(from, at, vec3(0, 1, 0), (which == 4 || which == 6) ? 20 : 40, aspect, 0.0, (from - at).length(), 0.0,
                   1.0, true).describe();
  } else {
    hittable *root = obj_model(obj_path);
    root->flatten(g_flat, transform());
    g_cam = camera(point3(1, 3, 7), point3(0, 2, 0), vec3(0, 1, 0), 60, aspect, 0.0,
                   (point3(1, 3, 7) - point3(0, 2, 0)).length(), 0.0, 1.0, true).describe();
  }
  return 0;
}
void rtx_host_set_instancing(int on) { g_instancing = on != 0; }
// n[0..1] = groups, instances
void rtx_host_counts3(int *n) { n[0] = (int)g_flat.groups.size(); n[1] = (int)g_flat.instances.size(); }
void rtx_host_get3(rt_group *g, rt_instance *i) {
  if (g) memcpy(g, g_flat.groups.data(), sizeof(rt_group) * g_flat.groups.size());
  if (i) memcpy(i, g_flat.instances.data(), sizeof(rt_instance) * g_flat.instances.size());
}
void rtx_host_counts(int *n) {
  n[0] = (int)g_flat.spheres.size(); n[1] = (int)g_flat.triangles.size();
  n[2] = (int)g_flat.quads.size(); n[3] = (int)g_flat.materials.size(); n[4] = g_flat.wants_accel;
}
// n[0..2] = media, perlin tables, images; dims = (w, h) per image
void rtx_host_counts2(int *n, int *dims) {
  n[0] = (int)g_flat.media.size(); n[1] = (int)g_flat.perlin_tables.size(); n[2] = (int)g_flat.images.size();
  for (size_t i = 0; i < g_flat.images.size() && dims; i++) { dims[2 * i] = g_flat.images[i].width; dims[2 * i + 1] = g_flat.images[i].height; }
}
void rtx_host_get2(rt_medium *m, rt_perlin *p, unsigned char **image_bytes) {
  if (m) memcpy(m, g_flat.media.data(), sizeof(rt_medium) * g_flat.media.size());
  if (p) memcpy(p, g_flat.perlin_tables.data(), sizeof(rt_perlin) * g_flat.perlin_tables.size());
  for (size_t i = 0; i < g_flat.images.size() && image_bytes; i++)
    memcpy(image_bytes[i], g_flat.images[i].rgb, (size_t)g_flat.images[i].width * g_flat.images[i].height * 3);
}
void rtx_host_get(rt_sphere *s, rt_triangle *t, rt_quad *q, rt_material *m, rt_camera *c) {
  if (s) memcpy(s, g_flat.spheres.data(), sizeof(rt_sphere) * g_flat.spheres.size());
  if (t) memcpy(t, g_flat.triangles.data(), sizeof(rt_triangle) * g_flat.triangles.size());
  if (q) memcpy(q, g_flat.quads.data(), sizeof(rt_quad) * g_flat.quads.size());
  if (m) memcpy(m, g_flat.materials.data(), sizeof(rt_material) * g_flat.materials.size());
  if (c) *c = g_cam;
}
}

extern "C" int rtx_host_write_png(const char *path, const unsigned char *rgb, int w, int h) {
  return rtx::write_png_file(path, rgb, w, h) ? 0 : 1;
}

// Parse + flatten a text scene file (include/rtx/scene_file.h). opts: profile, sky_gradient, max_depth,
// flags, width, height, spp; fopts: background rgb, t_min. Returns 0, or 1 with the message in err.
extern "C" int rtx_host_load_scene_file(const char *path, int width, int height, int *opts, float *fopts, char *err,
                                        int errcap) {
  try {
    scene_file sf = load_scene_file(path);
    g_flat = fresh_flat();
    static scene_file keep; // the flattened image pointers refer to textures owned by the scene
    keep = sf;
    if (keep.opt.instancing) g_flat.instancing = true; // the file's `flags instancing`
    keep.world.flatten(g_flat, transform());
    const int w = width > 0 ? width : keep.width, h = height > 0 ? height : keep.height;
    g_cam = keep.make_camera(w, h).describe();
    opts[0] = keep.opt.profile; opts[1] = keep.opt.sky_gradient ? 1 : 0; opts[2] = keep.opt.max_depth;
    opts[3] = (int)keep.opt.flags; opts[4] = w; opts[5] = h; opts[6] = keep.spp;
    for (int a = 0; a < 3; a++) fopts[a] = (float)keep.opt.background[a];
    fopts[3] = (float)keep.opt.t_min;
    return 0;
  } catch (const std::exception &e) {
    snprintf(err, errcap, "%s", e.what());
    return 1;
  }
}
