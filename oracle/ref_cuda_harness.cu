// oracle/ref_cuda_harness.cu — ORACLE L0-CUDA (test infrastructure, NOT product code).
//
// Executes the reference's UNMODIFIED CUDA-tree device code on a GPU and records what it computes,
// so that the plain-C restatement (oracle L1, orc32_*) and the B200 kernels can be pinned to the
// reference's own results for the rows the CPU renderer does not cover (moving spheres, triangles,
// rects, checker / noise / image textures, emitters, constant media, translate / rotate_y).
//
// Nothing is copied: this file #includes the reference translation unit where it lies under
// $(REFERENCE) (with `main` renamed), exactly as oracle/ref_harness.cpp does for the CPU renderer:
//   TREE == 1  rt_next_week/cuda/main.cu   (+ its headers sphere.h, moving_sphere.h, aarect.h, box.h,
//              aabb.h, bvh.h, hittable.h, hittable_list.h, material.h, texture.h, perlin.h,
//              constant_medium.h) — normals never flip (hittable.h:29), THICKNESS 0.1, t_min 1e-3,
//              get_color main.cu:48-105
//   TREE == 2  triangles/cuda/obj_render.cu (+ include/*.h incl. triangle.h) — normals face the
//              ray (include/hittable.h:29), THICKNESS 0.01, get_color obj_render.cu:20-86
// Objects are created with device-side `new` in a <<<1,1>>> kernel, as the reference's
// create_world does, from a flattened scene file written by tools/make_cuda_golden.py; every object
// gets its own material instance so that hit_record::mat_ptr identifies the object that was hit.
// Random numbers: cuRAND XORWOW, curand_init(seed + ray, 0, 0) per ray (restated by the oracle's
// replay generator, orc_xorwow_*), so scatter / get_color / constant_medium::hit can be replayed.
//
// Build (oracle/Makefile, only where /root/reference exists):
//   nvcc -arch=sm_100 [-fmad=false] -DTREE=1 -I$(REFERENCE)/rt_next_week/cuda ...
// Run on the GPU box:  oracle/_ref/ref_cuda_nw <in.bin> <out.bin>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>
#include <cuda.h>
#include <curand_kernel.h>
#include <thrust/device_vector.h>
#include <thrust/execution_policy.h>
#include <thrust/sort.h>

// perlin keeps its tables private (perlin.h:73-79); the harness must read them back so that the
// other implementations can be given the very same tables. The reference source is not touched.
#define private public
#define main ref_main_unused
#if TREE == 1
#include "main.cu"
#else
#include "obj_render.cu"
#endif
#undef main
#undef private

#define HCK(x)                                                                           \
  do {                                                                                   \
    cudaError_t e_ = (x);                                                                \
    if (e_ != cudaSuccess) {                                                             \
      fprintf(stderr, "%s failed: %s (%s:%d)\n", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                           \
    }                                                                                    \
  } while (0)

// ---- file records (all little-endian, 4-byte fields) — mirrored by tools/make_cuda_golden.py
struct HHeader {
  int magic, tree, n_spheres, n_tris, n_quads, n_mats, n_images, n_media, n_rays, n_texq, n_colorrays, reserved;
  float t_min, t_max, background[3], time0, time1, pad;
  unsigned long long seed;
};
struct HSphere { float c0[3], radius, c1[3]; int material; float time0, time1; int moving, reserved; }; // = rt_sphere
struct HTri { float v0[3], v1[3], v2[3], vn0[3], vn1[3], vn2[3]; int material, pad; };
struct HQuad { int axis; float a0, a1, b0, b1, k; int material; };                                      // = rt_quad
struct HMat { int type, texture; float albedo[3], param, albedo2[3], reserved; };                       // = rt_material
struct HInst { float angle_deg, off[3]; int flag; }; // translate(rotate_y(object, angle), off) when flag
struct HMedium { int shape; float p0[3], p1[3], angle_deg, off[3], density, albedo[3]; };
struct HTexQ { int material; float u, v, p[3]; };
struct HHit { int hit, obj; float t, p[3], n[3], u, v; int front_face; };
struct HScatter { int ok; float att[3], o[3], d[3], tm, emitted[3]; int draws; };
struct HMediumHit { int hit; float t, p[3]; int draws; };

struct World {
  hittable **objs;      // [n_prims + n_media] list order: spheres, triangles, quads, media
  hittable **objs_bvh;  // copy (the bvh_node constructor sorts it in place)
  material **obj_mat;   // per object: its own material instance (identifies the object)
  abstract_texture **tex;
  hittable **world_list, **world_bvh, **world_list_media, **world_bvh_media;
  color **background;
  int n_prims, n_objs;
};

__device__ int find_obj(const World &w, const material *m) {
  for (int i = 0; i < w.n_objs; i++)
    if (w.obj_mat[i] == m) return i;
  return -1;
}

__device__ abstract_texture *make_texture(const HMat &m, unsigned char **img_data, const int *img_wh, curandState *rs) {
  color a(m.albedo[0], m.albedo[1], m.albedo[2]);
  if (m.texture == 1) return new checker_texture(a, color(m.albedo2[0], m.albedo2[1], m.albedo2[2]));
  if (m.texture == 2) return new noise_texture(m.albedo2[0], rs);
  if (m.texture == 3) {
    const int k = (int)m.albedo2[0];
    return new image_texture(img_data[k], img_wh[2 * k], img_wh[2 * k + 1]);
  }
  return new solid_color(a);
}

__device__ material *make_material(const HMat &m, abstract_texture *tex) {
  switch (m.type) {
  case 0: return new lambertian(tex);
  case 1: return new metal(tex, m.param);
  case 2: return new dielectric(m.param);
  case 3: return new diffuse_light(tex);
  default: return new isotropic(tex);
  }
}

__global__ void k_build(HHeader H, const HSphere *sph, const HTri *tri, const HQuad *quad, const HMat *mats, const HInst *inst,
                        const HMedium *media, unsigned char **img_data, const int *img_wh, World w, curandState *rs,
                        float *perlin_out, float *inst_sc, float *med_sc) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  curand_init(H.seed, 0, 0, rs);
  int n_noise = 0;
  for (int m = 0; m < H.n_mats; m++) {
    w.tex[m] = make_texture(mats[m], img_data, img_wh, rs);
    if (mats[m].texture == 2) { // dump the tables the constructor drew (perlin.h:9-19)
      perlin *pn = ((noise_texture *)w.tex[m])->noise;
      float *o = perlin_out + (size_t)n_noise * (256 * 3 + 3 * 256);
      for (int i = 0; i < 256; i++) { o[3 * i] = pn->ranvec[i].x(); o[3 * i + 1] = pn->ranvec[i].y(); o[3 * i + 2] = pn->ranvec[i].z(); }
      for (int i = 0; i < 256; i++) { o[768 + i] = (float)pn->perm_x[i]; o[1024 + i] = (float)pn->perm_y[i]; o[1280 + i] = (float)pn->perm_z[i]; }
      n_noise++;
    }
  }
  int k = 0;
  for (int i = 0; i < H.n_spheres; i++, k++) {
    const HSphere &s = sph[i];
    w.obj_mat[k] = make_material(mats[s.material], w.tex[s.material]);
    if (s.moving)
      w.objs[k] = new moving_sphere(vec3(s.c0[0], s.c0[1], s.c0[2]), vec3(s.c1[0], s.c1[1], s.c1[2]), s.time0, s.time1, s.radius, w.obj_mat[k]);
    else
      w.objs[k] = new sphere(vec3(s.c0[0], s.c0[1], s.c0[2]), s.radius, w.obj_mat[k]);
  }
  for (int i = 0; i < H.n_tris; i++, k++) {
    const HTri &t = tri[i];
    w.obj_mat[k] = make_material(mats[t.material], w.tex[t.material]);
#if TREE == 2
    w.objs[k] = new triangle(vec3(t.v0[0], t.v0[1], t.v0[2]), vec3(t.v1[0], t.v1[1], t.v1[2]), vec3(t.v2[0], t.v2[1], t.v2[2]),
                             vec3(t.vn0[0], t.vn0[1], t.vn0[2]), vec3(t.vn1[0], t.vn1[1], t.vn1[2]), vec3(t.vn2[0], t.vn2[1], t.vn2[2]),
                             w.obj_mat[k]);
#else
    w.objs[k] = new sphere(vec3(0, -1e9f, 0), 1.0f, w.obj_mat[k]); // the rt_next_week tree has no triangle
#endif
  }
  for (int i = 0; i < H.n_quads; i++, k++) {
    const HQuad &q = quad[i];
    w.obj_mat[k] = make_material(mats[q.material], w.tex[q.material]);
    if (q.axis == 2) w.objs[k] = new xy_rect(q.a0, q.a1, q.b0, q.b1, q.k, w.obj_mat[k]);
    else if (q.axis == 1) w.objs[k] = new xz_rect(q.a0, q.a1, q.b0, q.b1, q.k, w.obj_mat[k]);
    else w.objs[k] = new yz_rect(q.a0, q.a1, q.b0, q.b1, q.k, w.obj_mat[k]);
  }
  for (int i = 0; i < k; i++) {
    inst_sc[2 * i] = 0.f; inst_sc[2 * i + 1] = 1.f;
    if (inst[i].flag) {
      rotate_y *ry = new rotate_y(w.objs[i], inst[i].angle_deg);
      inst_sc[2 * i] = ry->sin_theta; inst_sc[2 * i + 1] = ry->cos_theta; // sinf / cosf as the device evaluates them
      w.objs[i] = new translate(ry, vec3(inst[i].off[0], inst[i].off[1], inst[i].off[2]));
    }
  }
  for (int m = 0; m < H.n_media; m++, k++) {
    const HMedium &md = media[m];
    hittable *b;
    material *bm = new dielectric(1.5f); // the boundary's own material is never used (main.cu:297-305)
    med_sc[2 * m] = 0.f; med_sc[2 * m + 1] = 1.f;
    if (md.shape == 0) b = new sphere(vec3(md.p0[0], md.p0[1], md.p0[2]), md.p1[0], bm);
    else {
      rotate_y *ry = new rotate_y(new box(vec3(md.p0[0], md.p0[1], md.p0[2]), vec3(md.p1[0], md.p1[1], md.p1[2]), bm), md.angle_deg);
      med_sc[2 * m] = ry->sin_theta; med_sc[2 * m + 1] = ry->cos_theta;
      b = new translate(ry, vec3(md.off[0], md.off[1], md.off[2]));
    }
    constant_medium *cm = new constant_medium(b, md.density, color(md.albedo[0], md.albedo[1], md.albedo[2]));
    w.objs[k] = cm;
    w.obj_mat[k] = cm->phase_function;
  }
  *w.background = new color(H.background[0], H.background[1], H.background[2]);
  *w.world_list = new hittable_list(w.objs, w.n_prims);
  *w.world_list_media = new hittable_list(w.objs, w.n_objs);
  for (int i = 0; i < w.n_objs; i++) w.objs_bvh[i] = w.objs[i];
  *w.world_bvh = w.n_prims > 0 ? (hittable *)new bvh_node(w.objs_bvh, 0, w.n_prims, H.time0, H.time1, rs) : *w.world_list;
  // the reference's scenes put the media INTO the BVH (main.cu:297-305 add them to the list it builds from)
  for (int i = 0; i < w.n_objs; i++) w.objs_bvh[i] = w.objs[i];
  *w.world_bvh_media = w.n_objs > 0 ? (hittable *)new bvh_node(w.objs_bvh, 0, w.n_objs, H.time0, H.time1, rs) : *w.world_list_media;
}

__device__ ray load_ray(const float *rays, int i) {
  const float *r = rays + 8 * (size_t)i;
  return ray(vec3(r[0], r[1], r[2]), vec3(r[4], r[5], r[6]), r[3]);
}

__device__ int draws_between(const curandState &a, const curandState &b) { return (int)((b.d - a.d) / 362437u); }

// closest hit over the list (hittable_list::hit) and through bvh_node::hit, surfaces only
__global__ void k_hits(HHeader H, World w, const float *rays, HHit *out_list, HHit *out_bvh) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H.n_rays) return;
  const ray r = load_ray(rays, i);
  curandState st;
  curand_init(H.seed + 1 + i, 0, 0, &st);
  for (int pass = 0; pass < 2; pass++) {
    hit_record rec;
    hittable *world = pass == 0 ? *w.world_list : *w.world_bvh;
    HHit o;
    memset(&o, 0, sizeof o);
    o.obj = -1;
    if (world->hit(r, H.t_min, H.t_max, rec, &st)) {
      o.hit = 1; o.obj = find_obj(w, rec.mat_ptr); o.t = rec.t;
      o.p[0] = rec.p.x(); o.p[1] = rec.p.y(); o.p[2] = rec.p.z();
      o.n[0] = rec.normal.x(); o.n[1] = rec.normal.y(); o.n[2] = rec.normal.z();
      o.u = rec.u; o.v = rec.v; o.front_face = rec.front_face ? 1 : 0;
    }
    (pass == 0 ? out_list : out_bvh)[i] = o;
  }
}

// material::scatter + emitted at the list's closest hit, with the ray's own XORWOW stream
__global__ void k_scatter(HHeader H, World w, const float *rays, HScatter *out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H.n_rays) return;
  const ray r = load_ray(rays, i);
  curandState st;
  curand_init(H.seed + 1 + i, 0, 0, &st);
  const curandState st0 = st;
  HScatter o;
  memset(&o, 0, sizeof o);
  o.ok = -1; // no hit
  hit_record rec;
  if ((*w.world_list)->hit(r, H.t_min, H.t_max, rec, &st)) {
    ray scattered(vec3(0, 0, 0), vec3(0, 0, 0), 0);
    vec3 att(0, 0, 0);
    const color em = rec.mat_ptr->emitted(rec.u, rec.v, rec.p);
    o.ok = rec.mat_ptr->scatter(r, rec, att, scattered, &st) ? 1 : 0;
    for (int c = 0; c < 3; c++) {
      o.att[c] = att[c]; o.emitted[c] = em[c];
      o.o[c] = scattered.origin()[c]; o.d[c] = scattered.direction()[c];
    }
    o.tm = scattered.time();
  }
  o.draws = draws_between(st0, st);
  out[i] = o;
}

// constant_medium::hit of every medium for every ray (t_min, t_max as the integrator passes them)
__global__ void k_media(HHeader H, World w, const float *rays, HMediumHit *out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H.n_rays) return;
  const ray r = load_ray(rays, i);
  for (int m = 0; m < H.n_media; m++) {
    curandState st;
    curand_init(H.seed + 1 + i, 0, 0, &st);
    const curandState st0 = st;
    hit_record rec;
    HMediumHit o;
    memset(&o, 0, sizeof o);
    if (w.objs[w.n_prims + m]->hit(r, H.t_min, H.t_max, rec, &st)) {
      o.hit = 1; o.t = rec.t;
      o.p[0] = rec.p.x(); o.p[1] = rec.p.y(); o.p[2] = rec.p.z();
    }
    o.draws = draws_between(st0, st);
    out[(size_t)i * H.n_media + m] = o;
  }
}

// the integrator get_color (main.cu:48-105 / obj_render.cu:20-86) on the first n_colorrays rays: list world
// and bvh world (both with the media), each from the ray's own stream
__global__ void k_color(HHeader H, World w, const float *rays, float *out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H.n_colorrays) return;
  const ray r = load_ray(rays, i);
  for (int pass = 0; pass < 2; pass++) {
    curandState st;
    curand_init(H.seed + 1 + i, 0, 0, &st);
    const curandState st0 = st;
    const vec3 c = get_color(r, w.background, pass == 0 ? w.world_list_media : w.world_bvh_media, &st);
    float *o = out + ((size_t)i * 2 + pass) * 4;
    o[0] = c.x(); o[1] = c.y(); o[2] = c.z(); o[3] = (float)draws_between(st0, st);
  }
}

__global__ void k_texq(HHeader H, World w, const HTexQ *q, float *out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H.n_texq) return;
  const color c = w.tex[q[i].material]->value(q[i].u, q[i].v, point3(q[i].p[0], q[i].p[1], q[i].p[2]));
  out[3 * i] = c.x(); out[3 * i + 1] = c.y(); out[3 * i + 2] = c.z();
}

// known answers of the generator itself: the first 8 curand_uniform of seeds 0..15
__global__ void k_xorwow(unsigned long long seed, float *out) {
  const int i = threadIdx.x;
  curandState st;
  curand_init(seed + i, 0, 0, &st);
  for (int k = 0; k < 8; k++) out[8 * i + k] = curand_uniform(&st);
}

template <class T> static T *upload(const std::vector<T> &v) {
  T *d = nullptr;
  HCK(cudaMalloc(&d, std::max<size_t>(v.size(), 1) * sizeof(T)));
  if (!v.empty()) HCK(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return d;
}
template <class T> static void rd(FILE *f, std::vector<T> &v, size_t n) {
  v.resize(n);
  if (n && fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "short input file\n"); exit(2); }
}
template <class T> static void wr(FILE *f, const T *d_ptr, size_t n) {
  std::vector<T> h(n);
  if (n) HCK(cudaMemcpy(h.data(), d_ptr, n * sizeof(T), cudaMemcpyDeviceToHost));
  const unsigned long long cnt = n * sizeof(T);
  fwrite(&cnt, 8, 1, f);
  if (n) fwrite(h.data(), sizeof(T), n, f);
}

int main(int argc, char **argv) {
  if (argc < 3) { fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]); return 2; }
  FILE *fi = fopen(argv[1], "rb");
  if (!fi) { perror(argv[1]); return 2; }
  HHeader H;
  if (fread(&H, sizeof H, 1, fi) != 1 || H.magic != 0x31484352 || H.tree != TREE) { fprintf(stderr, "bad header (tree %d)\n", TREE); return 2; }
  std::vector<HSphere> sph; std::vector<HTri> tri; std::vector<HQuad> quad; std::vector<HMat> mats; std::vector<HInst> inst;
  std::vector<HMedium> media; std::vector<HTexQ> texq; std::vector<float> rays; std::vector<int> img_wh;
  rd(fi, sph, H.n_spheres); rd(fi, tri, H.n_tris); rd(fi, quad, H.n_quads); rd(fi, mats, H.n_mats);
  const int n_prims = H.n_spheres + H.n_tris + H.n_quads, n_objs = n_prims + H.n_media;
  rd(fi, inst, n_prims); rd(fi, media, H.n_media); rd(fi, img_wh, 2 * (size_t)H.n_images);
  std::vector<unsigned char *> img_ptrs(H.n_images, nullptr);
  for (int k = 0; k < H.n_images; k++) {
    std::vector<unsigned char> px;
    rd(fi, px, 3 * (size_t)img_wh[2 * k] * img_wh[2 * k + 1]);
    img_ptrs[k] = upload(px);
  }
  rd(fi, texq, H.n_texq); rd(fi, rays, 8 * (size_t)H.n_rays);
  fclose(fi);

  HCK(cudaDeviceSetLimit(cudaLimitStackSize, 32768ULL));                        // main.cu:483
  HCK(cudaDeviceSetLimit(cudaLimitMallocHeapSize, 512ULL * 1024ULL * 1024ULL)); // obj_render.cu:789
  World w;
  memset(&w, 0, sizeof w);
  w.n_prims = n_prims; w.n_objs = n_objs;
  HCK(cudaMalloc(&w.objs, sizeof(void *) * std::max(n_objs, 1)));
  HCK(cudaMalloc(&w.objs_bvh, sizeof(void *) * std::max(n_objs, 1)));
  HCK(cudaMalloc(&w.obj_mat, sizeof(void *) * std::max(n_objs, 1)));
  HCK(cudaMalloc(&w.tex, sizeof(void *) * std::max(H.n_mats, 1)));
  HCK(cudaMalloc(&w.world_list, sizeof(void *))); HCK(cudaMalloc(&w.world_bvh, sizeof(void *)));
  HCK(cudaMalloc(&w.world_list_media, sizeof(void *))); HCK(cudaMalloc(&w.world_bvh_media, sizeof(void *)));
  HCK(cudaMalloc(&w.background, sizeof(void *)));
  curandState *rs; HCK(cudaMalloc(&rs, sizeof(curandState)));
  int n_noise = 0;
  for (auto &m : mats) n_noise += m.texture == 2;
  float *d_perlin; HCK(cudaMalloc(&d_perlin, sizeof(float) * 1536 * std::max(n_noise, 1)));
  float *d_inst_sc, *d_med_sc;
  HCK(cudaMalloc(&d_inst_sc, sizeof(float) * 2 * std::max(n_prims, 1))); HCK(cudaMalloc(&d_med_sc, sizeof(float) * 2 * std::max(H.n_media, 1)));
  unsigned char **d_img = upload(img_ptrs);
  int *d_wh = upload(img_wh);
  HSphere *d_sph = upload(sph); HTri *d_tri = upload(tri); HQuad *d_quad = upload(quad); HMat *d_mats = upload(mats);
  HInst *d_inst = upload(inst); HMedium *d_media = upload(media); HTexQ *d_texq = upload(texq); float *d_rays = upload(rays);
  k_build<<<1, 1>>>(H, d_sph, d_tri, d_quad, d_mats, d_inst, d_media, d_img, d_wh, w, rs, d_perlin, d_inst_sc, d_med_sc);
  HCK(cudaGetLastError()); HCK(cudaDeviceSynchronize());

  const int TB = 64, g = (H.n_rays + TB - 1) / TB;
  HHit *o_list, *o_bvh; HScatter *o_sc; HMediumHit *o_med; float *o_col, *o_tex, *o_x;
  HCK(cudaMalloc(&o_list, sizeof(HHit) * std::max(H.n_rays, 1))); HCK(cudaMalloc(&o_bvh, sizeof(HHit) * std::max(H.n_rays, 1)));
  HCK(cudaMalloc(&o_sc, sizeof(HScatter) * std::max(H.n_rays, 1)));
  HCK(cudaMalloc(&o_med, sizeof(HMediumHit) * std::max(H.n_rays * H.n_media, 1)));
  HCK(cudaMalloc(&o_col, sizeof(float) * 8 * std::max(H.n_colorrays, 1)));
  HCK(cudaMalloc(&o_tex, sizeof(float) * 3 * std::max(H.n_texq, 1)));
  HCK(cudaMalloc(&o_x, sizeof(float) * 128));
  if (H.n_rays) {
    k_hits<<<g, TB>>>(H, w, d_rays, o_list, o_bvh);
    HCK(cudaGetLastError()); HCK(cudaDeviceSynchronize());
    k_scatter<<<g, TB>>>(H, w, d_rays, o_sc);
    HCK(cudaGetLastError()); HCK(cudaDeviceSynchronize());
    if (H.n_media) { k_media<<<g, TB>>>(H, w, d_rays, o_med); HCK(cudaGetLastError()); HCK(cudaDeviceSynchronize()); }
  }
  if (H.n_colorrays) { k_color<<<(H.n_colorrays + TB - 1) / TB, TB>>>(H, w, d_rays, o_col); HCK(cudaGetLastError()); HCK(cudaDeviceSynchronize()); }
  if (H.n_texq) { k_texq<<<(H.n_texq + TB - 1) / TB, TB>>>(H, w, d_texq, o_tex); HCK(cudaGetLastError()); HCK(cudaDeviceSynchronize()); }
  k_xorwow<<<1, 16>>>(H.seed, o_x);
  HCK(cudaGetLastError()); HCK(cudaDeviceSynchronize());

  FILE *fo = fopen(argv[2], "wb");
  if (!fo) { perror(argv[2]); return 2; }
  fwrite(&H, sizeof H, 1, fo);
  wr(fo, o_list, H.n_rays); wr(fo, o_bvh, H.n_rays); wr(fo, o_sc, H.n_rays); wr(fo, o_med, (size_t)H.n_rays * H.n_media);
  wr(fo, o_col, 8 * (size_t)H.n_colorrays); wr(fo, o_tex, 3 * (size_t)H.n_texq); wr(fo, d_perlin, 1536 * (size_t)n_noise);
  wr(fo, o_x, 128);
  wr(fo, d_inst_sc, 2 * (size_t)n_prims); wr(fo, d_med_sc, 2 * (size_t)H.n_media);
  fclose(fo);
  fprintf(stderr, "ref_cuda_harness tree %d: %d objects (%d media), %d rays, %d colour rays, %d texture queries -> %s\n", TREE,
          n_objs, H.n_media, H.n_rays, H.n_colorrays, H.n_texq, argv[2]);
  return 0;
}
