/* rt_capi.h — C ABI of the B200 path-tracer core (libb200rt.so).
 *
 * The reference (xyloid/a_dive_into_ray_tracing) has no FFI; its renderers are
 * single translation units whose host code launches CUDA kernels directly.
 * Each entry point below replaces one of those launch sites / host loops and
 * is what a maintainer's host C++ would call instead:
 *
 *   rt_create / rt_destroy      cudaMalloc/cudaFree + checkCudaErrors->exit(99)
 *                               accelerated-rt-cuda/final.cu:13-24,176-203,234-245
 *   rt_scene_upload             create_world<<<1,1>>> (device-side `new` of every
 *                               sphere/material/camera) final.cu:100-143,
 *                               rt_next_week/cuda/main.cu:386-467,
 *                               triangles/cuda/obj_render.cu:631-741
 *   rt_accel_build              `new bvh_node(list,0,n,t0,t1,rng)`
 *                               rt_next_week/cuda/bvh.h:139-196 (one GPU thread)
 *   rt_trace_closest            hittable::hit — hittable_list.h:20-34 (CPU),
 *                               accelerated-rt-cuda/hittable_list.h:22-37,
 *                               rt_next_week/cuda/bvh.h:78-137   (parity hook)
 *   rt_render                   worker() rt_in_one_weekend/main.cpp:267-290;
 *                               render_init+render<<<grid,8x8>>> final.cu:62-96,
 *                               main.cu:113-149, obj_render.cu:94-130
 *   rt_resolve                  write_color rt_in_one_weekend/color.h:14-28;
 *                               the /ns, sqrt in final.cu:91-95 and the
 *                               int(255.99*x) loop final.cu:223-232
 *   rt_last_error               the std::cerr text of check_cuda final.cu:15-24
 *
 * Conventions: plain pointers and sizes only; every function returns an
 * rt_status (0 = OK) and never throws, prints or exits; the caller owns all
 * host buffers, the library owns all device buffers it allocates; a context is
 * bound to ONE CUDA device and is single-host-thread; separate contexts are
 * independent (multi-GPU = one context per device, see rt_render's spp range).
 *
 * Pixel indexing everywhere: index = j*W + i with j = 0 the BOTTOM row, as in
 * the reference (main.cpp:274-275, final.cu:80-81); writers flip rows.
 */
#ifndef RT_CAPI_H
#define RT_CAPI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RT_CAPI_VERSION 5

typedef struct rt_ctx rt_ctx;

typedef enum rt_status {
  RT_OK = 0,
  RT_ERR_INVALID = 1,  /* bad argument / malformed scene */
  RT_ERR_CUDA = 2,     /* a CUDA runtime call failed; see rt_last_error */
  RT_ERR_STATE = 3,    /* call order (e.g. render before scene upload) */
  RT_ERR_NOMEM = 4,
  RT_ERR_NODEVICE = 5  /* no CUDA device: there is NO CPU fallback */
} rt_status;

/* Semantic profile = which reference renderer defines the expected image
 * (SURVEY.md §7 table). */
typedef enum rt_profile {
  RT_PROFILE_WEEKEND_CPU = 0, /* rt_in_one_weekend (book v3, front_face, true Lambertian) */
  RT_PROFILE_FINAL_CU = 1,    /* accelerated-rt-cuda/final.cu */
  RT_PROFILE_NEXT_WEEK = 2    /* rt_next_week/cuda + triangles/cuda (emission, background) */
} rt_profile;

/* rt_config.flags / rt_scene_desc.flags */
#define RT_FLAG_FLIP_NORMALS 1u      /* triangles/cuda/include/hittable.h:29 (normal faces the ray) */
#define RT_FLAG_DEPTH_BACKGROUND 2u  /* obj_render.cu:78-83: depth exhausted -> unwound background */
#define RT_FLAG_COUNTERS 4u          /* maintain node/prim test counters (slower kernel variant) */
#define RT_FLAG_REFERENCE_MEDIUM 8u  /* constant_medium.h:66: the scattered ray leaves from the boundary ENTRY point
                                      * r.at(rec1.t) (entry clamped to the ray origin), not from the scatter point */

typedef enum rt_prim_type {
  RT_PRIM_SPHERE = 0,
  RT_PRIM_TRIANGLE = 1,
  RT_PRIM_QUAD = 2,
  RT_PRIM_MEDIUM = 3, /* constant_medium: never returned by rt_trace_closest (stochastic) */
  RT_PRIM_INSTANCE = 4, /* leaf of the top-level tree naming an rt_instance: internal, never returned */
  RT_PRIM_BOX = 5       /* leaf naming six rects that form a `box` (box.h:41-58), tested together: internal, the
                         * hit reports the rect (RT_PRIM_QUAD) that was hit */
} rt_prim_type;
/* primitive id returned by rt_trace_closest: (type << 28) | index-within-type; -1 = miss */
#define RT_PRIM_ID(type, index) ((int32_t)(((uint32_t)(type) << 28) | (uint32_t)(index)))
#define RT_PRIM_TYPE_OF(id) ((int)(((uint32_t)(id)) >> 28))
#define RT_PRIM_INDEX_OF(id) ((int)(((uint32_t)(id)) & 0x0FFFFFFFu))

typedef enum rt_material_type {
  RT_MAT_LAMBERTIAN = 0,
  RT_MAT_METAL = 1,
  RT_MAT_DIELECTRIC = 2,
  RT_MAT_DIFFUSE_LIGHT = 3,
  RT_MAT_ISOTROPIC = 4 /* phase function of a constant_medium (rt_next_week/cuda/material.h:178-195) */
} rt_material_type;

typedef enum rt_texture_type {
  RT_TEX_SOLID = 0,
  RT_TEX_CHECKER = 1,
  RT_TEX_NOISE = 2, /* noise_texture (texture.h:55-75): albedo * 0.5 (1 + sin(scale z + 10 turb(scale p))) */
  RT_TEX_IMAGE = 3  /* image_texture (texture.h:77-124) looked up with the primitive's (u, v) */
} rt_texture_type;

typedef struct rt_config {
  int32_t device;   /* CUDA device ordinal */
  int32_t profile;  /* rt_profile */
  uint32_t flags;
  uint32_t reserved;
  uint64_t seed;    /* Philox key */
} rt_config;

/* sphere / moving_sphere (sphere.h:10-11, moving_sphere.h:12-15) */
typedef struct rt_sphere {
  float center0[3];
  float radius;      /* may be negative (hollow-glass trick, main.cpp:209) */
  float center1[3];  /* == center0 when not moving */
  int32_t material;
  float time0, time1;
  int32_t moving;    /* 0 = sphere, 1 = moving_sphere */
  int32_t reserved;
} rt_sphere;

/* triangle (triangles/cuda/include/triangle.h:17-53): vertices as stored by the
 * ctor and the oriented, UN-normalised face normal it computes (:42-44). */
typedef struct rt_triangle {
  float v0[3], v1[3], v2[3];
  float normal[3];
  int32_t material;
} rt_triangle;

/* xy_rect / xz_rect / yz_rect (aarect.h:15-17,70-72,126-128).
 * axis = the constant axis: 2 -> xy_rect (a=x,b=y), 1 -> xz_rect (a=x,b=z),
 * 0 -> yz_rect (a=y,b=z). */
typedef struct rt_quad {
  int32_t axis;
  float a0, a1, b0, b1, k;
  int32_t material;
} rt_quad;

typedef struct rt_material {
  int32_t type;      /* rt_material_type */
  int32_t texture;   /* rt_texture_type (albedo / emit) */
  float albedo[3];   /* solid colour, checker `even` (texture.h:33-53), noise tint (reference: 1,1,1) */
  float param;       /* metal: fuzz (already clamped to <=1); dielectric: index */
  float albedo2[3];  /* checker `odd`; noise: {scale, perlin table index, -}; image: {image index, -, -} */
  float reserved;
} rt_material;

/* perlin (rt_next_week/cuda/perlin.h:9-19,76-100): the 256 random gradient vectors and the three
 * permutation tables of one `perlin` object, generated by the host. */
typedef struct rt_perlin {
  float ranvec[256][3];
  int32_t perm_x[256], perm_y[256], perm_z[256];
} rt_perlin;

/* image_texture data (texture.h:83-88): 3 bytes per pixel, row 0 = top. */
typedef struct rt_image {
  int32_t width, height;
  const uint8_t *rgb;
} rt_image;

/* constant_medium (rt_next_week/cuda/constant_medium.h:10-33) over a convex boundary:
 * shape 0 = sphere(center = p0, radius = p1[0]); shape 1 = box(p0, p1) rotated about y by the
 * angle with (sin_y, cos_y) and then translated by `offset` (box.h + hittable.h rotate_y/translate,
 * main.cu:297-305). material must be RT_MAT_ISOTROPIC. */
typedef struct rt_medium {
  int32_t shape;
  float p0[3], p1[3];
  float sin_y, cos_y;
  float offset[3];
  float density;
  int32_t material;
} rt_medium;

/* camera as its constructor leaves it (camera.h:8-45; rt_next_week/cuda/camera.h:25-61) */
typedef struct rt_camera {
  float origin[3];
  float lower_left_corner[3];
  float horizontal[3];
  float vertical[3];
  float u[3], v[3], w[3];
  float lens_radius;
  float time0, time1;
} rt_camera;

/* ---- version 5: first-class instancing (two-level BVH). Replaces the reference's `translate` / `rotate_y`
 * wrappers (rt_next_week/cuda/hittable.h:49-190; per triangle in triangles/cuda/obj_render.cu:498-511, around the
 * 1000-sphere cluster in rt_next_week/cuda/main.cu:373-381): the wrapped object keeps its OBJECT-space primitives and
 * its own bottom-level tree, the ray is moved into object space on entry and t is preserved, exactly as
 * translate::hit / rotate_y::hit do (hittable.h:66-79,156-190).
 *
 * rt_group = one object: ranges of the scene's primitive arrays, given in object space. Primitives covered by any
 * group are NOT part of the world level; groups may share primitives. An empty group is allowed.
 * rt_instance = one placement of a group: p_world = M p_object with M = {m[0..3]; m[4..7]; m[8..11]} (rows of a
 * 3x4 matrix: rotation | translation). M must be RIGID (orthonormal rotation, det +1, no scale): only then is the
 * ray parameter t the same in both spaces, which is what the reference's wrappers rely on. */
typedef struct rt_group {
  int32_t first_sphere, n_spheres;
  int32_t first_triangle, n_triangles;
  int32_t first_quad, n_quads;
  int32_t reserved[2];
} rt_group;

typedef struct rt_instance {
  float m[12];
  int32_t group;
  int32_t reserved[3];
} rt_instance;

typedef struct rt_scene_desc {
  int32_t n_spheres;   const rt_sphere *spheres;
  int32_t n_triangles; const rt_triangle *triangles;
  int32_t n_quads;     const rt_quad *quads;
  int32_t n_materials; const rt_material *materials;
  rt_camera camera;
  float background[3]; /* profile 2: constant background (main.cu:96) */
  int32_t sky_gradient;/* 1: lerp(white,(.5,.7,1)) miss shader (main.cpp:80-82) */
  float t_min;         /* 1e-3 (main.cpp:65) / 1e-5 (obj_render.cu:33) */
  int32_t max_depth;   /* 50 */
  uint32_t flags;      /* RT_FLAG_FLIP_NORMALS | RT_FLAG_DEPTH_BACKGROUND */
  uint32_t reserved;
  /* version 2: participating media and procedural / image textures (profile 2 only) */
  int32_t n_media;     const rt_medium *media;
  int32_t n_perlin;    const rt_perlin *perlin;
  int32_t n_images;    const rt_image *images;
  /* version 5: instancing (profile 2 only) */
  int32_t n_groups;    const rt_group *groups;
  int32_t n_instances; const rt_instance *instances;
} rt_scene_desc;

/* 32-byte packed BVH node as downloaded by rt_accel_download (device layout).
 * Threaded (stackless) traversal: on a box hit go to `child` (inner) or test
 * the leaf's primitives and go to `escape`; on a miss go to `escape`.
 * index == n_nodes terminates. */
typedef struct rt_bvh_node {
  float bmin[3];
  int32_t escape;
  float bmax[3];
  int32_t payload; /* >= 0: index of first child (inner);
                    * < 0: leaf, ~payload = first << 3 | (count - 1): the leaf holds
                    * leaf_prims[first .. first + count), count <= 8 */
} rt_bvh_node;

typedef struct rt_stats_t {
  uint64_t paths;          /* camera rays started since the last rt_stats_reset */
  uint64_t segments;       /* closest-hit queries ("path-bounces") */
  uint64_t box_tests;      /* only with RT_FLAG_COUNTERS */
  uint64_t prim_tests;     /* only with RT_FLAG_COUNTERS */
  uint64_t kernel_launches;/* kernels of this library launched */
  float ms_upload, ms_build, ms_render, ms_resolve; /* last call of each phase (CUDA events) */
  int32_t n_nodes, n_big_prims, smem_bytes, block_threads, grid_blocks, regs_per_thread;
  int32_t smem_plan;       /* residency plan of the last render / trace launch: 2 = scene + eight node orderings in
                            * shared memory, 1 = scene + one ordering, 3 = nodes only, 0 = everything through L1/L2 */
  int32_t reserved;
  float ms_k_render, ms_k_combine; /* version 4: mean CUDA-event duration of k_render / k_combine over the (last <= 64)
                                    * render launches since rt_stats_reset, on whichever stream they ran */
} rt_stats_t;

int rt_version(void);
int rt_device_count(void);

int rt_create(rt_ctx **out, const rt_config *cfg);
void rt_destroy(rt_ctx *ctx);
const char *rt_last_error(const rt_ctx *ctx);

int rt_scene_upload(rt_ctx *ctx, const rt_scene_desc *scene);

/* quality: 0 = LBVH (Morton + radix sort + Karras), 1 = + SAH refinement. */
int rt_accel_build(rt_ctx *ctx, int quality);
/* nodes: threaded BVH; leaf_prims: RT_PRIM_IDs in depth-first leaf order; big_prims: the
 * always-tested oversized primitives. Any output pointer may be NULL (counts only).
 * Two-level scenes: the node array holds every tree - the top level first (root = node 0; its leaves may be
 * RT_PRIM_INSTANCE ids), then one slot per group -, n_nodes counts the slots' records and records a tree does not
 * use are undefined; leaves may also carry RT_PRIM_BOX ids (six rects of a box). */
int rt_accel_download(rt_ctx *ctx, rt_bvh_node *nodes, int cap_nodes, int *n_nodes, int32_t *leaf_prims,
                      int cap_leaf, int *n_leaf, int32_t *big_prims, int cap_big, int *n_big);

/* rays: [n][8] floats = origin.xyz, time, direction.xyz, unused. use_accel 0 =
 * brute force over the flattened arrays, 1 = through the BVH with a plain one-ray-per-thread loop,
 * 2 = through the BVH with the RENDER KERNEL's own machinery (the TRACE instantiation of k_render: the
 * same shared-memory residency plan and quadrant node orderings, warp-voted search bursts, hardware
 * reciprocals in the ray setup, direct leaf payloads and regeneration by ballot rank as rt_render),
 * so that the closest-hit parity tests cover the code that is timed. */
int rt_trace_closest(rt_ctx *ctx, const float *rays, int n, float t_min, float t_max, int use_accel,
                     int32_t *prim_id, float *t);
/* Same for instanced scenes: inst_id[k] = the rt_instance whose object holds the hit primitive, -1 for a
 * world-level primitive or a miss (may be NULL). */
int rt_trace_closest_inst(rt_ctx *ctx, const float *rays, int n, float t_min, float t_max, int use_accel,
                          int32_t *prim_id, int32_t *inst_id, float *t);

/* Transform-only update of an instanced scene (animation): new matrices for ALL n == n_instances instances
 * (group indices must not change). Only the top level is rebuilt - instance boxes + one tree over
 * (world primitives, instances); the objects' own trees and every primitive array stay as they are on the device.
 * The reference has no counterpart: its transforms are baked into a device-side object graph that
 * create_world<<<1,1>>> must rebuild from scratch (rt_next_week/cuda/main.cu:386-467). */
int rt_instances_update(rt_ctx *ctx, const rt_instance *instances, int n);

/* Accumulate samples [spp_begin, spp_begin+spp_count) of every pixel into the
 * context's accumulation buffer (float4 per pixel: sum R,G,B and sample count).
 * Counter-based RNG: the result for a given (pixel, sample) does not depend on
 * which call / which GPU rendered it. */
int rt_render(rt_ctx *ctx, int width, int height, int spp_begin, int spp_count);
/* Same, into a caller-provided DEVICE buffer (float[H*W*4]) on a caller stream
 * (cudaStream_t as void*; NULL = the context's stream). No host sync.
 * ONE render in flight per context: the work counter, the per-chunk partial frames and the statistics words
 * are per-context scratch, so two un-synchronised renders of the same context on different streams race.
 * Use one context per concurrent render (contexts are cheap: the scene of config 2 is 60 KB). */
int rt_render_device(rt_ctx *ctx, int width, int height, int spp_begin, int spp_count, float *d_accum,
                     void *stream);
/* Image-space split (SURVEY.md 8e alternative / 8f: tile split for interactive latency): as rt_render_device,
 * but only the rows [row_begin, row_end) of the frame (row 0 = bottom) are rendered and accumulated; every
 * other pixel of d_accum is left untouched. A pixel's samples do not depend on the split. */
int rt_render_rows_device(rt_ctx *ctx, int width, int height, int row_begin, int row_end, int spp_begin,
                          int spp_count, float *d_accum, void *stream);
int rt_accum_clear(rt_ctx *ctx);
int rt_accum_download(rt_ctx *ctx, float *rgba, size_t n_floats);     /* host <- device */
int rt_accum_upload(rt_ctx *ctx, int width, int height, const float *rgba, size_t n_floats); /* resume */
void *rt_accum_device_ptr(rt_ctx *ctx);

/* Resolve = /count, sqrt (gamma 2), clamp, quantise — per profile:
 * 0: int(256*clamp(sqrt(x),0,0.999)) color.h:21-27; 1,2: int(255.99*sqrt(x)) clamped to 255.
 * linear_rgb: [H][W][3] mean radiance (bottom row first) or NULL;
 * rgb8: [H][W][3] bytes, TOP row first (PPM order) or NULL. */
int rt_resolve(rt_ctx *ctx, float *linear_rgb, uint8_t *rgb8);
int rt_resolve_device(rt_ctx *ctx, int width, int height, const float *d_accum, float *linear_rgb,
                      uint8_t *rgb8, void *stream);

/* Denoiser feature buffers ("denoise-ready AOVs"; the reference renders one shot and has
 * none). Per pixel, the mean over `spp` camera samples - sample k of a pixel is the same primary ray
 * rt_render traces - of the first SURFACE hit's albedo (the texture colour the integrator multiplies by; 1 for
 * dielectrics; emission clamped to 1; the miss colour for rays that leave the scene), shading normal (facing
 * the ray where the profile flips normals; 0 on a miss) and ray parameter t (0 on a miss).
 * aov: host float[H][W][8] = albedo.rgb, normal.xyz, t, hit fraction; bottom row first. */
int rt_render_aov(rt_ctx *ctx, int width, int height, int spp, float *aov);

int rt_stats(rt_ctx *ctx, rt_stats_t *out);
int rt_stats_reset(rt_ctx *ctx);
int rt_sync(rt_ctx *ctx);

/* Page-locked host memory for the buffers the caller hands to rt_resolve / rt_accum_download /
 * rt_scene_upload: a device <-> host copy of pinned memory runs at link speed without the driver's staging
 * copy (2.9 MB rgb8 frame: 0.12 ms instead of ~0.4 ms). Any host pointer works everywhere; pinned ones are
 * faster. rt_host_free(NULL) is a no-op. */
void *rt_host_alloc(rt_ctx *ctx, size_t bytes);
void rt_host_free(rt_ctx *ctx, void *p);

/* ---- multi-GPU: the one exchange step of the path (north_star (4): "per-GPU sample
 * buffers combined by an NCCL reduce"). One context per GPU, each renders its share of the samples
 * (rt_render / rt_render_device with its own [spp_begin, spp_begin+spp_count)); rt_reduce sums the frames
 * onto `root`. Replaces the std::thread join + shared image array of main.cpp:292-355.
 *
 * One PROCESS per GPU (torchrun, MPI): rank 0 calls rt_comm_unique_id, ships the 128 bytes to the other
 * ranks by its own means, every rank calls rt_comm_init.
 * One process, several GPUs (apps/render_cli --gpus N): rt_comm_init_all on the array of contexts (one host
 * thread per context must then call rt_reduce concurrently, as NCCL requires of a single-process group).
 * NCCL is loaded at the first comm call (dlopen libnccl.so.2); without it these return RT_ERR_STATE. */
#define RT_COMM_ID_BYTES 128
int rt_comm_unique_id(void *id128);
int rt_comm_init(rt_ctx *ctx, const void *id128, int rank, int nranks);
int rt_comm_init_all(rt_ctx **ctxs, int n);
/* flags of rt_reduce */
#define RT_REDUCE_UNIFORM_COUNT 1u /* every pixel of a rank's frame holds the same sample count (the sample
                                    * split): only R,G,B travel (+ one count word), 25 % fewer bytes */
/* Sum-reduce the float4 frames of all ranks onto `root`'s d_accum (NULL = the context's own accumulation
 * buffer), asynchronously on `stream` (NULL = the context's stream): pack -> ncclReduce(sum, fp32) -> unpack
 * on root. Non-root frames are left unchanged. width*height must match on all ranks. */
int rt_reduce(rt_ctx *ctx, int width, int height, float *d_accum, int root, unsigned flags, void *stream);

/* Diagnostic: measured FP32 FMA throughput of the device (TFLOP/s, FMA = 2 flops), the
 * roofline denominator for this compute-bound path. */
int rt_measure_fp32_peak(rt_ctx *ctx, float *tflops);

#ifdef __cplusplus
}
#endif
#endif /* RT_CAPI_H */
