// b200rt.cu — the C ABI (include/rt_capi.h) over the sm_100a kernels.
// Single translation unit: nvcc -gencode arch=compute_100a,code=sm_100a (Makefile).
// There is NO CPU fallback in this library: without a CUDA device rt_create fails
// with RT_ERR_NODEVICE.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <dlfcn.h>
#include <nccl.h> // types and prototypes only: the library is loaded with dlopen at the first rt_comm_* call

#include "bvh_build.cuh"
#include "scene_flatten.h"
#include "render_kernels.cuh"

#define RT_MAX_BIG 32
#define RT_BIG_FRAC 0.30f
#define RT_BIG_ROUNDS 3

// ------------------------------------------------------------------ build kernels
__global__ void k_prim_box(BuildArrays B) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B.n_prims) body_prim_box(B, i);
}
__global__ void k_classify(BuildArrays B, int round, float frac) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B.n_prims) body_classify(B, i, round, frac);
}
// Oversized primitives as a (count, indices) list: read back through pinned memory (<= RT_MAX_BIG + 1 words,
// the only device -> host traffic of the classification step).
__global__ void k_big_collect(BuildArrays B, int *list) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B.n_prims && B.big_flag[i] == 1) {
    const int slot = atomicAdd(list, 1);
    if (slot < RT_MAX_BIG) list[1 + slot] = i;
  }
}
struct BigList {
  int n;
  int gid[RT_MAX_BIG]; // ascending
};
// The remaining primitives in ascending order: position = index - (big primitives before it), no scan needed
// because the big list is tiny. Also the per-sphere "is big" bytes of the brute-force parity hook.
__global__ void k_small_list(BuildArrays B, BigList bl, int *small_gid, uint8_t *sph_is_big) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B.n_prims) return;
  int before = 0;
  bool big = false;
  for (int k = 0; k < bl.n; k++) { before += bl.gid[k] < i ? 1 : 0; big = big || bl.gid[k] == i; }
  if (i < B.n_spheres) sph_is_big[i] = big ? 1 : 0;
  if (!big && small_gid) small_gid[i - before] = i;
}
__global__ void k_morton(BuildArrays B, int final_round) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B.n_pad) body_morton(B, i, final_round);
}
__global__ void k_bitonic(unsigned long long *a, int n_pad, int j, int k) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_pad) body_bitonic(a, i, j, k);
}
// all (k, j) steps with j < tile inside one block's shared memory
__global__ void k_bitonic_smem(unsigned long long *a, int n_pad, int k_begin, int k_end, int j_begin) {
  extern __shared__ unsigned long long s_keys[];
  const int tile = blockDim.x * 2;
  const int base = blockIdx.x * tile;
  for (int i = threadIdx.x; i < tile; i += blockDim.x) s_keys[i] = a[base + i];
  __syncthreads();
  for (int k = k_begin; k <= k_end; k <<= 1) {
    for (int j = (k == k_begin ? j_begin : k >> 1); j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < tile; i += blockDim.x) {
        int gi = base + i, ixj = i ^ j;
        if (ixj > i) {
          unsigned long long x = s_keys[i], y = s_keys[ixj];
          bool up = (gi & k) == 0;
          if ((x > y) == up) { s_keys[i] = y; s_keys[ixj] = x; }
        }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < tile; i += blockDim.x) a[base + i] = s_keys[i];
}
// ---- LSD radix sort of (30-bit Morton code, primitive) pairs: 8 passes of 4-bit digits,
// stable, multi-block. Per pass: per-block digit histograms -> exclusive scan over
// [digit][block] -> stable scatter (rank inside a 256-item round from __match_any_sync,
// across warps from a 16 x 8 shared table, across rounds from running per-digit bases).
// Input pairs arrive ordered by primitive index, so the stable result equals sorting the
// unique 64-bit keys code << 32 | primitive (what the bitonic fallback sorts).
#define RSORT_THREADS 256
#define RSORT_ITEMS 16
#define RSORT_TILE (RSORT_THREADS * RSORT_ITEMS)
__global__ void k_radix_split(const unsigned long long *keys64, unsigned *codes, int *gids, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { codes[i] = (unsigned)(keys64[i] >> 32); gids[i] = (int)(keys64[i] & 0xffffffffull); }
}
__global__ void k_radix_join(const unsigned *codes, const int *gids, unsigned long long *keys64, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keys64[i] = ((unsigned long long)codes[i] << 32) | (unsigned)gids[i];
}
__global__ void __launch_bounds__(RSORT_THREADS) k_radix_hist(const unsigned *keys, int n, int shift, int *hist, int nblocks) {
  __shared__ int cnt[16];
  if (threadIdx.x < 16) cnt[threadIdx.x] = 0;
  __syncthreads();
  const int base = blockIdx.x * RSORT_TILE;
  for (int r = 0; r < RSORT_ITEMS; r++) {
    int i = base + r * RSORT_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&cnt[(keys[i] >> shift) & 15u], 1);
  }
  __syncthreads();
  if (threadIdx.x < 16) hist[threadIdx.x * nblocks + blockIdx.x] = cnt[threadIdx.x];
}
__global__ void __launch_bounds__(1024) k_radix_scan(int *hist, int total) { // in-place exclusive scan, one block
  __shared__ int part[1024];
  const int per = (total + 1023) / 1024;
  const int b = threadIdx.x * per, e = min(b + per, total);
  int s = 0;
  for (int i = b; i < e; i++) s += hist[i];
  part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) { int acc = 0; for (int i = 0; i < 1024; i++) { int t = part[i]; part[i] = acc; acc += t; } }
  __syncthreads();
  int acc = part[threadIdx.x];
  for (int i = b; i < e; i++) { int t = hist[i]; hist[i] = acc; acc += t; }
}
__global__ void __launch_bounds__(RSORT_THREADS) k_radix_scatter(const unsigned *keys_in, const int *vals_in, unsigned *keys_out,
                                                             int *vals_out, int n, int shift, const int *offs, int nblocks) {
  __shared__ int run[16], tot[16], wcnt[RSORT_THREADS / 32][17];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < 16) run[threadIdx.x] = offs[threadIdx.x * nblocks + blockIdx.x];
  const int base = blockIdx.x * RSORT_TILE;
  for (int r = 0; r < RSORT_ITEMS; r++) {
    if (threadIdx.x < (RSORT_THREADS / 32) * 17) (&wcnt[0][0])[threadIdx.x] = 0;
    __syncthreads();
    const int i = base + r * RSORT_THREADS + threadIdx.x;
    const bool valid = i < n;
    unsigned key = 0;
    int val = 0, d = 16;
    if (valid) { key = keys_in[i]; val = vals_in[i]; d = (int)((key >> shift) & 15u); }
    const unsigned same = __match_any_sync(0xffffffffu, d);
    const int rank = __popc(same & ((1u << lane) - 1u));
    if (rank == 0) wcnt[warp][d] = __popc(same);
    __syncthreads();
    if (threadIdx.x < 16) {
      int acc = 0;
      for (int w = 0; w < RSORT_THREADS / 32; w++) { int t = wcnt[w][threadIdx.x]; wcnt[w][threadIdx.x] = acc; acc += t; }
      tot[threadIdx.x] = acc;
    }
    __syncthreads();
    if (valid) {
      const int pos = run[d] + wcnt[warp][d] + rank;
      keys_out[pos] = key;
      vals_out[pos] = val;
    }
    __syncthreads();
    if (threadIdx.x < 16) run[threadIdx.x] += tot[threadIdx.x];
  }
}
__global__ void k_karras(BuildArrays B) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B.n_small - 1) body_karras(B, i);
}
__global__ void k_fit(BuildArrays B, int rotate) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B.n_small) body_fit(B, i, rotate);
}
__global__ void k_order(BuildArrays B) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B.n_small - 1) body_order(B, i);
}
__global__ void k_pack(BuildArrays B) { // blockIdx.y = ordering (ray-direction octant)
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 2 * B.n_small - 1) body_pack(B, i, (int)blockIdx.y);
}

// ------------------------------------------------------------------ context
#define RT_N_BUILD_TEMPS 24
#define RT_EV_RING 64
struct DevBuf {
  void *p = nullptr;
  size_t bytes = 0;
};

// One build unit = one tree: the whole scene, the top level of an instanced scene (world primitives + instances)
// or one group (an object in its own space). All units share the builder kernels, the grow-only temporaries of the
// context and the packed node / leaf_prims arrays, where each unit owns a slot.
struct BuildUnit {
  int first[3] = {0, 0, 0}, count[3] = {0, 0, 0}; // sub-ranges of the scene's sphere / triangle / quad arrays
  int n_inst = 0;                     // instances (top level only): builder gids after the geometry
  bool top_level = false;             // the top level of a two-level scene: geometry owned by groups is not its business
  bool classify = true;               // split oversized primitives off into the always-tested lists
  int link_base = 0, leaf_base = 0, end_link = -1;
  int stride_nodes = -1;              // nodes per ordering of the packed array; -1: this unit alone (2 n_small - 1)
  // results
  int n_small = 0, n_big = 0, kept_nodes = 0;
  float root_box[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};

struct rt_ctx {
  rt_config cfg;
  std::string err;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // render launches since the last rt_stats_reset: a ring of event triples (before k_render, after it, after
  // k_combine) so that rt_stats can report the MEAN kernel durations of asynchronously queued frames
  cudaEvent_t ev_k[RT_EV_RING][3];
  int ev_k_count = 0; // launches recorded since the reset (ring slot = count % RT_EV_RING)
  cudaStream_t last_render_stream = nullptr; // rt_stats waits for it: renders on caller streams update the device counters
  int sm_count = 0, max_smem_optin = 0;
  // host copy of the scene
  bool have_scene = false, have_accel = false;
  std::vector<rt_sphere> spheres;
  std::vector<rt_triangle> tris;
  std::vector<rt_quad> quads;
  std::vector<rt_material> mats;
  rt_scene_desc desc;
  bool general = false;
  bool ext = false; // media or noise/image textures: the extended kernel variant
  bool spheres_only = false; // profile 2 with nothing but spheres (and no media / instances): the SPH instantiations
  // device scene
  // d_scene: ONE arena holding every flattened array of the scene (16-byte aligned slots, zero padded), filled by
  // one copy from the pinned staging block h_stage; the d_raw_* pointers (builder input) point into it too
  DevBuf d_scene, d_nodes, d_big, d_leaf_prims, d_bigq, d_sph_is_big, d_raw_sph, d_raw_tri, d_raw_quad;
  DevBuf build_tmp[RT_N_BUILD_TEMPS]; // rt_accel_build temporaries: grow-only, never freed between builds
  ncclComm_t comm = nullptr; // rt_comm_init / rt_comm_init_all
  int comm_rank = 0, comm_size = 1;
  DevBuf d_red;              // packed frame travelling through ncclReduce
  void *h_pin = nullptr, *h_stage = nullptr, *h_out = nullptr; // pinned host blocks: small read-backs, scene staging, frame read-back
  size_t h_pin_bytes = 0, h_stage_bytes = 0, h_out_bytes = 0;
  DevScene S;
  DevCamera cam;
  ShadeParams sp;
  std::vector<int32_t> big_ids;
  // two-level scenes (rt_group / rt_instance)
  bool two_level = false;
  std::vector<rt_group> groups;
  std::vector<rt_instance> instances;
  std::vector<uint8_t> grouped;      // per primitive (spheres, triangles, quads): owned by a group
  std::vector<BuildUnit> group_units; // layout + root box of every group's tree
  BuildUnit top_unit;
  std::vector<int> group_root_off, h_small;
  std::vector<float> group_frames; // per group {qbase[3], qscale[3]}
  std::vector<int> box_first_quad; // boxes recognised by scene_flatten.h: first of the six rects
  std::vector<uint8_t> h_mask;     // build_unit scratch
  std::vector<int> h_unit_boxes;
  std::vector<float4> h_units;     // unit table of the shared-memory staging loop (render_kernels.cuh)
  DevBuf d_units;
  std::vector<float4> h_inst_rec, h_inst_box;
  DevBuf d_exclude, d_inst_lo, d_inst_hi, d_inst_ids; // d_exclude: a view into d_scene
  int total_nodes = 0, build_quality = 1;
  int n_leaf_prims = 0, max_leaf = 1;
  float root_box[8] = {0, 0, 0, 0, 0, 0, 0, 0}; // {bmin.xyz, -, bmax.xyz, -} of the BVH root
  // frame
  int W = 0, H = 0;
  DevBuf d_accum, d_partial, d_counter, d_stats, d_linear, d_rgb8, d_rays, d_ids, d_ts;
  rt_stats_t stats;
  unsigned long long launches = 0;
};

static int fail(rt_ctx *c, int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (c) c->err = buf;
  return code;
}
#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(ctx, RT_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

static size_t pad16(size_t b) { return (b + 15) & ~(size_t)15; }

static int dev_reserve(rt_ctx *ctx, DevBuf &b, size_t bytes) {
  bytes = pad16(std::max<size_t>(bytes, 16));
  if (b.bytes >= bytes) return RT_OK;
  if (b.p) CK(cudaFree(b.p));
  b.p = nullptr; b.bytes = 0;
  CK(cudaMalloc(&b.p, bytes));
  b.bytes = bytes;
  return RT_OK;
}
static void dev_free(DevBuf &b) {
  if (b.p) cudaFree(b.p);
  b.p = nullptr; b.bytes = 0;
}
// grow-only pinned host block (contents are NOT preserved when it grows)
static int host_reserve(rt_ctx *ctx, void *&p, size_t &have, size_t bytes) {
  if (have >= bytes) return RT_OK;
  if (p) CK(cudaFreeHost(p));
  p = nullptr; have = 0;
  bytes = (bytes + 4095) & ~(size_t)4095;
  CK(cudaHostAlloc(&p, bytes, cudaHostAllocDefault));
  have = bytes;
  return RT_OK;
}
static int pin_reserve(rt_ctx *ctx, size_t bytes) { return host_reserve(ctx, ctx->h_pin, ctx->h_pin_bytes, bytes); }

// ------------------------------------------------------------------ NCCL (loaded on demand)
struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Reduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  std::string why;
};
static NcclApi &nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api;
  tried = true;
  // a copy already mapped into the process (torch's) wins, so that one process never holds two NCCLs
  const char *names[] = {getenv("B200RT_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char *nm : names) {
    if (!nm || !*nm) continue;
    api.handle = dlopen(nm, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (api.handle) break;
  }
  for (const char *nm : names) {
    if (api.handle) break;
    if (!nm || !*nm) continue;
    api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
  }
  if (!api.handle) { api.why = "NCCL not found (libnccl.so.2; set B200RT_NCCL_LIB)"; return api; }
#define SYM(field, name) \
  *(void **)(&api.field) = dlsym(api.handle, name); \
  if (!api.field) { api.why = std::string("NCCL symbol missing: ") + name; api.handle = nullptr; return api; }
  SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommInitAll, "ncclCommInitAll")
  SYM(CommDestroy, "ncclCommDestroy") SYM(Reduce, "ncclReduce") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  return api;
}
static void nccl_comm_destroy(ncclComm_t c) {
  NcclApi &a = nccl_api();
  if (a.handle && c) a.CommDestroy(c);
}

// frame <-> travelling buffer of rt_reduce: rgb only (+ one count word at the end) or all four lanes
__global__ void k_reduce_pack(const float4 *__restrict__ accum, float *__restrict__ red, int n_pix) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pix) return;
  const float4 a = accum[p];
  red[3 * (size_t)p + 0] = a.x; red[3 * (size_t)p + 1] = a.y; red[3 * (size_t)p + 2] = a.z;
  if (p == 0) red[3 * (size_t)n_pix] = a.w;
}
__global__ void k_reduce_unpack(float4 *__restrict__ accum, const float *__restrict__ red, int n_pix) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pix) return;
  accum[p] = make_float4(red[3 * (size_t)p + 0], red[3 * (size_t)p + 1], red[3 * (size_t)p + 2], red[3 * (size_t)n_pix]);
}

extern "C" {

int rt_version(void) { return RT_CAPI_VERSION; }

int rt_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

const char *rt_last_error(const rt_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int rt_create(rt_ctx **out, const rt_config *cfg) {
  if (!out || !cfg) return RT_ERR_INVALID;
  *out = nullptr;
  if (cfg->profile < 0 || cfg->profile > 2) return RT_ERR_INVALID;
  int n = rt_device_count();
  if (n <= 0) return RT_ERR_NODEVICE;
  if (cfg->device < 0 || cfg->device >= n) return RT_ERR_INVALID;
  rt_ctx *ctx = new (std::nothrow) rt_ctx();
  if (!ctx) return RT_ERR_NOMEM;
  ctx->cfg = *cfg;
  memset(&ctx->stats, 0, sizeof ctx->stats);
  memset(&ctx->S, 0, sizeof ctx->S);
  cudaError_t e = cudaSetDevice(cfg->device);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev0);
  if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev1);
  memset(ctx->ev_k, 0, sizeof ctx->ev_k);
  for (int i = 0; i < RT_EV_RING * 3 && e == cudaSuccess; i++) e = cudaEventCreate(&ctx->ev_k[i / 3][i % 3]);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, cfg->device);
  if (e == cudaSuccess)
    e = cudaDeviceGetAttribute(&ctx->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, cfg->device);
  if (e != cudaSuccess) {
    delete ctx;
    return RT_ERR_CUDA;
  }
  *out = ctx;
  return RT_OK;
}

void rt_destroy(rt_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->cfg.device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  // d_raw_* are views into d_scene
  if (ctx->comm) nccl_comm_destroy(ctx->comm);
  DevBuf *all[] = {&ctx->d_red, &ctx->d_scene, &ctx->d_nodes, &ctx->d_big, &ctx->d_leaf_prims, &ctx->d_bigq, &ctx->d_sph_is_big, &ctx->d_accum,
                   &ctx->d_partial, &ctx->d_counter, &ctx->d_stats, &ctx->d_linear, &ctx->d_rgb8, &ctx->d_rays, &ctx->d_ids, &ctx->d_ts,
                   &ctx->d_inst_lo, &ctx->d_inst_hi, &ctx->d_inst_ids, &ctx->d_units};
  for (DevBuf *b : all) dev_free(*b);
  for (DevBuf &b : ctx->build_tmp) dev_free(b);
  if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
  if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
  if (ctx->h_out) cudaFreeHost(ctx->h_out);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  for (int i = 0; i < RT_EV_RING * 3; i++) if (ctx->ev_k[i / 3][i % 3]) cudaEventDestroy(ctx->ev_k[i / 3][i % 3]);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

int rt_sync(rt_ctx *ctx) {
  if (!ctx) return RT_ERR_INVALID;
  CK(cudaSetDevice(ctx->cfg.device));
  CK(cudaStreamSynchronize(ctx->stream));
  return RT_OK;
}

// Flatten the scene description into the device SoA layout (rt_common.cuh DevScene).
int rt_scene_upload(rt_ctx *ctx, const rt_scene_desc *sc) {
  if (!ctx || !sc) return RT_ERR_INVALID;
  CK(cudaSetDevice(ctx->cfg.device));
  HostFlat F;
  {
    int frc = flatten_scene(sc, ctx->cfg.profile, F, ctx->err);
    if (frc) return frc;
  }
  const int profile = ctx->cfg.profile;
  const bool any_moving = F.any_moving;
  ctx->spheres.assign(sc->spheres, sc->spheres + sc->n_spheres);
  ctx->tris.assign(sc->triangles, sc->triangles + sc->n_triangles);
  ctx->quads.assign(sc->quads, sc->quads + sc->n_quads);
  ctx->mats.assign(sc->materials, sc->materials + sc->n_materials);
  ctx->desc = *sc;
  ctx->desc.spheres = ctx->spheres.data();
  ctx->desc.triangles = ctx->tris.data();
  ctx->desc.quads = ctx->quads.data();
  ctx->desc.materials = ctx->mats.data();
  ctx->desc.media = nullptr; ctx->desc.perlin = nullptr; ctx->desc.images = nullptr; // flattened below, not kept
  ctx->general = (profile == RT_PROFILE_NEXT_WEEK);
  ctx->ext = sc->n_media > 0;
  for (int i = 0; i < sc->n_materials; i++)
    if (sc->materials[i].texture >= RT_TEX_NOISE || sc->materials[i].type == RT_MAT_ISOTROPIC) ctx->ext = true;
  if (const char *e = getenv("B200RT_EXT")) ctx->ext = ctx->ext || (ctx->general && atoi(e) != 0); // measurement knob
  ctx->spheres_only = ctx->general && !ctx->ext && sc->n_triangles == 0 && sc->n_quads == 0 && sc->n_groups == 0 &&
                      sc->n_instances == 0;
  if (const char *e = getenv("B200RT_SPH")) ctx->spheres_only = ctx->spheres_only && atoi(e) != 0; // measurement knob

  CK(cudaEventRecord(ctx->ev0, ctx->stream));
  const int ns = sc->n_spheres, nt = sc->n_triangles, nq = sc->n_quads, nm = sc->n_materials;
  int rc;
  // ONE packed upload: every array gets a 16-byte aligned, zero-padded slot of the pinned staging block; one
  // cudaMemcpyAsync moves the block into the device arena (was: a memset + memcpy pair per array, 24 pairs)
  struct Slot { const void *src; size_t bytes, off; };
  std::vector<Slot> slots;
  size_t total = 0;
  auto add = [&](const void *src, size_t bytes) {
    slots.push_back({src, bytes, total});
    total += pad16(std::max<size_t>(bytes, 16));
    return slots.size() - 1;
  };
#define ADD(vec) add((vec).data(), (vec).size() * sizeof((vec)[0]))
  const size_t i_sph = ADD(F.sph), i_sph_k = ADD(F.sph_k), i_sph_mv = ADD(F.sph_mv), i_sph_t0 = ADD(F.sph_t0), i_tri = ADD(F.tri),
               i_tri_n = ADD(F.tri_n), i_quad = ADD(F.quad), i_sph_mat = ADD(F.sph_mat), i_tri_mat = ADD(F.tri_mat),
               i_quad_mat = ADD(F.quad_mat), i_mats = ADD(F.mats), i_raw_sph = ADD(ctx->spheres), i_raw_tri = ADD(ctx->tris),
               i_raw_quad = ADD(ctx->quads), i_media = ADD(F.media), i_pvec = ADD(F.perlin_vec), i_pperm = ADD(F.perlin_perm),
               i_img = ADD(F.image_bytes);
  std::vector<DevImage> dimg((size_t)sc->n_images);
  const size_t i_images = ADD(dimg); // filled below, once the arena's address is known
  std::vector<rt_group> groups_copy(sc->groups, sc->groups + sc->n_groups);
  const size_t i_inst = ADD(F.inst), i_excl = ADD(F.grouped), i_groups = ADD(groups_copy), i_box = ADD(F.box);
#undef ADD
  if ((rc = dev_reserve(ctx, ctx->d_scene, total))) return rc;
  if ((rc = host_reserve(ctx, ctx->h_stage, ctx->h_stage_bytes, total))) return rc;
  unsigned char *hs = (unsigned char *)ctx->h_stage, *ds = (unsigned char *)ctx->d_scene.p;
  for (int i = 0; i < sc->n_images; i++) {
    dimg[i].rgb = (const uint8_t *)(ds + slots[i_img].off) + F.image_offset[i];
    dimg[i].width = sc->images[i].width;
    dimg[i].height = sc->images[i].height;
  }
  slots[i_images].src = dimg.data();
  memset(hs, 0, total);
  for (const Slot &sl : slots)
    if (sl.bytes) memcpy(hs + sl.off, sl.src, sl.bytes);
  CK(cudaMemcpyAsync(ds, hs, total, cudaMemcpyHostToDevice, ctx->stream));
  if ((rc = dev_reserve(ctx, ctx->d_sph_is_big, (size_t)std::max(ns, 1)))) return rc;
  CK(cudaMemsetAsync(ctx->d_sph_is_big.p, 0, ctx->d_sph_is_big.bytes, ctx->stream));
  if ((rc = dev_reserve(ctx, ctx->d_big, 16))) return rc;
  if ((rc = dev_reserve(ctx, ctx->d_bigq, 32))) return rc;
  if ((rc = dev_reserve(ctx, ctx->d_leaf_prims, 16))) return rc;
  if ((rc = dev_reserve(ctx, ctx->d_nodes, 32))) return rc;
  ctx->d_raw_sph.p = ds + slots[i_raw_sph].off;
  ctx->d_raw_tri.p = ds + slots[i_raw_tri].off;
  ctx->d_raw_quad.p = ds + slots[i_raw_quad].off;

  DevScene &S = ctx->S;
  S.nodes = (const float4 *)ctx->d_nodes.p;
  S.sph = (const float4 *)(ds + slots[i_sph].off);
  S.sph_mv = (const float4 *)(ds + slots[i_sph_mv].off);
  S.sph_t0 = (const float *)(ds + slots[i_sph_t0].off);
  S.sph_k = (const float *)(ds + slots[i_sph_k].off);
  S.tri = (const float4 *)(ds + slots[i_tri].off);
  S.tri_n = (const float4 *)(ds + slots[i_tri_n].off);
  S.quad = (const float4 *)(ds + slots[i_quad].off);
  S.sph_mat = (const int32_t *)(ds + slots[i_sph_mat].off);
  S.tri_mat = (const int32_t *)(ds + slots[i_tri_mat].off);
  S.quad_mat = (const int32_t *)(ds + slots[i_quad_mat].off);
  S.mats = (const float4 *)(ds + slots[i_mats].off);
  S.big = (const int32_t *)ctx->d_big.p;
  S.leaf_prims = (const int32_t *)ctx->d_leaf_prims.p;
  S.media = (const float4 *)(ds + slots[i_media].off);
  S.perlin_vec = (const float4 *)(ds + slots[i_pvec].off);
  S.perlin_perm = (const uint8_t *)(ds + slots[i_pperm].off);
  S.images = (const DevImage *)(ds + slots[i_images].off);
  S.n_media = sc->n_media; S.n_perlin = sc->n_perlin; S.n_images = sc->n_images;
  S.inst = (const float4 *)(ds + slots[i_inst].off);
  S.n_inst = sc->n_instances;
  S.groups = (const int32_t *)(ds + slots[i_groups].off);
  S.box = (const float4 *)(ds + slots[i_box].off);
  S.n_boxes = (int)F.box.size() / 2;
  ctx->box_first_quad.clear();
  for (int b = 0; b < S.n_boxes; b++) ctx->box_first_quad.push_back(RT_F2I(F.box[2 * (size_t)b].w));
  ctx->d_exclude.p = ds + slots[i_excl].off; // (a view into the arena, like d_raw_*: never freed on its own)
  ctx->two_level = sc->n_groups > 0 || sc->n_instances > 0;
  ctx->groups.assign(sc->groups, sc->groups + sc->n_groups);
  ctx->instances.assign(sc->instances, sc->instances + sc->n_instances);
  ctx->grouped = F.grouped;
  ctx->group_units.clear();

  S.bigq = (const float4 *)ctx->d_bigq.p;
  S.n_nodes = 0; S.n_big = 0; S.n_bigq = 0;
  S.n_spheres = ns; S.n_tris = nt; S.n_quads = nq; S.n_mats = nm;
  S.any_moving = any_moving ? 1 : 0;

  const rt_camera &c = sc->camera;
  ctx->cam.origin = v3_from(c.origin);
  ctx->cam.llc = v3_from(c.lower_left_corner);
  ctx->cam.horizontal = v3_from(c.horizontal);
  ctx->cam.vertical = v3_from(c.vertical);
  ctx->cam.u = v3_from(c.u);
  ctx->cam.v = v3_from(c.v);
  ctx->cam.lens_radius = c.lens_radius;
  ctx->cam.time0 = c.time0;
  ctx->cam.time1 = c.time1;
  ctx->sp.background = v3_from(sc->background);
  ctx->sp.sky_gradient = sc->sky_gradient;
  ctx->sp.flags = sc->flags | ctx->cfg.flags;
  ctx->sp.t_min = sc->t_min;
  ctx->sp.max_depth = sc->max_depth;
  CK(cudaEventRecord(ctx->ev1, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaEventElapsedTime(&ctx->stats.ms_upload, ctx->ev0, ctx->ev1));
  ctx->have_scene = true;
  ctx->have_accel = false;
  ctx->big_ids.clear();
  ctx->stats.n_nodes = 0;
  ctx->stats.n_big_prims = 0;
  return RT_OK;
}

static int build_unit(rt_ctx *ctx, BuildUnit &U, int quality) {
  cudaStream_t st = ctx->stream;
  const int ns = U.count[0], nt = U.count[1], nq = U.count[2];
  const int n_geom = ns + nt + nq;
  int n = n_geom + U.n_inst;
  U.n_small = 0; U.n_big = 0; U.kept_nodes = 0;
  memset(U.root_box, 0, sizeof U.root_box);
  if (n == 0) return RT_OK;

  BuildArrays B;
  memset(&B, 0, sizeof B);
  B.n_prims = n; B.n_spheres = ns; B.n_tris = nt; B.n_quads = nq; B.n_inst = U.n_inst;
  B.spheres = (const rt_sphere *)ctx->d_raw_sph.p + U.first[0];
  B.tris = (const rt_triangle *)ctx->d_raw_tri.p + U.first[1];
  B.quads = (const rt_quad *)ctx->d_raw_quad.p + U.first[2];
  for (int k = 0; k < 3; k++) B.id_base[k] = U.first[k];
  B.inst_lo = (const float4 *)ctx->d_inst_lo.p; B.inst_hi = (const float4 *)ctx->d_inst_hi.p;
  // Geometry that is NOT a leaf of this unit: primitives owned by groups (top level of a two-level scene) and the
  // sides of the unit's boxes, which get one leaf per box instead (RT_PRIM_BOX, gids after the instances)
  std::vector<uint8_t> &mask = ctx->h_mask;
  std::vector<int> &unit_boxes = ctx->h_unit_boxes;
  mask.assign((size_t)std::max(n_geom, 1), 0);
  unit_boxes.clear();
  bool any_mask = false;
  if (U.top_level)
    for (int i = 0; i < n_geom; i++) { mask[i] = ctx->grouped[i]; any_mask = any_mask || mask[i]; }
  for (int b = 0; b < (int)ctx->box_first_quad.size(); b++) {
    const int fq = ctx->box_first_quad[b];
    if (fq < U.first[2] || fq + 6 > U.first[2] + nq) continue;
    if (U.top_level && ctx->grouped[(size_t)ctx->S.n_spheres + ctx->S.n_tris + fq]) continue;
    unit_boxes.push_back(b);
    for (int k = 0; k < 6; k++) mask[(size_t)ns + nt + (fq - U.first[2]) + k] = 1;
    any_mask = true;
  }
  const int n_boxes = (int)unit_boxes.size();
  n += n_boxes;
  B.n_prims = n; B.n_boxes = n_boxes;
  B.box_rec = ctx->S.box;
  B.link_base = U.link_base; B.leaf_base = U.leaf_base; B.end_link = U.end_link;
  B.thickness = (ctx->sp.flags & RT_FLAG_FLIP_NORMALS) ? 0.01f : 0.1f; // THICKNESS of the two trees
  // Build temporaries live in the context and only ever grow: a rebuild (every frame of an animated or
  // re-uploaded scene) calls neither cudaMalloc nor cudaFree, which would synchronise the device.
  enum { T_LO, T_HI, T_FLAG, T_BOUNDS, T_BIGLIST, T_SMALL, T_KEYS, T_LEFT, T_RIGHT, T_PARENT, T_NFLAG, T_SIZE, T_LCNT, T_NLO,
         T_NHI, T_SWAP, T_RK0, T_RK1, T_RV0, T_RV1, T_RHIST, T_MASK, T_BOXIDS, T_COUNT };
  static_assert(T_COUNT <= RT_N_BUILD_TEMPS, "build temporaries");
  DevBuf *T = ctx->build_tmp;
  int rc;
#define RSV(buf, bytes) do { if ((rc = dev_reserve(ctx, buf, (bytes)))) return rc; } while (0)
  RSV(T[T_LO], sizeof(float4) * (size_t)n); RSV(T[T_HI], sizeof(float4) * (size_t)n); RSV(T[T_FLAG], sizeof(int) * (size_t)n);
  RSV(T[T_BOUNDS], sizeof(BuildBounds) * 4); RSV(T[T_BIGLIST], sizeof(int) * (RT_MAX_BIG + 1));
  if ((rc = pin_reserve(ctx, 4096))) return rc;
  B.pbox_lo = (float4 *)T[T_LO].p; B.pbox_hi = (float4 *)T[T_HI].p; B.big_flag = (int *)T[T_FLAG].p;
  B.bounds = (BuildBounds *)T[T_BOUNDS].p;
  if (any_mask) {
    RSV(T[T_MASK], mask.size()); RSV(T[T_BOXIDS], sizeof(int) * (size_t)std::max(n_boxes, 1));
    CK(cudaMemcpyAsync(T[T_MASK].p, mask.data(), mask.size(), cudaMemcpyHostToDevice, st));
    if (n_boxes) CK(cudaMemcpyAsync(T[T_BOXIDS].p, unit_boxes.data(), sizeof(int) * (size_t)n_boxes, cudaMemcpyHostToDevice, st));
    B.exclude = (const uint8_t *)T[T_MASK].p;
    B.box_ids = (const int *)T[T_BOXIDS].p;
  }
  int *h_words = (int *)ctx->h_pin; // pinned: [0..32] big list, [40] kept nodes, [48..55] root box
  {
    BuildBounds *init = (BuildBounds *)(h_words + 64); // 4 x 24 bytes, pinned
    for (int r = 0; r < 4; r++)
      for (int a = 0; a < 3; a++) { init[r].lo[a] = 0x7fffffff; init[r].hi[a] = (int)0x80000000; }
    CK(cudaMemcpyAsync(T[T_BOUNDS].p, init, sizeof(BuildBounds) * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(T[T_BIGLIST].p, 0, sizeof(int) * (RT_MAX_BIG + 1), st));
  }
  const int TB = 256;
  const int gp = (n + TB - 1) / TB;
  k_prim_box<<<gp, TB, 0, st>>>(B);
  ctx->launches++;
  BigList bl;
  memset(&bl, 0, sizeof bl);
  int final_round = 0;
  if (U.classify) {
    for (int r = 0; r < RT_BIG_ROUNDS; r++) k_classify<<<gp, TB, 0, st>>>(B, r, RT_BIG_FRAC);
    k_big_collect<<<gp, TB, 0, st>>>(B, (int *)T[T_BIGLIST].p);
    ctx->launches += 1 + RT_BIG_ROUNDS;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(h_words, T[T_BIGLIST].p, sizeof(int) * (RT_MAX_BIG + 1), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st)); // host round trip 1 of 2: how many primitives stay in the tree sizes every later launch
    final_round = RT_BIG_ROUNDS;
    if (h_words[0] > RT_MAX_BIG) { // degenerate classification: keep everything in the tree
      final_round = 0;
    } else {
      bl.n = h_words[0];
      std::copy(h_words + 1, h_words + 1 + bl.n, bl.gid);
      std::sort(bl.gid, bl.gid + bl.n);
    }
    ctx->big_ids.clear();
    for (int k = 0; k < bl.n; k++) {
      const int i = bl.gid[k];
      ctx->big_ids.push_back(i < ns ? RT_PRIM_ID(RT_PRIM_SPHERE, U.first[0] + i)
                                    : (i < ns + nt ? RT_PRIM_ID(RT_PRIM_TRIANGLE, U.first[1] + i - ns)
                                                   : RT_PRIM_ID(RT_PRIM_QUAD, U.first[2] + i - ns - nt)));
    }
    // rects go to their own decoded list (no id fetch / type dispatch in the segment start)
    std::vector<int32_t> others;
    std::vector<float4> bigq;
    for (int32_t id : ctx->big_ids) {
      if (RT_PRIM_TYPE_OF(id) != RT_PRIM_QUAD) { others.push_back(id); continue; }
      const rt_quad &q = ctx->quads[RT_PRIM_INDEX_OF(id)];
      bigq.push_back(make_float4(q.k, q.a0, q.a1, RT_I2F(q.axis)));
      bigq.push_back(make_float4(q.b0, q.b1, RT_I2F(id), 0.f));
    }
    const int n_bigq = (int)bigq.size() / 2;
    // both lists through ONE pinned block and one copy each (zero padded to the reserved size)
    const size_t big_bytes = pad16(std::max<size_t>(others.size(), 4) * sizeof(int32_t));
    const size_t bigq_bytes = std::max<size_t>(bigq.size(), 2) * sizeof(float4);
    if ((rc = pin_reserve(ctx, 1024 + big_bytes + bigq_bytes))) return rc;
    unsigned char *hb = (unsigned char *)ctx->h_pin + 1024;
    h_words = (int *)ctx->h_pin;
    memset(hb, 0, big_bytes + bigq_bytes);
    if (!others.empty()) memcpy(hb, others.data(), others.size() * sizeof(int32_t));
    if (!bigq.empty()) memcpy(hb + big_bytes, bigq.data(), bigq.size() * sizeof(float4));
    RSV(ctx->d_big, big_bytes); RSV(ctx->d_bigq, bigq_bytes);
    CK(cudaMemcpyAsync(ctx->d_big.p, hb, big_bytes, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_bigq.p, hb + big_bytes, bigq_bytes, cudaMemcpyHostToDevice, st));
    ctx->S.big = (const int32_t *)ctx->d_big.p;
    ctx->S.n_big = (int)others.size();
    ctx->S.bigq = (const float4 *)ctx->d_bigq.p;
    ctx->S.n_bigq = n_bigq;
  }
  const int n_big = bl.n;
  U.n_big = n_big;
  // the primitives that stay in the tree, ascending
  int nsm;
  if (any_mask) { // the host knows which primitives are not leaves of this unit
    std::vector<int> &small = ctx->h_small;
    small.clear();
    int kb = 0;
    for (int i = 0; i < n; i++) {
      while (kb < bl.n && bl.gid[kb] < i) kb++;
      if ((kb < bl.n && bl.gid[kb] == i) || (i < n_geom && mask[i])) continue;
      small.push_back(i);
    }
    nsm = (int)small.size();
    RSV(T[T_SMALL], sizeof(int) * (size_t)std::max(nsm, 1));
    if (nsm) CK(cudaMemcpyAsync(T[T_SMALL].p, small.data(), sizeof(int) * (size_t)nsm, cudaMemcpyHostToDevice, st));
    k_small_list<<<gp, TB, 0, st>>>(B, bl, nullptr, (uint8_t *)ctx->d_sph_is_big.p + U.first[0]);
  } else {
    nsm = n - n_big;
    RSV(T[T_SMALL], sizeof(int) * (size_t)std::max(nsm, 1));
    k_small_list<<<gp, TB, 0, st>>>(B, bl, (int *)T[T_SMALL].p, (uint8_t *)ctx->d_sph_is_big.p + U.first[0]);
  }
  ctx->launches++;
  U.n_small = nsm;
  int n_nodes = nsm > 0 ? 2 * nsm - 1 : 0;
  if (nsm > 0) {
    int n_pad = 1;
    while (n_pad < nsm) n_pad <<= 1;
    if (n_pad < 2) n_pad = 2;
    B.n_small = nsm; B.n_pad = n_pad;
    RSV(T[T_KEYS], sizeof(unsigned long long) * (size_t)n_pad);
    RSV(T[T_LEFT], sizeof(int) * (size_t)nsm); RSV(T[T_RIGHT], sizeof(int) * (size_t)nsm);
    RSV(T[T_PARENT], sizeof(int) * (size_t)n_nodes); RSV(T[T_NFLAG], sizeof(int) * (size_t)nsm);
    RSV(T[T_SIZE], sizeof(int) * (size_t)n_nodes); RSV(T[T_LCNT], sizeof(int) * (size_t)n_nodes);
    RSV(T[T_SWAP], sizeof(int) * (size_t)nsm);
    RSV(T[T_NLO], sizeof(float4) * (size_t)n_nodes); RSV(T[T_NHI], sizeof(float4) * (size_t)n_nodes);
    const int stride_nodes = U.stride_nodes >= 0 ? U.stride_nodes : n_nodes;
    if (U.stride_nodes < 0) { // a single-level scene: the arrays are exactly this unit's (multi-unit scenes reserve them up front)
      RSV(ctx->d_leaf_prims, sizeof(int32_t) * (size_t)nsm);
      // eight packed copies, one per ray-direction octant (own front-to-back visiting order)
      RSV(ctx->d_nodes, RT_N_ORDERINGS * sizeof(float4) * 2 * (size_t)n_nodes);
    }
    B.small_gid = (const int *)T[T_SMALL].p;
    B.keys = (unsigned long long *)T[T_KEYS].p;
    B.left = (int *)T[T_LEFT].p; B.right = (int *)T[T_RIGHT].p; B.parent = (int *)T[T_PARENT].p;
    B.flag = (int *)T[T_NFLAG].p; B.size = (int *)T[T_SIZE].p; B.lcnt = (int *)T[T_LCNT].p;
    B.leaf_prims = (int32_t *)ctx->d_leaf_prims.p;
    B.swapmask = (int *)T[T_SWAP].p;
    B.packed_stride = 2 * stride_nodes;
    {
      const char *e = getenv("B200RT_MAX_LEAF"); // tuning knob (leaf size)
      int ml = e ? atoi(e) : 1; // while-while traversal: single-primitive leaves measured fastest
      B.max_leaf = ml < 1 ? 1 : (ml > 8 ? 8 : ml);
      if (U.n_inst) B.max_leaf = 1; // an instance leaf holds exactly one instance (the traversal enters it)
      ctx->max_leaf = std::max(ctx->max_leaf, B.max_leaf);
    }
    B.nbox_lo = (float4 *)T[T_NLO].p; B.nbox_hi = (float4 *)T[T_NHI].p;
    B.packed = (float4 *)ctx->d_nodes.p;
    k_morton<<<(n_pad + TB - 1) / TB, TB, 0, st>>>(B, final_round);
    ctx->launches++;
    const char *sort_env = getenv("B200RT_SORT");
    const int SORT_THREADS = 512, TILE = 2 * SORT_THREADS;
    // small scenes: the whole sort inside one CTA's shared memory (one launch instead of 26); both sorts order
    // the same unique 64-bit keys, so the tree does not depend on which one ran (tests: radix == bitonic)
    const bool use_bitonic = sort_env ? !strcmp(sort_env, "bitonic") : n_pad <= 2 * TILE;
    if (!use_bitonic) {
      // radix sort of (code, primitive) pairs, 8 x 4-bit passes, ping-pong buffers
      const int nb = (nsm + RSORT_TILE - 1) / RSORT_TILE;
      RSV(T[T_RK0], sizeof(unsigned) * (size_t)nsm); RSV(T[T_RK1], sizeof(unsigned) * (size_t)nsm);
      RSV(T[T_RV0], sizeof(int) * (size_t)nsm); RSV(T[T_RV1], sizeof(int) * (size_t)nsm);
      RSV(T[T_RHIST], sizeof(int) * 16 * (size_t)nb);
      unsigned *k0 = (unsigned *)T[T_RK0].p, *k1 = (unsigned *)T[T_RK1].p;
      int *v0 = (int *)T[T_RV0].p, *v1 = (int *)T[T_RV1].p, *hist = (int *)T[T_RHIST].p;
      k_radix_split<<<(nsm + TB - 1) / TB, TB, 0, st>>>(B.keys, k0, v0, nsm);
      for (int pass = 0; pass < 8; pass++) {
        k_radix_hist<<<nb, RSORT_THREADS, 0, st>>>(k0, nsm, 4 * pass, hist, nb);
        k_radix_scan<<<1, 1024, 0, st>>>(hist, 16 * nb);
        k_radix_scatter<<<nb, RSORT_THREADS, 0, st>>>(k0, v0, k1, v1, nsm, 4 * pass, hist, nb);
        std::swap(k0, k1);
        std::swap(v0, v1);
      }
      k_radix_join<<<(nsm + TB - 1) / TB, TB, 0, st>>>(k0, v0, B.keys, nsm);
      ctx->launches += 26;
      CK(cudaGetLastError());
    } else {
      // bitonic sort of the unique 64-bit keys: steps with j < TILE run in shared memory
      if (n_pad <= 2 * TILE && !(sort_env && n_pad > TILE)) {
        k_bitonic_smem<<<1, n_pad / 2, sizeof(unsigned long long) * (size_t)n_pad, st>>>(B.keys, n_pad, 2, n_pad, 1);
        ctx->launches++;
      } else {
        k_bitonic_smem<<<n_pad / TILE, SORT_THREADS, sizeof(unsigned long long) * TILE, st>>>(B.keys, n_pad, 2, TILE, 1);
        ctx->launches++;
        for (int k = 2 * TILE; k <= n_pad; k <<= 1) {
          int j = k >> 1;
          for (; j >= TILE; j >>= 1) {
            k_bitonic<<<(n_pad + TB - 1) / TB, TB, 0, st>>>(B.keys, n_pad, j, k);
            ctx->launches++;
          }
          k_bitonic_smem<<<n_pad / TILE, SORT_THREADS, sizeof(unsigned long long) * TILE, st>>>(B.keys, n_pad, k, k, j);
          ctx->launches++;
        }
      }
    }
    if (nsm > 1) {
      k_karras<<<(nsm - 1 + TB - 1) / TB, TB, 0, st>>>(B);
      ctx->launches++;
    }
    const int rounds = quality > 0 ? 1 + 2 * std::min(quality, 4) : 1;
    for (int r = 0; r < rounds; r++) {
      if (r > 0) CK(cudaMemsetAsync(B.flag, 0, sizeof(int) * (size_t)nsm, st));
      k_fit<<<(nsm + TB - 1) / TB, TB, 0, st>>>(B, (quality > 0 && r < rounds - 1) ? 1 : 0);
      ctx->launches++;
    }
    if (nsm > 1) k_order<<<(nsm - 1 + TB - 1) / TB, TB, 0, st>>>(B);
    k_pack<<<dim3((n_nodes + TB - 1) / TB, RT_N_ORDERINGS), TB, 0, st>>>(B);
    ctx->launches += 2;
    CK(cudaGetLastError());
    // nodes that survive leaf collapsing = kept size of the root (build node 0); the root's box (record 0 of every
    // ordering) = the frame of the render kernel's 16-bit quantised nodes / the object box of a group
    CK(cudaMemcpyAsync(h_words + 40, B.size, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h_words + 48, (const char *)ctx->d_nodes.p + 32 * (size_t)U.link_base, sizeof U.root_box,
                       cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st)); // host round trip 2 of 2
    U.kept_nodes = h_words[40];
    memcpy(U.root_box, h_words + 48, sizeof U.root_box);
  }
#undef RSV
  return RT_OK;
}

// World box of every instance from its group's root box (object space): the eight corners through M in double,
// rounded outwards.
static int upload_instance_boxes(rt_ctx *ctx) {
  const int ni = (int)ctx->instances.size();
  int rc;
  if ((rc = dev_reserve(ctx, ctx->d_inst_lo, sizeof(float4) * (size_t)std::max(ni, 1)))) return rc;
  if ((rc = dev_reserve(ctx, ctx->d_inst_hi, sizeof(float4) * (size_t)std::max(ni, 1)))) return rc;
  ctx->h_inst_box.assign(2 * (size_t)ni, make_float4(0, 0, 0, 0));
  for (int i = 0; i < ni; i++) {
    const rt_instance &I = ctx->instances[i];
    const float *rb = ctx->group_units[I.group].root_box;
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int c = 0; c < 8; c++) {
      const double p[3] = {(c & 1) ? rb[4] : rb[0], (c & 2) ? rb[5] : rb[1], (c & 4) ? rb[6] : rb[2]};
      for (int a = 0; a < 3; a++) {
        const double w = (double)I.m[4 * a] * p[0] + (double)I.m[4 * a + 1] * p[1] + (double)I.m[4 * a + 2] * p[2] + (double)I.m[4 * a + 3];
        lo[a] = std::min(lo[a], w); hi[a] = std::max(hi[a], w);
      }
    }
    float l[3], h[3];
    for (int a = 0; a < 3; a++) {
      // the object-space ray is rounded when it is transformed: pad by a few ulps of the box's reach
      const double e = 4e-7 * std::max(std::fabs(lo[a]), std::fabs(hi[a])) + 1e-9;
      l[a] = nextafterf((float)(lo[a] - e), -INFINITY);
      h[a] = nextafterf((float)(hi[a] + e), INFINITY);
    }
    ctx->h_inst_box[i] = make_float4(l[0], l[1], l[2], 0.f);
    ctx->h_inst_box[ni + i] = make_float4(h[0], h[1], h[2], 0.f);
  }
  if (ni) {
    CK(cudaMemcpyAsync(ctx->d_inst_lo.p, ctx->h_inst_box.data(), sizeof(float4) * (size_t)ni, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_inst_hi.p, ctx->h_inst_box.data() + ni, sizeof(float4) * (size_t)ni, cudaMemcpyHostToDevice, ctx->stream));
  }
  return RT_OK;
}

// The top level of an instanced scene: one tree over (world primitives, instances); its slot is the head of the
// packed array, sized for the worst case, so the groups' trees behind it never move.
static int build_top_level(rt_ctx *ctx, int quality) {
  int rc;
  if ((rc = upload_instance_boxes(ctx))) return rc;
  BuildUnit &U = ctx->top_unit;
  U = BuildUnit();
  U.count[0] = ctx->S.n_spheres; U.count[1] = ctx->S.n_tris; U.count[2] = ctx->S.n_quads;
  U.n_inst = (int)ctx->instances.size();
  U.top_level = true;
  U.classify = true;
  U.link_base = 0; U.leaf_base = 0;
  U.stride_nodes = ctx->total_nodes;
  U.end_link = ctx->total_nodes << RT_NODE_SHIFT;
  if ((rc = build_unit(ctx, U, quality))) return rc;
  memcpy(ctx->root_box, U.root_box, sizeof ctx->root_box);
  // no top-level tree at all (nothing but always-tested primitives): nothing to traverse
  ctx->S.n_nodes = U.n_small > 0 ? ctx->total_nodes : 0;
  ctx->stats.n_nodes = U.kept_nodes;
  for (const BuildUnit &G : ctx->group_units) ctx->stats.n_nodes += G.kept_nodes;
  ctx->stats.n_big_prims = U.n_big;
  // unit table for the shared-memory staging loop: per tree {first record, kept nodes, end code, -} {qbase} {qinv},
  // ascending by first record (the top level, then the groups in layout order)
  const int nu = 1 + (int)ctx->group_units.size();
  ctx->h_units.assign(3 * (size_t)nu, make_float4(0, 0, 0, 0));
  for (int u = 0; u < nu; u++) {
    const BuildUnit &B = u == 0 ? U : ctx->group_units[u - 1];
    float qb[3], qs[3], qi[3];
    quant_frame(B.root_box, RT_Q_MAX, qb, qs, qi);
    ctx->h_units[3 * (size_t)u] = make_float4(RT_I2F(B.link_base), RT_I2F(B.kept_nodes), RT_I2F(u == 0 ? 0 : RT_POP_SHARED), 0.f);
    ctx->h_units[3 * (size_t)u + 1] = make_float4(qb[0], qb[1], qb[2], 0.f);
    ctx->h_units[3 * (size_t)u + 2] = make_float4(qi[0], qi[1], qi[2], 0.f);
  }
  if ((rc = dev_reserve(ctx, ctx->d_units, sizeof(float4) * 3 * (size_t)nu))) return rc;
  CK(cudaMemcpyAsync(ctx->d_units.p, ctx->h_units.data(), sizeof(float4) * 3 * (size_t)nu, cudaMemcpyHostToDevice, ctx->stream));
  return RT_OK;
}

int rt_accel_build(rt_ctx *ctx, int quality) {
  if (!ctx) return RT_ERR_INVALID;
  if (!ctx->have_scene) return fail(ctx, RT_ERR_STATE, "rt_accel_build before rt_scene_upload");
  CK(cudaSetDevice(ctx->cfg.device));
  cudaStream_t st = ctx->stream;
  const int ns = ctx->S.n_spheres, nt = ctx->S.n_tris, nq = ctx->S.n_quads;
  const int n = ns + nt + nq;
  ctx->big_ids.clear();
  ctx->S.n_nodes = 0; ctx->S.n_big = 0; ctx->S.n_bigq = 0; ctx->n_leaf_prims = 0;
  ctx->stats.n_nodes = 0; ctx->stats.n_big_prims = 0;
  ctx->max_leaf = 1;
  ctx->build_quality = quality;
  memset(ctx->root_box, 0, sizeof ctx->root_box);
  if (n == 0) { ctx->have_accel = true; ctx->stats.ms_build = 0; return RT_OK; }
  CK(cudaEventRecord(ctx->ev0, st));
  int rc;
  if ((rc = dev_reserve(ctx, ctx->d_sph_is_big, (size_t)std::max(ns, 1)))) return rc;
  CK(cudaMemsetAsync(ctx->d_sph_is_big.p, 0, (size_t)std::max(ns, 1), st));
  if (!ctx->two_level) {
    BuildUnit U;
    U.count[0] = ns; U.count[1] = nt; U.count[2] = nq;
    if ((rc = build_unit(ctx, U, quality))) return rc;
    memcpy(ctx->root_box, U.root_box, sizeof ctx->root_box);
    ctx->n_leaf_prims = U.n_small;
    ctx->S.n_nodes = U.kept_nodes;
    ctx->S.node_stride = U.n_small > 0 ? (2 * U.n_small - 1) * 32 : 0;
    ctx->stats.n_nodes = U.kept_nodes;
    ctx->stats.n_big_prims = U.n_big;
  } else {
    // layout of the packed arrays: [top level: worst case 2 (world + instances) - 1 nodes][group 0][group 1]...
    const int ng = (int)ctx->groups.size(), ni = (int)ctx->instances.size();
    int n_world = 0;
    for (int i = 0; i < n; i++) n_world += ctx->grouped[i] ? 0 : 1;
    const int top_leaves = n_world + ni;
    // (every slot ends with one spare record: the sentinel of the shared-memory copies, k_render staging loop)
    int node_at = std::max(2 * top_leaves, 2), leaf_at = std::max(top_leaves, 1);
    ctx->group_units.assign(ng, BuildUnit());
    for (int g = 0; g < ng; g++) {
      const rt_group &G = ctx->groups[g];
      BuildUnit &U = ctx->group_units[g];
      U.first[0] = G.first_sphere; U.count[0] = G.n_spheres;
      U.first[1] = G.first_triangle; U.count[1] = G.n_triangles;
      U.first[2] = G.first_quad; U.count[2] = G.n_quads;
      U.classify = false;
      U.link_base = node_at; U.leaf_base = leaf_at;
      const int m = G.n_spheres + G.n_triangles + G.n_quads;
      node_at += 2 * m;
      leaf_at += m;
    }
    ctx->total_nodes = node_at;
    if ((long long)node_at * 32 * RT_N_ORDERINGS >= (1ll << 31))
      return fail(ctx, RT_ERR_INVALID, "two-level scene too large for 32-bit node offsets");
    if ((rc = dev_reserve(ctx, ctx->d_nodes, RT_N_ORDERINGS * sizeof(float4) * 2 * (size_t)node_at))) return rc;
    if ((rc = dev_reserve(ctx, ctx->d_leaf_prims, sizeof(int32_t) * (size_t)leaf_at))) return rc;
    ctx->n_leaf_prims = leaf_at;
    std::vector<int> root_off(ng, 0);
    for (int g = 0; g < ng; g++) {
      BuildUnit &U = ctx->group_units[g];
      U.stride_nodes = node_at;
      U.end_link = RT_POP_LINK(node_at);
      if ((rc = build_unit(ctx, U, quality))) return rc;
      root_off[g] = U.link_base << RT_NODE_SHIFT;
    }
    ctx->group_root_off = root_off;
    // instance records with the groups' root offsets and quantisation frames
    ctx->group_frames.assign(6 * (size_t)ng, 0.f);
    for (int g = 0; g < ng; g++) {
      float qinv[3];
      quant_frame(ctx->group_units[g].root_box, RT_Q_MAX, &ctx->group_frames[6 * (size_t)g], &ctx->group_frames[6 * (size_t)g + 3], qinv);
    }
    std::vector<float4> rec;
    if ((rc = make_instance_records(ctx->instances.data(), ni, ng, root_off.data(), rec, ctx->err, ctx->group_frames.data()))) return rc;
    ctx->h_inst_rec = rec;
    if (ni) CK(cudaMemcpyAsync((void *)ctx->S.inst, ctx->h_inst_rec.data(), sizeof(float4) * RT_INST_STRIDE * (size_t)ni, cudaMemcpyHostToDevice, st));
    ctx->S.node_stride = node_at * 32;
    if ((rc = build_top_level(ctx, quality))) return rc;
  }
  CK(cudaEventRecord(ctx->ev1, st));
  CK(cudaStreamSynchronize(st));
  CK(cudaEventElapsedTime(&ctx->stats.ms_build, ctx->ev0, ctx->ev1));
  ctx->S.nodes = (const float4 *)ctx->d_nodes.p;
  ctx->S.leaf_prims = (const int32_t *)ctx->d_leaf_prims.p;
  ctx->have_accel = true;
  return RT_OK;
}

int rt_instances_update(rt_ctx *ctx, const rt_instance *instances, int n) {
  if (!ctx || (n && !instances)) return RT_ERR_INVALID;
  if (!ctx->have_accel || !ctx->two_level) return fail(ctx, RT_ERR_STATE, "rt_instances_update needs a built two-level scene");
  if (n != (int)ctx->instances.size()) return fail(ctx, RT_ERR_INVALID, "rt_instances_update: %d instances, the scene has %d", n, (int)ctx->instances.size());
  for (int i = 0; i < n; i++)
    if (instances[i].group != ctx->instances[i].group) return fail(ctx, RT_ERR_INVALID, "rt_instances_update: instance %d changes its group", i);
  CK(cudaSetDevice(ctx->cfg.device));
  cudaStream_t st = ctx->stream;
  std::vector<float4> rec;
  int rc;
  if ((rc = make_instance_records(instances, n, (int)ctx->groups.size(), ctx->group_root_off.data(), rec, ctx->err,
                                  ctx->group_frames.data())))
    return rc;
  CK(cudaEventRecord(ctx->ev0, st));
  ctx->instances.assign(instances, instances + n);
  ctx->h_inst_rec = rec;
  if (n) CK(cudaMemcpyAsync((void *)ctx->S.inst, ctx->h_inst_rec.data(), sizeof(float4) * RT_INST_STRIDE * (size_t)n, cudaMemcpyHostToDevice, st));
  ctx->S.n_big = 0; ctx->S.n_bigq = 0;
  if ((rc = build_top_level(ctx, ctx->build_quality))) return rc;
  CK(cudaEventRecord(ctx->ev1, st));
  CK(cudaStreamSynchronize(st));
  CK(cudaEventElapsedTime(&ctx->stats.ms_build, ctx->ev0, ctx->ev1));
  return RT_OK;
}

int rt_accel_download(rt_ctx *ctx, rt_bvh_node *nodes, int cap_nodes, int *n_nodes, int32_t *leaf_prims, int cap_leaf,
                      int *n_leaf, int32_t *big_prims, int cap_big, int *n_big) {
  if (!ctx) return RT_ERR_INVALID;
  if (!ctx->have_accel) return fail(ctx, RT_ERR_STATE, "rt_accel_download before rt_accel_build");
  CK(cudaSetDevice(ctx->cfg.device));
  if (n_nodes) *n_nodes = ctx->S.n_nodes;
  if (n_leaf) *n_leaf = ctx->n_leaf_prims;
  if (n_big) *n_big = (int)ctx->big_ids.size();
  if (nodes) {
    if (cap_nodes < ctx->S.n_nodes) return fail(ctx, RT_ERR_INVALID, "node buffer too small");
    if (ctx->S.n_nodes)
      CK(cudaMemcpy(nodes, ctx->d_nodes.p, sizeof(rt_bvh_node) * (size_t)ctx->S.n_nodes, cudaMemcpyDeviceToHost));
    for (int i = 0; i < ctx->S.n_nodes; i++) { // device links are byte offsets; the ABI speaks indices
      nodes[i].escape >>= RT_NODE_SHIFT;
      if (nodes[i].payload >= 0) nodes[i].payload >>= RT_NODE_SHIFT;
    }
  }
  if (leaf_prims) {
    if (cap_leaf < ctx->n_leaf_prims) return fail(ctx, RT_ERR_INVALID, "leaf-primitive buffer too small");
    if (ctx->n_leaf_prims)
      CK(cudaMemcpy(leaf_prims, ctx->d_leaf_prims.p, sizeof(int32_t) * (size_t)ctx->n_leaf_prims,
                    cudaMemcpyDeviceToHost));
  }
  if (big_prims) {
    if (cap_big < (int)ctx->big_ids.size()) return fail(ctx, RT_ERR_INVALID, "big-primitive buffer too small");
    std::copy(ctx->big_ids.begin(), ctx->big_ids.end(), big_prims);
  }
  return RT_OK;
}

static int trace_through_render_kernel(rt_ctx *ctx, int n, float t_min, float t_max, cudaStream_t st);

static int trace_closest_impl(rt_ctx *ctx, const float *rays, int n, float t_min, float t_max, int use_accel, int32_t *prim_id,
                              int32_t *inst_id, float *t) {
  if (!ctx || n < 0 || (n && (!rays || !prim_id || !t))) return RT_ERR_INVALID;
  if (!ctx->have_scene) return fail(ctx, RT_ERR_STATE, "rt_trace_closest before rt_scene_upload");
  if (use_accel && !ctx->have_accel) return fail(ctx, RT_ERR_STATE, "use_accel=1 before rt_accel_build");
  if (n == 0) return RT_OK;
  CK(cudaSetDevice(ctx->cfg.device));
  int rc;
  if ((rc = dev_reserve(ctx, ctx->d_rays, sizeof(float) * 8 * (size_t)n))) return rc;
  if ((rc = dev_reserve(ctx, ctx->d_ids, sizeof(int32_t) * (size_t)n))) return rc;
  if ((rc = dev_reserve(ctx, ctx->d_ts, sizeof(float) * (size_t)n))) return rc;
  if ((rc = dev_reserve(ctx, ctx->d_inst_ids, sizeof(int32_t) * (size_t)n))) return rc;
  cudaStream_t st = ctx->stream;
  CK(cudaMemcpyAsync(ctx->d_rays.p, rays, sizeof(float) * 8 * (size_t)n, cudaMemcpyHostToDevice, st));
  DevScene S = ctx->S;
  if (!use_accel) { S.n_nodes = 0; S.n_big = 0; S.n_bigq = 0; }
  const int TB = 128, g = (n + TB - 1) / TB;
  const uint8_t *isbig = (const uint8_t *)ctx->d_sph_is_big.p;
  const float4 *dr = (const float4 *)ctx->d_rays.p;
  int32_t *di = (int32_t *)ctx->d_ids.p, *dn = (int32_t *)ctx->d_inst_ids.p;
  float *dt = (float *)ctx->d_ts.p;
  if (!ctx->two_level) CK(cudaMemsetAsync(dn, 0xff, sizeof(int32_t) * (size_t)n, st)); // every hit is a world-level primitive
  if (use_accel == 2) { // through k_render's own scheduler and traversal (TRACE instantiation)
    if ((rc = trace_through_render_kernel(ctx, n, t_min, t_max, st))) return rc;
  } else if (ctx->two_level) {
    k_trace_closest<2, true, true><<<g, TB, 0, st>>>(S, isbig, dr, n, t_min, t_max, use_accel, di, dt, (const uint8_t *)ctx->d_exclude.p, dn);
  } else
  switch (ctx->cfg.profile) {
  case 0: k_trace_closest<0, false><<<g, TB, 0, st>>>(S, isbig, dr, n, t_min, t_max, use_accel, di, dt); break;
  case 1: k_trace_closest<1, false><<<g, TB, 0, st>>>(S, isbig, dr, n, t_min, t_max, use_accel, di, dt); break;
  default: k_trace_closest<2, true><<<g, TB, 0, st>>>(S, isbig, dr, n, t_min, t_max, use_accel, di, dt); break;
  }
  ctx->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(prim_id, di, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(t, dt, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, st));
  if (inst_id) CK(cudaMemcpyAsync(inst_id, dn, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return RT_OK;
}

int rt_trace_closest(rt_ctx *ctx, const float *rays, int n, float t_min, float t_max, int use_accel, int32_t *prim_id,
                     float *t) {
  return trace_closest_impl(ctx, rays, n, t_min, t_max, use_accel, prim_id, nullptr, t);
}

int rt_trace_closest_inst(rt_ctx *ctx, const float *rays, int n, float t_min, float t_max, int use_accel, int32_t *prim_id,
                          int32_t *inst_id, float *t) {
  return trace_closest_impl(ctx, rays, n, t_min, t_max, use_accel, prim_id, inst_id, t);
}

// ------------------------------------------------------------------ render
typedef void (*render_kernel_t)(const RenderParams);
static render_kernel_t pick_render_kernel(int profile, int smem, bool count, bool ext, bool inst, bool sph) {
  if (sph) { // profile 2, spheres only (config 4)
#define PICKS(C) \
  return smem == 2 ? k_render<2, true, 2, C, false, false, false, true>                                 \
                   : (smem == 1 ? k_render<2, true, 1, C, false, false, false, true>                    \
                                : (smem == 3 ? k_render<2, true, 3, C, false, false, false, true> : k_render<2, true, 0, C, false, false, false, true>))
    if (count) { PICKS(true); }
    PICKS(false);
#undef PICKS
  }
  if (inst) { // two-level scenes: the general kernel's INST instantiations (the counting variant: global plan only)
    if (count) return ext ? k_render<2, true, 0, true, true, false, true> : k_render<2, true, 0, true, false, false, true>;
#define PICKI(E) \
  return smem == 2 ? k_render<2, true, 2, false, E, false, true>                                        \
                   : (smem == 1 ? k_render<2, true, 1, false, E, false, true>                           \
                                : (smem == 3 ? k_render<2, true, 3, false, E, false, true> : k_render<2, true, 0, false, E, false, true>))
    if (ext) { PICKI(true); }
    PICKI(false);
#undef PICKI
  }
#define PICK3(P, G, C, E) \
  (smem == 2 ? k_render<P, G, 2, C, E>                                                                  \
             : (smem == 1 ? k_render<P, G, 1, C, E> : (smem == 3 ? k_render<P, G, 3, C, E> : k_render<P, G, 0, C, E>)))
#define PICK(P, G, E) return count ? PICK3(P, G, true, E) : PICK3(P, G, false, E)
  if (profile == 0) { PICK(0, false, false); }
  if (profile == 1) { PICK(1, false, false); }
  if (ext) { PICK(2, true, true); }
  PICK(2, true, false);
#undef PICK
#undef PICK3
}

// Scheduler thresholds, the bytes of every scene array and the shared-memory residency plan of
// k_render (shared by the render launches and the TRACE parity hook, which must run the very same
// plan). smem: 2 = scene + eight octant orderings of the nodes, 1 = scene + one ordering, 3 = nodes only, 0 = global.
static void plan_scene_residency(rt_ctx *ctx, RenderParams &P, int block, int &smem, size_t &smem_bytes) {
  {
    const char *e = getenv("B200RT_BATCH"), *f = getenv("B200RT_FRAC8"); // tuning knobs
    // (re-swept after the FHFMA step made a BVH step cheaper against the other phases: a burst now runs until fewer
    // than 4/8 - before: 5/8 - of its entry lanes search; the sphere kernels of profiles 0 / 1 also shade a little
    // later. Config 2 per 250 spp: 33.17 ms at 26 / 4 against 33.45 ms at 24 / 5; config 3 per 150 spp 166.4 -> 164.3 ms,
    // config 4 per 200 spp 35.3 -> 34.9 ms, final scene per 200 spp 96.8 -> 95.2 ms at 24 / 4. The optimum is flat.)
    P.batch = e ? atoi(e) : (ctx->general ? 24 : 26);
    P.frac8 = f ? atoi(f) : 4;
    const char *lm = getenv("B200RT_LEAFMIN");
    P.leaf_min = lm ? atoi(lm) : 1; // measured: 1..5 within 0.5 %, larger values slower
    P.batch = std::max(1, std::min(P.batch, 32));
    const char *bl = getenv("B200RT_BATCH_LONG");
    P.batch_long = bl ? std::max(1, std::min(atoi(bl), 32)) : std::min(P.batch, 18);
    P.frac8 = std::max(0, std::min(P.frac8, 8));
  }
  const DevScene &S = ctx->S;
  // shared-memory node copies: 16-byte quantised records + one sentinel per copy (k_render staging loop)
  // (two-level scenes: S.n_nodes counts the slots of every tree, each of which ends with its own sentinel record)
  P.b_nodes = S.n_nodes ? 16 * (S.n_nodes + (ctx->two_level ? 0 : 1)) : 0;
  quant_frame(ctx->root_box, RT_Q_MAX, P.qbase, P.qscale, P.qinv);
  P.units = (const float4 *)ctx->d_units.p;
  P.n_units = ctx->two_level ? 1 + (int)ctx->group_units.size() : 0;
  P.b_sph = (int)pad16(sizeof(float4) * (size_t)S.n_spheres);
  P.b_sph_mat = (int)pad16(sizeof(int32_t) * (size_t)S.n_spheres);
  P.b_sph_k = (int)pad16(sizeof(float) * (size_t)S.n_spheres);
  P.b_mats = (int)pad16(sizeof(float4) * 2 * (size_t)S.n_mats);
  P.b_big = (int)pad16(sizeof(int32_t) * (size_t)S.n_big);
  P.b_bigq = ctx->general ? (int)pad16(sizeof(float4) * 2 * (size_t)S.n_bigq) : 0;
  P.b_leaf_prims = (int)pad16(sizeof(int32_t) * (size_t)ctx->n_leaf_prims);
  size_t scene_bytes = (size_t)P.b_nodes + P.b_sph + P.b_sph_k + P.b_sph_mat + P.b_mats + P.b_big + P.b_leaf_prims;
  if (ctx->general) {
    P.b_sph_mv = S.any_moving ? (int)pad16(sizeof(float4) * (size_t)S.n_spheres) : 0;
    P.b_sph_t0 = S.any_moving ? (int)pad16(sizeof(float) * (size_t)S.n_spheres) : 0;
    P.b_tri = (int)pad16(sizeof(float4) * 4 * (size_t)S.n_tris);
    P.b_tri_n = (int)pad16(sizeof(float4) * (size_t)S.n_tris);
    P.b_quad = (int)pad16(sizeof(float4) * 2 * (size_t)S.n_quads);
    P.b_tri_mat = (int)pad16(sizeof(int32_t) * (size_t)S.n_tris);
    P.b_quad_mat = (int)pad16(sizeof(int32_t) * (size_t)S.n_quads);
    P.b_box = (int)pad16(sizeof(float4) * 2 * (size_t)S.n_boxes);
    scene_bytes += (size_t)P.b_sph_mv + P.b_sph_t0 + P.b_tri + P.b_tri_n + P.b_quad + P.b_tri_mat + P.b_quad_mat + P.b_bigq + P.b_box;
  }
  const size_t acc_bytes = (size_t)(block / 32) * RT_ACC_WORDS * sizeof(unsigned); // two 32-pixel tiles per warp, 64-bit fixed-point sums
  // shared-memory plan: 2 = scene + eight octant orderings of the (quantised) nodes, 1 = scene, 0 = global
  smem = 0;
  if (scene_bytes + acc_bytes + 1024 > (size_t)ctx->max_smem_optin && ctx->general && S.n_boxes > 0 &&
      scene_bytes - P.b_quad - P.b_quad_mat + acc_bytes + 1024 <= (size_t)ctx->max_smem_optin) {
    // everything but the rect arrays fits, and the rects are mostly sides of boxes, whose records are resident:
    // leave the rects in global memory (shading reads axis / material / uv extents from there)
    scene_bytes -= (size_t)P.b_quad + P.b_quad_mat;
    P.b_quad = 0; P.b_quad_mat = 0;
  }
  if (scene_bytes + acc_bytes + 1024 <= (size_t)ctx->max_smem_optin) smem = 1;
  if (scene_bytes + (RT_N_ORDERINGS - 1) * (size_t)P.b_nodes + acc_bytes + 1024 <= (size_t)ctx->max_smem_optin) smem = 2;
  // nodes only: one node copy resident, primitives through L1/L2
  if (smem == 0 && (size_t)P.b_nodes + acc_bytes + 1024 <= (size_t)ctx->max_smem_optin && S.n_nodes > 0) smem = 3;
  // two-level scenes: every tree is quantised in its own frame (unit table); the staging loop looks the tree of a
  // record up by binary search, which is only worth it for a moderate number of trees
  if (ctx->two_level && (P.n_units > 4096 || (ctx->sp.flags & RT_FLAG_COUNTERS))) smem = 0;
  if (const char *e = getenv("B200RT_SMEM")) { // tuning knob: cap the plan (3 ranks between 0 and 1)
    const int cap = atoi(e);
    const int rank_of[4] = {0, 2, 3, 1}; // plan -> rank
    if (rank_of[smem] > rank_of[cap < 0 || cap > 3 ? 0 : cap]) smem = cap < 0 || cap > 3 ? 0 : cap;
  }
  // one-ordering plans with room to spare: 2 or 4 of the eight octant orderings (the per-lane plane selection of
  // those plans works for any ordering), chosen along the widest axes of the root box
  P.n_copies = 1; P.copy_oct[0] = P.copy_oct[1] = P.copy_oct[2] = P.copy_oct[3] = 0; P.oct_lut = 0u;
  // (only where the whole scene is resident: with primitives or textures still coming through L1 the extra shared
  // memory costs more cache than the better visiting order saves - measured on the rt_next_week final scene,
  // rects and textures in global memory: 113.8 ms with one copy, 117.6 ms with two)
  const bool all_resident = smem == 1 && !ctx->ext && !(ctx->general && S.n_quads > 0 && P.b_quad == 0);
  if ((all_resident || getenv("B200RT_COPIES")) && (smem == 1 || smem == 3) && S.n_nodes > 0 && S.node_stride != 0) {
    const size_t base = acc_bytes + 1024 + (smem == 3 ? (size_t)P.b_nodes : scene_bytes);
    int k = 1;
    if (base + 3 * (size_t)P.b_nodes <= (size_t)ctx->max_smem_optin) k = 4;
    else if (base + (size_t)P.b_nodes <= (size_t)ctx->max_smem_optin) k = 2;
    if (const char *e = getenv("B200RT_COPIES")) k = std::min(k, std::max(1, atoi(e))) >= 4 ? 4 : (std::min(k, std::max(1, atoi(e))) >= 2 ? 2 : 1); // measurement knob
    if (k > 1) {
      const float ext[3] = {ctx->root_box[4] - ctx->root_box[0], ctx->root_box[5] - ctx->root_box[1], ctx->root_box[6] - ctx->root_box[2]};
      int ax[3] = {0, 1, 2};
      std::sort(ax, ax + 3, [&](int a, int b) { return ext[a] > ext[b]; });
      const unsigned mask = k == 4 ? ((1u << ax[0]) | (1u << ax[1])) : (1u << ax[0]);
      int n = 0;
      int index_of[8];
      for (unsigned o = 0; o < 8; o++)
        if ((o & ~mask) == 0) { index_of[o] = n; P.copy_oct[n++] = (int)o; }
      for (unsigned o = 0; o < 8; o++) P.oct_lut |= (unsigned)index_of[o & mask] << (4 * o);
      P.n_copies = k;
    }
  }
  if (smem == 0) {
    // Global-memory node path: the octant-ordered copies pay while they stay cache
    // resident (measured: neutral up to 164 k nodes, one copy +14 % at 655 k nodes = 84 MB of copies).
    bool one_copy = RT_N_ORDERINGS * 32 * (size_t)S.n_nodes > ((size_t)48 << 20);
    if (const char *e = getenv("B200RT_ONECOPY")) one_copy = atoi(e) != 0; // measurement knob
    if (one_copy) P.S.node_stride = 0;
  }
  smem_bytes = acc_bytes + (smem == 3 ? (size_t)P.b_nodes : (smem ? scene_bytes : 0)) +
                            (smem == 2 ? (RT_N_ORDERINGS - 1) * (size_t)P.b_nodes : (size_t)(P.n_copies - 1) * P.b_nodes);
  // staging order in k_render: node copies first, then the sphere array
  P.off_sph = (smem == 2 ? RT_N_ORDERINGS : P.n_copies) * P.b_nodes;
  P.direct_leaf = ((!ctx->general || ctx->spheres_only) && (smem == 1 || smem == 2) && ctx->max_leaf == 1) ? 1 : 0;
}

static render_kernel_t pick_trace_kernel(int profile, int smem, bool inst, bool sph) {
  if (sph)
    return smem == 2 ? k_render<2, true, 2, false, false, true, false, true>
                     : (smem == 1 ? k_render<2, true, 1, false, false, true, false, true>
                                  : (smem == 3 ? k_render<2, true, 3, false, false, true, false, true> : k_render<2, true, 0, false, false, true, false, true>));
  if (inst)
    return smem == 2 ? k_render<2, true, 2, false, false, true, true>
                     : (smem == 1 ? k_render<2, true, 1, false, false, true, true>
                                  : (smem == 3 ? k_render<2, true, 3, false, false, true, true> : k_render<2, true, 0, false, false, true, true>));
#define PICKT(P, G) \
  return smem == 2 ? k_render<P, G, 2, false, false, true>                                              \
                   : (smem == 1 ? k_render<P, G, 1, false, false, true>                                 \
                                : (smem == 3 ? k_render<P, G, 3, false, false, true> : k_render<P, G, 0, false, false, true>))
  if (profile == 0) { PICKT(0, false); }
  if (profile == 1) { PICKT(1, false); }
  PICKT(2, true);
#undef PICKT
}

// rt_trace_closest(use_accel = 2): the rays in d_rays through the TRACE instantiation of k_render
// with the same residency plan, scheduler thresholds, grid and block as a render launch.
static int trace_through_render_kernel(rt_ctx *ctx, int n, float t_min, float t_max, cudaStream_t st) {
  RenderParams P;
  memset(&P, 0, sizeof P);
  P.S = ctx->S; P.cam = ctx->cam; P.sp = ctx->sp;
  P.sp.t_min = t_min;
  P.W = 8; P.H = 4; P.y0 = 0; P.y1 = 4; P.tiles_x = 1;
  P.trace_rays = (const float4 *)ctx->d_rays.p;
  P.trace_id = (int32_t *)ctx->d_ids.p;
  P.trace_t = (float *)ctx->d_ts.p;
  P.trace_inst = ctx->two_level ? (int32_t *)ctx->d_inst_ids.p : nullptr;
  P.trace_tmax = t_max;
  P.n_rays = n;
  P.trace_item = 256; // 8 "samples" of a 32-lane pool, like a render work item
  if (const char *e = getenv("B200RT_TRACE_ITEM")) P.trace_item = std::max(1, atoi(e));
  P.n_work = (n + P.trace_item - 1) / P.trace_item;
  P.n_tiles = P.n_work; // item w: tile w, chunk 0
  P.n_chunks = 1; P.chunk_spp = 1; P.spp_count = 1;
  const int grid = ctx->sm_count, block = RT_BLOCK_OF(ctx->general);
  int rc;
  if ((rc = dev_reserve(ctx, ctx->d_counter, 16))) return rc;
  if (!ctx->d_stats.p) {
    if ((rc = dev_reserve(ctx, ctx->d_stats, 64))) return rc;
    CK(cudaMemsetAsync(ctx->d_stats.p, 0, 64, st));
  }
  P.work_counter = (int *)ctx->d_counter.p;
  P.stats = (unsigned long long *)ctx->d_stats.p;
  int smem = 0;
  size_t smem_bytes = 0;
  plan_scene_residency(ctx, P, block, smem, smem_bytes);
  render_kernel_t kern = pick_trace_kernel(ctx->cfg.profile, smem, ctx->two_level, ctx->spheres_only);
  CK(cudaFuncSetAttribute((const void *)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  CK(cudaMemsetAsync(P.work_counter, 0, 4, st));
  kern<<<grid, block, smem_bytes, st>>>(P);
  CK(cudaGetLastError());
  ctx->stats.smem_bytes = (int)smem_bytes;
  ctx->stats.smem_plan = smem;
  return RT_OK;
}

static int render_into(rt_ctx *ctx, int W, int H, int spp_begin, int spp_count, float *d_accum, cudaStream_t st,
                       bool timed, int y0 = 0, int y1 = -1) {
  if (y1 < 0) y1 = H;
  if (W < 2 || H < 2 || spp_count < 0 || spp_begin < 0) return fail(ctx, RT_ERR_INVALID, "bad frame parameters");
  if (y0 < 0 || y1 > H || y0 > y1) return fail(ctx, RT_ERR_INVALID, "bad row range");
  if (y0 == y1) return RT_OK;
  if (!ctx->have_scene) return fail(ctx, RT_ERR_STATE, "rt_render before rt_scene_upload");
  if (!ctx->have_accel) {
    int rc = rt_accel_build(ctx, 1);
    if (rc) return rc;
  }
  if (spp_count == 0) return RT_OK;
  const size_t n_pix = (size_t)W * H;
  RenderParams P;
  memset(&P, 0, sizeof P);
  P.S = ctx->S; P.cam = ctx->cam; P.sp = ctx->sp;
  P.W = W; P.H = H;
  P.y0 = y0; P.y1 = y1;
  P.tiles_x = (W + RT_TILE_W - 1) / RT_TILE_W;
  P.n_tiles = P.tiles_x * ((y1 - y0 + RT_TILE_H - 1) / RT_TILE_H);
  const int grid = ctx->sm_count;
  const int block = RT_BLOCK_OF(ctx->general);
  const int n_warps = grid * (block / 32);
  // chunking: a work item is (tile, chunk of samples). Measured on B200: the
  // end-of-frame imbalance (few items per warp) costs far more than the per-item drain
  // tail, so aim for >= 64 items per resident warp while keeping >= 8 samples per item.
  int n_chunks = (int)((64LL * n_warps + P.n_tiles - 1) / P.n_tiles);
  n_chunks = std::min(n_chunks, std::max(spp_count >= 4 ? 2 : 1, spp_count / 12));
  if (const char *e = getenv("B200RT_CHUNKS")) n_chunks = atoi(e); // tuning knob
  n_chunks = std::max(1, std::min(n_chunks, std::min(spp_count, RT_MAX_CHUNKS)));
  {
    // Chunk sizes fall geometrically (ratio q per level, at least 1 sample): the work list is chunk-major, so the
    // last items every warp fetches are the small ones and the end-of-frame imbalance shrinks with them, while
    // most samples still travel in large items (an item boundary costs a partial drain of the warp).
    // Measured (B200, config 2): uniform chunks leave ~0.6 ms of tail per launch, which is what an 8-GPU sample
    // split (62 spp per GPU) loses against 1/8 of the single-GPU frame.
    double q = 0.65;
    if (const char *e = getenv("B200RT_CHUNK_RATIO")) q = std::min(1.0, std::max(0.05, atof(e))); // tuning knob; 1 = uniform
    std::vector<double> w((size_t)n_chunks);
    double sum = 0;
    for (int c = 0; c < n_chunks; c++) { w[c] = std::pow(q, c); sum += w[c]; }
    int given = 0;
    double acc = 0;
    P.chunk_begin[0] = 0;
    for (int c = 0; c < n_chunks; c++) {
      acc += w[c];
      int end = (int)std::llround(acc / sum * spp_count);
      end = std::max(end, given + 1);                          // every chunk holds at least one sample
      end = std::min(end, spp_count - (n_chunks - 1 - c));     // and leaves one for each later chunk
      if (c == n_chunks - 1) end = spp_count;
      P.chunk_begin[c + 1] = end;
      given = end;
    }
  }
  P.chunk_spp = P.chunk_begin[1];
  P.n_chunks = n_chunks;
  P.spp_begin = spp_begin; P.spp_count = spp_count;
  P.n_work = P.n_tiles * P.n_chunks;
  int rc;
  if ((rc = dev_reserve(ctx, ctx->d_partial, sizeof(float4) * n_pix * (size_t)P.n_chunks))) return rc;
  if ((rc = dev_reserve(ctx, ctx->d_counter, 16))) return rc;
  if (!ctx->d_stats.p) {
    if ((rc = dev_reserve(ctx, ctx->d_stats, 64))) return rc;
    CK(cudaMemsetAsync(ctx->d_stats.p, 0, 64, st));
  }
  P.partial = (float4 *)ctx->d_partial.p;
  P.work_counter = (int *)ctx->d_counter.p;
  P.stats = (unsigned long long *)ctx->d_stats.p;
  P.seed_lo = (uint32_t)(ctx->cfg.seed & 0xffffffffu);
  P.seed_hi = (uint32_t)(ctx->cfg.seed >> 32);
  philox_round_keys(P.seed_lo, P.seed_hi, P.philox_rk);
  int smem = 0;
  size_t smem_bytes = 0;
  plan_scene_residency(ctx, P, block, smem, smem_bytes);
  const bool count = (ctx->sp.flags & RT_FLAG_COUNTERS) != 0;
  render_kernel_t kern = pick_render_kernel(ctx->cfg.profile, smem, count, ctx->ext, ctx->two_level, ctx->spheres_only);
  CK(cudaFuncSetAttribute((const void *)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  cudaFuncAttributes fa;
  CK(cudaFuncGetAttributes(&fa, (const void *)kern));
  CK(cudaMemsetAsync(P.work_counter, 0, 4, st));
  if (timed) CK(cudaEventRecord(ctx->ev0, st));
  cudaEvent_t *evk = ctx->ev_k[ctx->ev_k_count % RT_EV_RING];
  CK(cudaEventRecord(evk[0], st));
  ctx->last_render_stream = st;
  kern<<<grid, block, smem_bytes, st>>>(P);
  CK(cudaGetLastError());
  CK(cudaEventRecord(evk[1], st));
  const int first_px = y0 * W, n_band = (y1 - y0) * W;
  k_combine<<<(unsigned)((n_band + 255) / 256), 256, 0, st>>>((float4 *)d_accum, P.partial, (int)n_pix, P.n_chunks, first_px,
                                                              n_band);
  CK(cudaGetLastError());
  ctx->launches += 2;
  CK(cudaEventRecord(evk[2], st));
  ctx->ev_k_count++;
  if (timed) CK(cudaEventRecord(ctx->ev1, st));
  ctx->stats.smem_bytes = (int)smem_bytes;
  ctx->stats.smem_plan = smem;
  ctx->stats.block_threads = block;
  ctx->stats.grid_blocks = grid;
  ctx->stats.regs_per_thread = fa.numRegs;
  return RT_OK;
}

static int ensure_frame(rt_ctx *ctx, int W, int H) {
  if (ctx->W == W && ctx->H == H && ctx->d_accum.p) return RT_OK;
  int rc = dev_reserve(ctx, ctx->d_accum, sizeof(float4) * (size_t)W * H);
  if (rc) return rc;
  CK(cudaMemsetAsync(ctx->d_accum.p, 0, sizeof(float4) * (size_t)W * H, ctx->stream));
  ctx->W = W; ctx->H = H;
  return RT_OK;
}

int rt_render(rt_ctx *ctx, int width, int height, int spp_begin, int spp_count) {
  if (!ctx) return RT_ERR_INVALID;
  CK(cudaSetDevice(ctx->cfg.device));
  int rc = ensure_frame(ctx, width, height);
  if (rc) return rc;
  rc = render_into(ctx, width, height, spp_begin, spp_count, (float *)ctx->d_accum.p, ctx->stream, true);
  if (rc) return rc;
  CK(cudaStreamSynchronize(ctx->stream));
  if (spp_count > 0) CK(cudaEventElapsedTime(&ctx->stats.ms_render, ctx->ev0, ctx->ev1));
  return RT_OK;
}

int rt_render_device(rt_ctx *ctx, int width, int height, int spp_begin, int spp_count, float *d_accum, void *stream) {
  if (!ctx || !d_accum) return RT_ERR_INVALID;
  CK(cudaSetDevice(ctx->cfg.device));
  return render_into(ctx, width, height, spp_begin, spp_count, d_accum, stream ? (cudaStream_t)stream : ctx->stream,
                     false);
}

int rt_render_rows_device(rt_ctx *ctx, int width, int height, int row_begin, int row_end, int spp_begin, int spp_count,
                          float *d_accum, void *stream) {
  if (!ctx || !d_accum) return RT_ERR_INVALID;
  CK(cudaSetDevice(ctx->cfg.device));
  return render_into(ctx, width, height, spp_begin, spp_count, d_accum, stream ? (cudaStream_t)stream : ctx->stream,
                     false, row_begin, row_end);
}

int rt_accum_clear(rt_ctx *ctx) {
  if (!ctx) return RT_ERR_INVALID;
  CK(cudaSetDevice(ctx->cfg.device));
  if (ctx->d_accum.p) CK(cudaMemsetAsync(ctx->d_accum.p, 0, sizeof(float4) * (size_t)ctx->W * ctx->H, ctx->stream));
  return RT_OK;
}

int rt_accum_download(rt_ctx *ctx, float *rgba, size_t n_floats) {
  if (!ctx || !rgba) return RT_ERR_INVALID;
  if (!ctx->d_accum.p) return fail(ctx, RT_ERR_STATE, "no frame");
  if (n_floats != (size_t)ctx->W * ctx->H * 4) return fail(ctx, RT_ERR_INVALID, "size mismatch");
  CK(cudaSetDevice(ctx->cfg.device));
  CK(cudaMemcpyAsync(rgba, ctx->d_accum.p, n_floats * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return RT_OK;
}

int rt_accum_upload(rt_ctx *ctx, int width, int height, const float *rgba, size_t n_floats) {
  if (!ctx || !rgba) return RT_ERR_INVALID;
  if (n_floats != (size_t)width * height * 4) return fail(ctx, RT_ERR_INVALID, "size mismatch");
  CK(cudaSetDevice(ctx->cfg.device));
  int rc = ensure_frame(ctx, width, height);
  if (rc) return rc;
  CK(cudaMemcpyAsync(ctx->d_accum.p, rgba, n_floats * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return RT_OK;
}

int rt_render_aov(rt_ctx *ctx, int width, int height, int spp, float *aov) {
  if (!ctx || !aov || width < 2 || height < 2 || spp < 1) return RT_ERR_INVALID;
  if (!ctx->have_scene) return fail(ctx, RT_ERR_STATE, "rt_render_aov before rt_scene_upload");
  CK(cudaSetDevice(ctx->cfg.device));
  if (!ctx->have_accel) {
    int rc = rt_accel_build(ctx, 1);
    if (rc) return rc;
  }
  const size_t n_pix = (size_t)width * height;
  int rc = dev_reserve(ctx, ctx->d_linear, sizeof(float) * 8 * n_pix);
  if (rc) return rc;
  cudaStream_t st = ctx->stream;
  const unsigned grid = (unsigned)((n_pix + 127) / 128);
  const uint32_t k0 = (uint32_t)(ctx->cfg.seed & 0xffffffffu), k1 = (uint32_t)(ctx->cfg.seed >> 32);
  float *out = (float *)ctx->d_linear.p;
  switch (ctx->cfg.profile) {
  case 0: k_aov<0, false><<<grid, 128, 0, st>>>(ctx->S, ctx->cam, ctx->sp, width, height, spp, k0, k1, out); break;
  case 1: k_aov<1, false><<<grid, 128, 0, st>>>(ctx->S, ctx->cam, ctx->sp, width, height, spp, k0, k1, out); break;
  default:
    if (ctx->two_level) k_aov<2, true, true><<<grid, 128, 0, st>>>(ctx->S, ctx->cam, ctx->sp, width, height, spp, k0, k1, out);
    else k_aov<2, true><<<grid, 128, 0, st>>>(ctx->S, ctx->cam, ctx->sp, width, height, spp, k0, k1, out);
    break;
  }
  ctx->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(aov, out, sizeof(float) * 8 * n_pix, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return RT_OK;
}

void *rt_accum_device_ptr(rt_ctx *ctx) { return ctx ? ctx->d_accum.p : nullptr; }

int rt_resolve_device(rt_ctx *ctx, int width, int height, const float *d_accum, float *linear_rgb, uint8_t *rgb8,
                      void *stream) {
  if (!ctx || !d_accum || width < 1 || height < 1) return RT_ERR_INVALID;
  CK(cudaSetDevice(ctx->cfg.device));
  cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
  const size_t n_pix = (size_t)width * height;
  int rc;
  if (linear_rgb && (rc = dev_reserve(ctx, ctx->d_linear, sizeof(float) * 3 * n_pix))) return rc;
  if (rgb8 && (rc = dev_reserve(ctx, ctx->d_rgb8, 3 * n_pix))) return rc;
  CK(cudaEventRecord(ctx->ev0, st));
  k_resolve<<<(unsigned)((n_pix + 255) / 256), 256, 0, st>>>((const float4 *)d_accum, width, height, ctx->cfg.profile,
                                                             linear_rgb ? (float *)ctx->d_linear.p : nullptr,
                                                             rgb8 ? (uint8_t *)ctx->d_rgb8.p : nullptr);
  ctx->launches++;
  CK(cudaGetLastError());
  // device -> host: straight into the caller's buffer when it is page-locked (rt_host_alloc, cudaHostRegister,
  // torch pinned tensors); otherwise through the context's pinned block + one host memcpy, which beats the
  // driver's pageable path (it stages through small internal buffers and blocks the stream meanwhile)
  auto is_pinned = [](const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
  };
  const size_t lin_bytes = linear_rgb ? sizeof(float) * 3 * n_pix : 0, rgb_bytes = rgb8 ? 3 * n_pix : 0;
  const bool lin_direct = linear_rgb && is_pinned(linear_rgb), rgb_direct = rgb8 && is_pinned(rgb8);
  const size_t stage_lin = lin_direct ? 0 : pad16(lin_bytes), stage_rgb = rgb_direct ? 0 : rgb_bytes;
  if (stage_lin + stage_rgb) {
    if ((rc = host_reserve(ctx, ctx->h_out, ctx->h_out_bytes, stage_lin + stage_rgb))) return rc;
  }
  unsigned char *ho = (unsigned char *)ctx->h_out;
  if (linear_rgb) CK(cudaMemcpyAsync(lin_direct ? (void *)linear_rgb : (void *)ho, ctx->d_linear.p, lin_bytes, cudaMemcpyDeviceToHost, st));
  if (rgb8) CK(cudaMemcpyAsync(rgb_direct ? (void *)rgb8 : (void *)(ho + stage_lin), ctx->d_rgb8.p, rgb_bytes, cudaMemcpyDeviceToHost, st));
  CK(cudaEventRecord(ctx->ev1, st));
  CK(cudaStreamSynchronize(st));
  if (linear_rgb && !lin_direct) memcpy(linear_rgb, ho, lin_bytes);
  if (rgb8 && !rgb_direct) memcpy(rgb8, ho + stage_lin, rgb_bytes);
  CK(cudaEventElapsedTime(&ctx->stats.ms_resolve, ctx->ev0, ctx->ev1));
  return RT_OK;
}

int rt_resolve(rt_ctx *ctx, float *linear_rgb, uint8_t *rgb8) {
  if (!ctx) return RT_ERR_INVALID;
  if (!ctx->d_accum.p) return fail(ctx, RT_ERR_STATE, "rt_resolve before rt_render");
  return rt_resolve_device(ctx, ctx->W, ctx->H, (const float *)ctx->d_accum.p, linear_rgb, rgb8, nullptr);
}

// FP32 FMA issue-rate microbenchmark: the roofline denominator of this compute-bound
// path (MEASURED_PEAKS.json only holds HBM and bf16-tensor peaks).
__global__ void k_fma_peak(float *out, int iters, float a, float b) {
  float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      x0 = __fmaf_rn(x0, a, b); x1 = __fmaf_rn(x1, a, b); x2 = __fmaf_rn(x2, a, b); x3 = __fmaf_rn(x3, a, b);
      x4 = __fmaf_rn(x4, a, b); x5 = __fmaf_rn(x5, a, b); x6 = __fmaf_rn(x6, a, b); x7 = __fmaf_rn(x7, a, b);
    }
  }
  float s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (s == 12345.678f) out[0] = s;
}

int rt_measure_fp32_peak(rt_ctx *ctx, float *tflops) {
  if (!ctx || !tflops) return RT_ERR_INVALID;
  CK(cudaSetDevice(ctx->cfg.device));
  int rc = dev_reserve(ctx, ctx->d_counter, 16);
  if (rc) return rc;
  const int iters = 4096, threads = 512, grid = ctx->sm_count * 4;
  float best = 0.f;
  for (int rep = 0; rep < 4; rep++) {
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    k_fma_peak<<<grid, threads, 0, ctx->stream>>>((float *)ctx->d_counter.p, iters, 0.999f, 0.001f);
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    double flops = 2.0 * 64.0 * (double)iters * (double)threads * (double)grid;
    float tf = (float)(flops / (ms * 1e-3) / 1e12);
    if (rep > 0 && tf > best) best = tf;
  }
  ctx->launches += 4;
  *tflops = best;
  return RT_OK;
}

int rt_stats(rt_ctx *ctx, rt_stats_t *out) {
  if (!ctx || !out) return RT_ERR_INVALID;
  CK(cudaSetDevice(ctx->cfg.device));
  unsigned long long h[4] = {0, 0, 0, 0};
  if (ctx->d_stats.p) {
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->last_render_stream && ctx->last_render_stream != ctx->stream) CK(cudaStreamSynchronize(ctx->last_render_stream));
    CK(cudaMemcpy(h, ctx->d_stats.p, sizeof h, cudaMemcpyDeviceToHost));
  }
  ctx->stats.paths = h[0]; ctx->stats.segments = h[1]; ctx->stats.box_tests = h[2]; ctx->stats.prim_tests = h[3];
  ctx->stats.kernel_launches = ctx->launches;
  { // mean kernel durations of the (last <= RT_EV_RING) render launches since the reset, whichever stream they ran on
    const int n = std::min(ctx->ev_k_count, RT_EV_RING);
    double sr = 0, sc = 0;
    int got = 0;
    for (int i = 0; i < n; i++) {
      float a = 0, b = 0;
      if (cudaEventSynchronize(ctx->ev_k[i][2]) == cudaSuccess && cudaEventElapsedTime(&a, ctx->ev_k[i][0], ctx->ev_k[i][1]) == cudaSuccess &&
          cudaEventElapsedTime(&b, ctx->ev_k[i][1], ctx->ev_k[i][2]) == cudaSuccess) { sr += a; sc += b; got++; }
    }
    ctx->stats.ms_k_render = got ? (float)(sr / got) : 0.f;
    ctx->stats.ms_k_combine = got ? (float)(sc / got) : 0.f;
    cudaGetLastError();
  }
  *out = ctx->stats;
  return RT_OK;
}

void *rt_host_alloc(rt_ctx *ctx, size_t bytes) {
  if (!ctx || bytes == 0) return nullptr;
  if (cudaSetDevice(ctx->cfg.device) != cudaSuccess) return nullptr;
  void *p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}

void rt_host_free(rt_ctx *ctx, void *p) {
  (void)ctx;
  if (p) cudaFreeHost(p);
}

// ------------------------------------------------------------------ multi-GPU reduce
int rt_comm_unique_id(void *id128) {
  if (!id128) return RT_ERR_INVALID;
  NcclApi &a = nccl_api();
  if (!a.handle) return RT_ERR_STATE;
  static_assert(sizeof(ncclUniqueId) == RT_COMM_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId id;
  if (a.GetUniqueId(&id) != ncclSuccess) return RT_ERR_CUDA;
  memcpy(id128, &id, sizeof id);
  return RT_OK;
}

int rt_comm_init(rt_ctx *ctx, const void *id128, int rank, int nranks) {
  if (!ctx || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return RT_ERR_INVALID;
  NcclApi &a = nccl_api();
  if (!a.handle) return fail(ctx, RT_ERR_STATE, "%s", a.why.c_str());
  CK(cudaSetDevice(ctx->cfg.device));
  if (ctx->comm) { a.CommDestroy(ctx->comm); ctx->comm = nullptr; }
  ncclUniqueId id;
  memcpy(&id, id128, sizeof id);
  ncclResult_t r = a.CommInitRank(&ctx->comm, nranks, id, rank);
  if (r != ncclSuccess) { ctx->comm = nullptr; return fail(ctx, RT_ERR_CUDA, "ncclCommInitRank: %s", a.GetErrorString(r)); }
  ctx->comm_rank = rank; ctx->comm_size = nranks;
  return RT_OK;
}

int rt_comm_init_all(rt_ctx **ctxs, int n) {
  if (!ctxs || n < 1 || n > 64) return RT_ERR_INVALID;
  for (int i = 0; i < n; i++) if (!ctxs[i]) return RT_ERR_INVALID;
  rt_ctx *ctx = ctxs[0];
  NcclApi &a = nccl_api();
  if (!a.handle) return fail(ctx, RT_ERR_STATE, "%s", a.why.c_str());
  int devs[64];
  ncclComm_t comms[64];
  for (int i = 0; i < n; i++) {
    devs[i] = ctxs[i]->cfg.device;
    if (ctxs[i]->comm) { a.CommDestroy(ctxs[i]->comm); ctxs[i]->comm = nullptr; }
  }
  ncclResult_t r = a.CommInitAll(comms, n, devs);
  if (r != ncclSuccess) return fail(ctx, RT_ERR_CUDA, "ncclCommInitAll: %s", a.GetErrorString(r));
  for (int i = 0; i < n; i++) { ctxs[i]->comm = comms[i]; ctxs[i]->comm_rank = i; ctxs[i]->comm_size = n; }
  return RT_OK;
}

int rt_reduce(rt_ctx *ctx, int width, int height, float *d_accum, int root, unsigned flags, void *stream) {
  if (!ctx || width < 1 || height < 1) return RT_ERR_INVALID;
  if (!d_accum) d_accum = (float *)ctx->d_accum.p;
  if (!d_accum) return fail(ctx, RT_ERR_STATE, "rt_reduce: no frame");
  if (ctx->comm_size == 1 && !ctx->comm) return RT_OK; // a single rank: nothing to exchange
  if (!ctx->comm) return fail(ctx, RT_ERR_STATE, "rt_reduce before rt_comm_init");
  if (root < 0 || root >= ctx->comm_size) return RT_ERR_INVALID;
  NcclApi &a = nccl_api();
  CK(cudaSetDevice(ctx->cfg.device));
  cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
  const size_t n_pix = (size_t)width * height;
  ncclResult_t r;
  if (flags & RT_REDUCE_UNIFORM_COUNT) {
    int rc = dev_reserve(ctx, ctx->d_red, sizeof(float) * (3 * n_pix + 4));
    if (rc) return rc;
    float *red = (float *)ctx->d_red.p;
    const unsigned g = (unsigned)((n_pix + 255) / 256);
    k_reduce_pack<<<g, 256, 0, st>>>((const float4 *)d_accum, red, (int)n_pix);
    CK(cudaGetLastError());
    r = a.Reduce(red, red, 3 * n_pix + 1, ncclFloat32, ncclSum, root, ctx->comm, st);
    if (r == ncclSuccess && ctx->comm_rank == root) {
      k_reduce_unpack<<<g, 256, 0, st>>>((float4 *)d_accum, red, (int)n_pix);
      CK(cudaGetLastError());
      ctx->launches++;
    }
    ctx->launches++;
  } else {
    r = a.Reduce(d_accum, d_accum, 4 * n_pix, ncclFloat32, ncclSum, root, ctx->comm, st);
  }
  if (r != ncclSuccess) return fail(ctx, RT_ERR_CUDA, "ncclReduce: %s", a.GetErrorString(r));
  return RT_OK;
}

int rt_stats_reset(rt_ctx *ctx) {
  if (!ctx) return RT_ERR_INVALID;
  CK(cudaSetDevice(ctx->cfg.device));
  if (ctx->d_stats.p) CK(cudaMemsetAsync(ctx->d_stats.p, 0, 64, ctx->stream));
  ctx->launches = 0;
  ctx->ev_k_count = 0;
  return RT_OK;
}

} // extern "C"
