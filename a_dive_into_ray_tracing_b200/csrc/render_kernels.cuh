// render_kernels.cuh — the sm_100a kernels of the hot path.
//
//  k_render      persistent-thread path tracer. Replaces worker()
//                (rt_in_one_weekend/main.cpp:267-290) and render_init+render<<<grid,8x8>>>
//                (accelerated-rt-cuda/final.cu:62-96, rt_next_week/cuda/main.cu:113-149,
//                triangles/cuda/obj_render.cu:94-130): one thread per pixel looping over
//                all samples and up to 50 bounces there; here one resident CTA per SM,
//                the whole scene staged in shared memory, and every warp draining a
//                pool of (pixel, sample) items of one 8x4 tile: lanes whose path ended
//                take the next items by ballot/popc rank ("warp-aggregated work fetch"),
//                so all 32 lanes trace a segment in every iteration (path regeneration);
//                tiles are fetched from a global atomic counter.
//  k_combine     fixed-order sum of the per-chunk partial frames into the accumulation
//                buffer (deterministic: no float atomics on global memory).
//  k_resolve     /count, sqrt, clamp, 8-bit (color.h:14-28 / final.cu:91-95,223-232).
//  k_trace_closest  one ray per thread closest-hit query (parity hook).
#pragma once
#include <cuda_fp16.h>
#include "shade.cuh"

// Threads per CTA (one CTA per SM). The sphere-only kernels fit 64 registers -> 1024
// threads; the general kernel (triangles, rects, moving spheres, textures) is capped at 64 too
// (a few bytes of spills): measured faster than 768 threads x 72 registers.
#ifndef RT_BLOCK
#define RT_BLOCK 1024
#endif
#ifndef RT_BLOCK_GENERAL
#define RT_BLOCK_GENERAL 1024
#endif
// BVH steps between two warp votes in the search burst. Measured (B200, after the r1i instruction
// diet): sphere-only kernel 2/3/4 steps -> 78.2/76.8/77.5 ms (config 2); general kernel
// 173/171/167 ms (config 3, 150 spp).
// The extended kernel (media, noise / image textures) is bound by instruction fetch: a shorter unrolled burst is
// faster there (final scene 106.9 -> 101.9 ms per 200 spp with 2 steps).
#ifndef RT_SPV_EXT
#define RT_SPV_EXT 2
#endif
#ifndef RT_SPV_SPHERE
#define RT_SPV_SPHERE 3
#endif
#define RT_STEPS_PER_VOTE(GENERAL, EXT) ((EXT) ? RT_SPV_EXT : ((GENERAL) ? 4 : RT_SPV_SPHERE))
#define RT_BLOCK_OF(GENERAL) ((GENERAL) ? RT_BLOCK_GENERAL : RT_BLOCK)
#define RT_TILE_W 8
#define RT_TILE_H 4
#define RT_MAX_CHUNKS 64
#define RT_ACC_WORDS 384 // per-warp accumulator words: 2 tiles x 3 channels x {lo, hi} x 32 pixels
#define RT_Q_MAX 46335 // largest quantised plane: (q << 16) + 0x4B000000 must not carry out of 32 bits (k_render)
// two-level scenes, shared-memory plans: the link that ends an object's tree (shared addresses are multiples of 16
// above it, 0 = finished, negative = leaf payload)
#define RT_POP_SHARED 8
// Shared-memory plans: the planes of a node record as two fp16 numbers per word, taken STRAIGHT from the loaded word by
// sm_100's mixed-precision multiply-add FHFMA (PTX fma.rn.f32.f16: f32 = f16 * f16 + f32, either half of a 32-bit
// register selected in the instruction): t = v_h * S_h + C is ONE instruction per plane and the record needs no decode at
// all - the step of the octant-copy plans is LDS.128 + 6 FHFMA + 4 min/max + the link selection (19-20 instructions
// against 25-26 with 16-bit integer planes: 3 LEA.HI on the half-rate ALU pipe + 3 IMAD + 6 FFMA; ncu r2f had that pipe
// as the limiter, 13 of the step's 22 instructions). The one-copy plans swap the halves of a word per lane by the sign
// of the lane's direction (3 PRMT instead of 6) and run the same six FHFMA.
// The plane coordinate v = q / 23 >= 0 (q = the 16-bit coordinate over the tree's box, 0..46335; v < 2048) is rounded
// OUTWARDS to fp16 after the same two-step padding, and the ray keeps its slope per axis as an fp16 PAIR rounded apart,
// S_entry <= S <= S_exit (round-down | round-up): with v >= 0 the product errs to the conservative side for either sign of
// S; a slope beyond the fp16 range saturates to max-finite / infinity, i.e. a half-open slab, and 0 * inf = NaN is dropped
// by min / max. The boxes stay conservative - results are unchanged (parity through this traversal is bit-exact against
// brute force) - and the looser planes cost 0.7 % more box tests. Measured on config 2: 72.9 ms with integer planes,
// 68.4 ms with fp16 planes converted by HADD2.F32 + FFMA, 65.9 ms with FHFMA. RT_HALF_PLANES=0 keeps the integer form
// for A/B measurements.
#ifndef RT_HALF_PLANES
#define RT_HALF_PLANES 1
#endif
#define RT_HALF_KINV 23.0f // v = q / 23
#ifndef RT_N_ORDERINGS
#define RT_N_ORDERINGS 8 // node orderings = ray-direction octants (bvh_build.cuh body_order)
#endif

struct RenderParams {
  DevScene S;
  DevCamera cam;
  ShadeParams sp;
  int W, H, tiles_x, n_tiles;
  int y0, y1; // rows rendered by this launch: [y0, y1) (the whole frame: 0, H)
  int n_chunks, chunk_spp, spp_begin, spp_count, n_work;
  // samples [chunk_begin[c], chunk_begin[c + 1]) (relative to spp_begin) form chunk c: sizes DECREASE towards the end of
  // the work list, so that the frame's tail - warps waiting for the last items - is as short as the smallest chunk
  int chunk_begin[RT_MAX_CHUNKS + 1];
  float4 *partial;               // [n_chunks][H*W]
  int *work_counter;
  unsigned long long *stats;     // paths, segments, box tests, prim tests
  uint32_t seed_lo, seed_hi;
  uint32_t philox_rk[20]; // the ten Philox round keys of (seed_lo, seed_hi): philox_round_keys
  int batch;  // shade/regenerate when this many lanes are DONE or DEAD (warp-voted scheduler)
  int batch_long; // the same for warps whose paths run long (general kernels)
  int leaf_min; // run the pending primitive tests when this many lanes wait (or nobody searches)
  int frac8;  // a BVH burst ends when fewer than frac8/8 of its entry lanes are still searching
  // bytes of each array staged to shared memory (all multiples of 16)
  int b_nodes, b_sph, b_sph_k, b_sph_mv, b_sph_t0, b_tri, b_tri_n, b_quad, b_sph_mat, b_tri_mat, b_quad_mat, b_mats, b_big, b_leaf_prims, b_bigq, b_box;
  // sphere-only kernels with single-primitive leaves: the staged leaf payloads name the sphere
  // directly (~(sphere << 3)) and the sphere array sits at this byte offset of the shared copy
  int direct_leaf, off_sph;
  // 16-byte quantised nodes of the shared-memory copies: plane = qbase + q * qscale per axis (q = 0..65535
  // over the root box), see the staging loop of k_render
  float qbase[3], qscale[3], qinv[3];
  // two-level scenes: one entry per tree, ascending by first record - {as_int first record, kept nodes, end code}
  // {qbase.xyz} {qinv.xyz}: the staging loop quantises every tree in its own frame (the top level's is qbase / qinv)
  const float4 *units;
  int n_units;
  // one-ordering plans (SMEM 1 / 3) with room to spare hold 2 or 4 of the builder's eight octant orderings: copy c
  // is ordering copy_oct[c]; a ray of octant o walks copy (oct_lut >> 4 o) & 7 (the ordering that agrees with its
  // direction signs on the one or two axes along which the scene is widest)
  int n_copies, copy_oct[4];
  unsigned oct_lut;
  // TRACE instantiation (parity hook, rt_trace_closest use_accel = 2): the pool of a work item is a
  // run of caller-supplied rays instead of (pixel, sample) pairs; a finished traversal writes
  // (primitive id, t) instead of being shaded
  const float4 *trace_rays; // [n_rays][2]: {origin, time} {direction, -}
  int32_t *trace_id;
  int32_t *trace_inst; // two-level scenes: instance of the hit primitive (may be null)
  float *trace_t;
  float trace_tmax;
  int n_rays, trace_item; // rays per work item
};

__device__ __forceinline__ const void *stage_to_smem(unsigned char *smem, int &off, const void *src, int bytes) {
  unsigned char *dst = smem + off;
  const int n16 = bytes >> 4;
  const float4 *s = (const float4 *)src;
  float4 *d = (float4 *)dst;
  for (int i = threadIdx.x; i < n16; i += blockDim.x) d[i] = __ldg(s + i);
  off += bytes;
  return dst;
}

// SMEM: 3 = only the node array (one copy) staged in shared memory, primitives and materials read
// through L1/L2 (scenes whose nodes fit but whose primitives do not);
// 0 = scene read from global memory (L1/L2); 1 = whole scene staged in shared memory;
// 2 = as 1, with FOUR copies of the node array, one per sign combination of (d.x, d.z), whose
// box planes are pre-swapped so that the slab test needs no min/max on those axes.
// EXT: media + noise/image textures (rt_next_week scenes 3-8), profile 2 only.
// TRACE: the closest-hit parity hook THROUGH this kernel's own traversal: same staging, same lane
// state machine, same warp-voted search bursts / primitive-test rounds / regeneration by ballot
// rank, same ray_precompute_fast and quadrant copies - only the two ends differ: a new "path" is a
// caller ray (no camera, no Philox) and a finished traversal is written out instead of shaded.
// INST: two-level scenes (rt_group / rt_instance; general kernel, global-memory plan only). A top-level leaf may
// name an instance: the lane moves its ray into the object's space (rigid map: t is preserved, as
// translate::hit / rotate_y::hit do, rt_next_week/cuda/hittable.h:66-79,156-190), remembers where the top-level walk
// goes on, and continues the SAME search bursts in the object's tree; that tree's end link (RT_POP_LINK) is a
// fourth lane state handled in the primitive-test phase: restore the world ray, resume the top level.
// SPH: a profile-2 scene of spheres only (config 4: moving spheres, checker texture, lights): the general kernel's
// semantics with the sphere kernels' machinery - no primitive type dispatch, leaf payloads that name the sphere,
// three steps per vote (the weekend scene under profile 2: 80.0 ms through the full general kernel, 68.1 ms through
// the sphere kernel of profile 1).
template <int PROFILE, bool GENERAL, int SMEM, bool COUNT, bool EXT, bool TRACE = false, bool INST = false, bool SPH = false>
__global__ void __launch_bounds__(RT_BLOCK_OF(GENERAL), 1) k_render(const __grid_constant__ RenderParams P) {
  constexpr bool FH = RT_HALF_PLANES != 0; // fp16 planes + FHFMA in the shared-memory plans
  constexpr bool PRIMS = GENERAL && !SPH; // primitives of several types (type dispatch, rect lists, leaf_prims indirection)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  DevScene S = P.S;
  int off = 0;
  // 32-bit shared-window address of the staged node copies (they start the dynamic segment)
  // (asm volatile: computed ONCE. As a plain __cvta_generic_to_shared the compiler rematerialises it - S2UR CgaCtaId +
  // UMOV + ULEA, a special-register read on the critical path - at every burst entry and phase: config 2 65.24 -> 64.49 ms,
  // config 3 167.8 -> 166.3 ms per 150 spp, no additional spills.)
  uint32_t nodes_s = 0u;
  if (SMEM) asm volatile("{ .reg .u64 t; cvta.to.shared.u64 t, %1; cvt.u32.u64 %0, t; }" : "=r"(nodes_s) : "l"(smem_raw));
  if (SMEM) {
    // Nodes go to shared memory as 16-BYTE QUANTISED records (the global array keeps 32-byte float nodes):
    //   {x planes, y planes, z planes, link}: each plane word = two 16-bit coordinates over the root box
    //   (entry | exit << 16; fp16 numbers by default - see RT_HALF_PLANES above - or fixed-point integers), rounded
    //   outwards with two steps of padding, so the slab test stays conservative: a quantised box only ever adds
    //   false-positive box hits, never changes a result.
    // ONE LDS.128 per BVH step instead of two: ncu r1j had the shared-memory data pipe at 89 % of its
    // wavefront peak (13.5 wavefronts per step for 32 lanes fetching 32 bytes each at unrelated addresses).
    // One link word is enough in the depth-first threaded layout: the first child of an inner node and the
    // continuation after a leaf are both the NEXT record, so an inner node stores only its escape (absolute
    // 32-bit shared address of the target record in the same copy, 0 = traversal finished) and a leaf only
    // its payload (< 0). A sentinel record (empty box, link 0) follows the last node of every copy.
    // SMEM == 2: EIGHT copies, one per ray-direction octant (own front-to-back child order from the builder),
    // with entry / exit planes pre-selected for the octant's direction signs: the step needs no min/max
    // to order a slab's two planes.
    {
      S.nodes = (const float4 *)(smem_raw + off);
      const int nn = S.n_nodes, rec = INST ? nn : nn + 1; // (two-level scenes: every tree's slot ends with its own sentinel)
      const int copies = SMEM == 2 ? RT_N_ORDERINGS : P.n_copies;
      for (int i = threadIdx.x; i < copies * rec; i += blockDim.x) {
        const int q = i / rec, k = i - q * rec;
        uint4 w;
        // two-level scenes: the tree this record belongs to (binary search in the unit table), its frame, where its
        // kept nodes end (= its sentinel) and the code that ends it (0 = finished, RT_POP_SHARED = back to the top level)
        float fb[3] = {P.qbase[0], P.qbase[1], P.qbase[2]}, fi[3] = {P.qinv[0], P.qinv[1], P.qinv[2]};
        int u_end = nn, u_code = 0;
        if (INST) {
          int lo_u = 0, hi_u = P.n_units - 1;
          while (lo_u < hi_u) {
            const int mid = (lo_u + hi_u + 1) >> 1;
            if (RT_F2I(__ldg(P.units + 3 * mid).x) <= k) lo_u = mid; else hi_u = mid - 1;
          }
          const float4 u0 = __ldg(P.units + 3 * lo_u), u1 = __ldg(P.units + 3 * lo_u + 1), u2 = __ldg(P.units + 3 * lo_u + 2);
          u_end = RT_F2I(u0.x) + RT_F2I(u0.y);
          u_code = RT_F2I(u0.z);
          fb[0] = u1.x; fb[1] = u1.y; fb[2] = u1.z; fi[0] = u2.x; fi[1] = u2.y; fi[2] = u2.z;
        }
        if (INST ? k >= u_end : k == nn) {
          // sentinel: entry plane 65535, exit plane 0 on every axis: t(entry) > t(exit) strictly for every ray
          // with a finite non-zero 1/d, so it is always missed and its link (0) ends the traversal
          // (entry = the plane the ray reaches LAST: q = 65535 for a positive direction, q = 0 for a negative one)
          const unsigned fwd = (unsigned)RT_Q_MAX + 0x4B000000u, bwd = ((unsigned)RT_Q_MAX << 16) + 0x4B000000u;
          w = make_uint4((q & 1) ? bwd : fwd, (q & 2) ? bwd : fwd, (q & 4) ? bwd : fwd, 0u);
          // one-copy plans keep (low plane | high plane << 16): low = RT_Q_MAX, high = 0 is missed by either direction
          if (SMEM != 2) w = make_uint4((unsigned)RT_Q_MAX, (unsigned)RT_Q_MAX, (unsigned)RT_Q_MAX, 0u);
          if (FH) {
            // fp16: entry = +60000 | exit = -60000 for a positive direction, mirrored for a negative one (a slope is finite
            // or infinite, never zero or NaN: the entry lies beyond the exit for every ray); the one-copy plans keep
            // (low | high << 16) = (+60000 | -60000), which the per-lane swap turns into the same two cases
            const unsigned fwdh = 0xFB537B53u, bwdh = 0x7B53FB53u;
            w = make_uint4((SMEM == 2 && (q & 1)) ? bwdh : fwdh, (SMEM == 2 && (q & 2)) ? bwdh : fwdh, (SMEM == 2 && (q & 4)) ? bwdh : fwdh, 0u);
          }
          if (INST) w.w = (unsigned)u_code; // (records between a tree's sentinel and the next tree are never reached)
        } else {
          const float4 *src = P.S.nodes + (SMEM == 2 ? q : P.copy_oct[q]) * (P.S.node_stride >> 4) + 2 * k;
          const float4 lo = __ldg(src), hi = __ldg(src + 1);
          const float l3[3] = {lo.x, lo.y, lo.z}, h3[3] = {hi.x, hi.y, hi.z};
          unsigned pw[3];
#pragma unroll
          for (int a = 0; a < 3; a++) {
            const float fl = floorf((l3[a] - fb[a]) * fi[a]) - 2.0f, fh = ceilf((h3[a] - fb[a]) * fi[a]) + 2.0f;
            const unsigned ql = (unsigned)fminf(fmaxf(fl, 0.0f), (float)RT_Q_MAX), qh = (unsigned)fminf(fmaxf(fh, 0.0f), (float)RT_Q_MAX);
            const bool neg = SMEM == 2 && ((q >> a) & 1);
            pw[a] = (neg ? (qh | (ql << 16)) : (ql | (qh << 16))) + (SMEM == 2 ? 0x4B000000u : 0u);
            if (FH) { // v = q / 23, both roundings outwards
              const unsigned hl = __half_as_ushort(__float2half_rd(__fmul_rd((float)ql, 1.0f / RT_HALF_KINV)));
              const unsigned hh = __half_as_ushort(__float2half_ru(__fmul_ru((float)qh, 1.0f / RT_HALF_KINV)));
              pw[a] = neg ? (hh | (hl << 16)) : (hl | (hh << 16));
            }
          }
          const int base_q = (int)nodes_s + q * P.b_nodes;
          const int esc = RT_F2I(lo.w) >> RT_NODE_SHIFT; // node index within the copy
          int pay = RT_F2I(hi.w), link;
          if (pay < 0) { // leaf
            if (!PRIMS && P.direct_leaf) pay = ~(__ldg(P.S.leaf_prims + ((~pay) >> 3)) << 3);
            link = pay;
          } else if (INST) { // a link that leaves the tree is its end code (global form: finished / RT_POP_LINK)
            link = esc >= nn ? u_code : base_q + (esc << 4);
          } else { // inner: its first child is record k + 1 by construction
            link = esc == nn ? 0 : base_q + (esc << 4);
          }
          w = make_uint4(pw[0], pw[1], pw[2], (unsigned)link);
        }
        ((uint4 *)(smem_raw + off + q * P.b_nodes))[k] = w;
      }
      off += copies * P.b_nodes;
    }
    if (SMEM != 3) {
    S.sph = (const float4 *)stage_to_smem(smem_raw, off, P.S.sph, P.b_sph);
    S.sph_mat = (const int32_t *)stage_to_smem(smem_raw, off, P.S.sph_mat, P.b_sph_mat);
    S.sph_k = (const float *)stage_to_smem(smem_raw, off, P.S.sph_k, P.b_sph_k);
    S.mats = (const float4 *)stage_to_smem(smem_raw, off, P.S.mats, P.b_mats);
    S.big = (const int32_t *)stage_to_smem(smem_raw, off, P.S.big, P.b_big);
    S.leaf_prims = (const int32_t *)stage_to_smem(smem_raw, off, P.S.leaf_prims, P.b_leaf_prims);
    if (GENERAL) {
      S.sph_mv = (const float4 *)stage_to_smem(smem_raw, off, P.S.sph_mv, P.b_sph_mv);
      S.sph_t0 = (const float *)stage_to_smem(smem_raw, off, P.S.sph_t0, P.b_sph_t0);
      S.tri = (const float4 *)stage_to_smem(smem_raw, off, P.S.tri, P.b_tri);
      S.tri_n = (const float4 *)stage_to_smem(smem_raw, off, P.S.tri_n, P.b_tri_n);
      // (b_quad == 0 with rects present: they all but fit - the rects stay in global memory, read at shading time
      // and for the few that are not sides of a box; the box records below carry what the traversal needs)
      if (P.b_quad) S.quad = (const float4 *)stage_to_smem(smem_raw, off, P.S.quad, P.b_quad);
      S.tri_mat = (const int32_t *)stage_to_smem(smem_raw, off, P.S.tri_mat, P.b_tri_mat);
      if (P.b_quad_mat) S.quad_mat = (const int32_t *)stage_to_smem(smem_raw, off, P.S.quad_mat, P.b_quad_mat);
      S.bigq = (const float4 *)stage_to_smem(smem_raw, off, P.S.bigq, P.b_bigq);
      if (P.b_box) S.box = (const float4 *)stage_to_smem(smem_raw, off, P.S.box, P.b_box);
    }
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // Per-warp pixel accumulators, ORDER-INDEPENDENT: each channel of each pixel is a 64-bit fixed-point sum (32.32,
  // two 32-bit words: [2 tiles in flight][3 channels][lo, hi][32 pixels]) fed by native integer ATOMS.ADD with an explicit
  // carry. Integer addition commutes, so a pixel's sum does not depend on which lane traced which sample or in
  // which order the paths ended - run-to-run bit-identical frames although work items are handed out by a racing
  // atomic counter and overlap inside a warp. (The float atomicAdd this replaces is a CAS spin loop on shared
  // memory and made the fp32 summation order schedule-dependent.) Lanes address different pixels -> different banks.
  unsigned *acc = (unsigned *)(smem_raw + off) + warp * RT_ACC_WORDS;
  auto acc_add = [&](int pix_, V3f L_) {
    const float c3[3] = {L_.x, L_.y, L_.z};
    unsigned *a = acc + (pix_ >> 5) * 192 + (pix_ & 31);
#ifdef RT_ACC_FLOAT // measurement build: the schedule-dependent float sums this scheme replaced
    for (int c = 0; c < 3; c++) atomicAdd((float *)(a + c * 64), c3[c]);
    return;
#endif
#pragma unroll
    for (int c = 0; c < 3; c++) {
      // clamp to [0, 2^30]: fmaxf drops a NaN (-> 0), the upper bound keeps 2^32 * x inside 63 bits
      const long long q = __float2ll_rn(fminf(fmaxf(c3[c], 0.0f), 1073741824.0f) * 4294967296.0f);
      const unsigned lo = (unsigned)q, hi = (unsigned)((unsigned long long)q >> 32);
      const unsigned old = atomicAdd(a + c * 64, lo);
      const unsigned carry = (old + lo) < lo ? 1u : 0u;
      atomicAdd(a + c * 64 + 32, hi + carry);
    }
  };
  __syncthreads();

  const unsigned FULL = 0xffffffffu;
  const unsigned lt_mask = (1u << lane) - 1u;
  // Lane state machine with WARP-VOTED phases (design evaluated with tests/emu/warpsim:
  // 75 -> ~60 warp-instructions per segment). Every lane owns one path; its state is
  //   SEARCH  walking the threaded BVH (one node per step)
  //   LEAF    a hit leaf's primitives are pending
  //   DONE    traversal finished: shade / terminate
  //   DEAD    no path: take the next (pixel, sample) item of the tile's pool
  // The warp (1) steps the BVH while at least `frac` of the lanes that entered the burst
  // are still searching, (2) runs the pending primitive tests together, (3) shades and
  // regenerates once `batch` lanes are DONE/DEAD (or nobody is traversing) while the
  // remaining lanes keep their traversal state in registers.
  // Encoding: `alive` (the lane owns a path) and `node`:
  //   SEARCH   the next BVH node: its shared address (> 0; shared copies) or its byte offset
  //            in [0, node_end) (global path)
  //   node == node_end   traversal finished (DONE if alive, else DEAD); 0 for the shared copies
  //   node < 0           LEAF: a hit leaf is pending, node = its payload
  //                      ~(first << 3 | count-1) into leaf_prims; `resume` = where to go on
  const int node_end = SMEM ? 0 : (S.n_nodes << RT_NODE_SHIFT);
  const int node_pop = SMEM ? RT_POP_SHARED : RT_POP_LINK(S.n_nodes); // INST: the end link of an object's tree
#define RT_SEARCHING(n) (SMEM ? ((n) > (INST ? RT_POP_SHARED : 0)) : ((unsigned)(n) < (unsigned)node_end))
  bool alive = false;
  int node = node_end, resume = 0;
  HitAcc h;
  h.t = INFINITY; h.id = -1;
  RayPre pre;
  pre.inv_d = v3(0, 0, 0); pre.ood = v3(0, 0, 0); pre.inv_a = 0.f;
  // quantised-node slab test (shared copies), per axis: t(plane v) = v * qSh + qC, v and the slope pair qSh in fp16
  // (integer form: t(plane q) = (2^23 + q) * qS + qC)
  V3f qS = v3(0, 0, 0), qC = v3(0, 0, 0);
  // one-copy plans: per-lane byte-permute selectors (fp16 form: selE swaps the halves of a plane word for a negative
  // direction; integer form: selE / selX extract the entry / exit plane of each axis) - see the step
  struct U3 { unsigned x, y, z; } selE = {0x7610u, 0x7610u, 0x7610u}, selX = {0x7632u, 0x7632u, 0x7632u};
  U3 qSh = {0u, 0u, 0u}; // RT_HALF_PLANES: the slope per axis as an fp16 pair (entry: rounded down | exit: rounded up)
  const float t_min = P.sp.t_min;
  // Work items (tile, chunk of samples) are OVERLAPPED: when the current item's pool is
  // drained its in-flight paths become the "old" item and the warp starts regenerating
  // from the next item at once, so no lane idles through an item's drain tail. Two
  // 32-pixel accumulators per warp; a lane's path remembers its accumulator in bit 5 of pix.
  int pool_next = 0, pool_end = 0; // warp-uniform
  int path_score = 0;              // warp-uniform: > 0 = this warp's paths are long (see the shading threshold)
  int tile_x0 = 0, tile_y0 = 0, s0 = 0, chunk = 0, chunk_n = 0, cur_buf = 0;
  int old_x0 = 0, old_y0 = 0, old_chunk = 0, old_chunk_n = 0, old_buf = 0, old_inflight = 0;
  bool have_cur = false, old_valid = false, more_work = true;
  Ray r;
  r.o = v3(0, 0, 0); r.d = v3(0, 0, 1); r.tm = 0;
  // INST: while a lane walks an object's tree, `r` is the object-space ray and these hold the world ray, where the
  // top-level walk resumes, the instance being walked and the instance of the closest hit so far
  V3f w_o = v3(0, 0, 0), w_d = v3(0, 0, 1);
  int resume_top = 0, cur_inst = -1, hit_inst = -1;
  V3f beta = v3(1, 1, 1), L = v3(0, 0, 0);
  int bounce = 0, pix = 0, smp = 0, pixel_index = 0;
  unsigned n_seg = 0, n_paths = 0;
  TraceCounters cnt;
  cnt.box_tests = 0; cnt.prim_tests = 0;
  // node fetch: 2 x LDS.128 straight from a 32-bit shared address (no generic-pointer
  // arithmetic in the loop); global path for scenes that do not fit shared memory
  const char *nodes_g = (const char *)S.nodes;
  const char *nodes_q = nodes_g; // SMEM == 0: the current ray's quadrant copy in global memory

  // Point the lane at the root of a tree for the ray in `r` / `pre`: root_index = the root's record index (0 = the
  // scene's / top level's; an object's tree of a two-level scene otherwise), fb / fs = that tree's quantisation frame.
  auto enter_tree = [&](int root_index, V3f fb, V3f fs) {
    // sign BITS of 1/d (covers d = -0): which octant ordering (own child order; in shared memory also
    // pre-selected entry / exit planes) this ray walks
    const unsigned octant = ((unsigned)RT_F2I(pre.inv_d.x) >> 31) | (((unsigned)RT_F2I(pre.inv_d.y) >> 31) << 1) |
                            (((unsigned)RT_F2I(pre.inv_d.z) >> 31) << 2);
    if (SMEM == 0) {
      node = root_index << RT_NODE_SHIFT;
      nodes_q = nodes_g + octant * (unsigned)S.node_stride;
    } else { // the root's shared address in this ray's copy
      const unsigned copy = SMEM == 2 ? octant : ((P.oct_lut >> (4u * octant)) & 7u);
      node = (int)(nodes_s + copy * (unsigned)P.b_nodes) + (root_index << 4);
      // (integer form: plane q enters the slab test as the float 2^23 + q - one PRMT builds it: 0x4B00 | q -, so
      // t = (qbase + q * qscale - o) / d = (2^23 + q) * qS + qC with qS = qscale / d, qC = (qbase - o) / d - 2^23 qS)
      qS = v3(fs.x * pre.inv_d.x, fs.y * pre.inv_d.y, fs.z * pre.inv_d.z);
      if (FH) {
        // x = fb + 23 v fs, so t = v * (23 fs / d) + (fb - o) / d; the slope as the fp16 pair (rounded down | rounded up << 16)
        const V3f sl = RT_HALF_KINV * qS;
        qC = v3((fb.x - r.o.x) * pre.inv_d.x, (fb.y - r.o.y) * pre.inv_d.y, (fb.z - r.o.z) * pre.inv_d.z);
        qSh.x = (unsigned)__half_as_ushort(__float2half_rd(sl.x)) | ((unsigned)__half_as_ushort(__float2half_ru(sl.x)) << 16);
        qSh.y = (unsigned)__half_as_ushort(__float2half_rd(sl.y)) | ((unsigned)__half_as_ushort(__float2half_ru(sl.y)) << 16);
        qSh.z = (unsigned)__half_as_ushort(__float2half_rd(sl.z)) | ((unsigned)__half_as_ushort(__float2half_ru(sl.z)) << 16);
      } else
      qC = v3(RT_FMA(-8388608.0f, qS.x, (fb.x - r.o.x) * pre.inv_d.x), RT_FMA(-8388608.0f, qS.y, (fb.y - r.o.y) * pre.inv_d.y),
              RT_FMA(-8388608.0f, qS.z, (fb.z - r.o.z) * pre.inv_d.z));
      if (SMEM != 2 && FH) { // a negative direction enters through the HIGH plane: swap the halves of the word
        selE.x = (octant & 1u) ? 0x1032u : 0x3210u;
        selE.y = (octant & 2u) ? 0x1032u : 0x3210u;
        selE.z = (octant & 4u) ? 0x1032u : 0x3210u;
      } else if (SMEM != 2) {
        selE.x = (octant & 1u) ? 0x7632u : 0x7610u; selX.x = selE.x ^ 0x0022u;
        selE.y = (octant & 2u) ? 0x7632u : 0x7610u; selX.y = selE.y ^ 0x0022u;
        selE.z = (octant & 4u) ? 0x7632u : 0x7610u; selX.z = selE.z ^ 0x0022u;
      }
    }
  };
  // start a segment: per-ray constants, the always-tested big primitives, enter at the root
  unsigned media_seed = 0u; // EXT: the fourth word of the Philox block that produced the current ray (apply_media)
  auto begin_segment = [&]() {
    pre = ray_precompute_fast(r);
    h.t = TRACE ? P.trace_tmax : INFINITY; h.id = -1;
    if (INST) { cur_inst = -1; hit_inst = -1; }
    for (int i = 0; i < S.n_big; i++) {
      const int32_t id = S.big[i];
      if (COUNT) cnt.prim_tests++;
      const bool is_sphere = !PRIMS || RT_PRIM_TYPE_OF(id) == RT_PRIM_SPHERE;
      if (is_sphere) hit_prim<PROFILE, GENERAL, true, SPH>(S, id, r, pre, t_min, h);
      else hit_prim<PROFILE, GENERAL, false, SPH>(S, id, r, pre, t_min, h);
    }
    if (PRIMS) { // always-tested rects (walls): decoded records, the same address for every lane
      for (int i = 0; i < S.n_bigq; i++) {
        if (COUNT) cnt.prim_tests++;
        const float4 q0 = S.bigq[2 * i], q1 = S.bigq[2 * i + 1];
        hit_quad(q0, q1, r, pre, t_min, h, RT_F2I(q1.z));
      }
    }
    // participating media: a sampled scatter event becomes the initial closest hit
    if (EXT && S.n_media) h = apply_media(S.media, S.n_media, r.o, r.d, media_seed, h);
    enter_tree(0, v3(P.qbase[0], P.qbase[1], P.qbase[2]), v3(P.qscale[0], P.qscale[1], P.qscale[2]));
    if (S.n_nodes == 0) node = node_end; // no tree at all: finished at once
  };

  for (;;) {
    // ---- phase 1: a burst of BVH steps (branch-free body: lanes that are not searching
    // fetch node 0 and discard the result)
    const int ns0 = __popc(__ballot_sync(FULL, RT_SEARCHING(node)));
    if (ns0 > 0) {
      const int thr = max(1, (ns0 * P.frac8) >> 3);
      for (;;) {
        if (__popc(__ballot_sync(FULL, RT_SEARCHING(node))) < thr) break;
#pragma unroll
        for (int u = 0; u < RT_STEPS_PER_VOTE(PRIMS, EXT); u++) {
          const bool searching = RT_SEARCHING(node);
          if (SMEM != 0) {
            // lanes that are not searching (node <= 0) fetch the first record and discard the result
            const int at = max(node, (int)nodes_s);
            unsigned w0, w1, w2;
            int link;
            asm("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(link) : "r"(at));
            if (COUNT) cnt.box_tests += searching ? 1u : 0u;
            // Integer form (RT_HALF_PLANES=0), octant copies: decode on the FMA pipe (ncu r2e: with PRMT the half-rate ALU
            // pipe was the limiter at 74 %, the FMA pipe at 25 %): the stored word is (exit << 16 | entry) + 0x4B000000, so
            //   exit bits  = mad.hi(w, 2^16, 0x4B000000 - 0x4B00) = 0x4B000000 + exit      (planes <= 46335: no carry)
            //   entry bits = mad.lo(exit bits, -2^16, w)          = 0x4B000000 + entry     (mod 2^32)
            // i.e. the floats 2^23 + q, two integer multiply-adds per word.
            unsigned bx1 = 0, by1 = 0, bz1 = 0, bx0 = 0, by0 = 0, bz0 = 0;
            float x0, x1, y0, y1, z0, z1;
            if (FH) {
              // FHFMA: t = plane (fp16 half of the record word) * slope (fp16 half of the pair) + C, fp32 result
              if (SMEM != 2) { // one copy for every direction: each lane puts its entry plane into the low half
                asm("prmt.b32 %0, %0, %0, %1;" : "+r"(w0) : "r"(selE.x));
                asm("prmt.b32 %0, %0, %0, %1;" : "+r"(w1) : "r"(selE.y));
                asm("prmt.b32 %0, %0, %0, %1;" : "+r"(w2) : "r"(selE.z));
              }
#define RT_FHFMA(d, w, sh, c, HALF)                                                                                      \
  asm("{ .reg .b16 pl, ph, sl, sx; mov.b32 {pl, ph}, %1; mov.b32 {sl, sx}, %2; fma.rn.f32.f16 %0, " HALF ", %3; }"   \
      : "=f"(d) : "r"(w), "r"(sh), "f"(c))
              RT_FHFMA(x0, w0, qSh.x, qC.x, "pl, sl"); RT_FHFMA(x1, w0, qSh.x, qC.x, "ph, sx");
              RT_FHFMA(y0, w1, qSh.y, qC.y, "pl, sl"); RT_FHFMA(y1, w1, qSh.y, qC.y, "ph, sx");
              RT_FHFMA(z0, w2, qSh.z, qC.z, "pl, sl"); RT_FHFMA(z1, w2, qSh.z, qC.z, "ph, sx");
#undef RT_FHFMA
            } else if (SMEM == 2) {
              asm("mad.hi.u32 %0, %1, 65536, 0x4AFFB500;" : "=r"(bx1) : "r"(w0));
              asm("mad.hi.u32 %0, %1, 65536, 0x4AFFB500;" : "=r"(by1) : "r"(w1));
              asm("mad.hi.u32 %0, %1, 65536, 0x4AFFB500;" : "=r"(bz1) : "r"(w2));
              asm("mad.lo.u32 %0, %1, 0xFFFF0000, %2;" : "=r"(bx0) : "r"(bx1), "r"(w0));
              asm("mad.lo.u32 %0, %1, 0xFFFF0000, %2;" : "=r"(by0) : "r"(by1), "r"(w1));
              asm("mad.lo.u32 %0, %1, 0xFFFF0000, %2;" : "=r"(bz0) : "r"(bz1), "r"(w2));
            } else {
              // One copy for every direction: the word is (low plane | high plane << 16) and each LANE picks its entry
              // and exit half with a byte permute whose selector depends on the sign of its direction (0x7610: low
              // half under 0x4B00, 0x7632: high half) - extraction and selection in one instruction, so the slab
              // needs no min/max to order its two planes here either.
              asm("prmt.b32 %0, %1, 0x4B000000, %2;" : "=r"(bx0) : "r"(w0), "r"(selE.x));
              asm("prmt.b32 %0, %1, 0x4B000000, %2;" : "=r"(by0) : "r"(w1), "r"(selE.y));
              asm("prmt.b32 %0, %1, 0x4B000000, %2;" : "=r"(bz0) : "r"(w2), "r"(selE.z));
              asm("prmt.b32 %0, %1, 0x4B000000, %2;" : "=r"(bx1) : "r"(w0), "r"(selX.x));
              asm("prmt.b32 %0, %1, 0x4B000000, %2;" : "=r"(by1) : "r"(w1), "r"(selX.y));
              asm("prmt.b32 %0, %1, 0x4B000000, %2;" : "=r"(bz1) : "r"(w2), "r"(selX.z));
            }
            if (!FH) {
              x0 = RT_FMA(__uint_as_float(bx0), qS.x, qC.x); x1 = RT_FMA(__uint_as_float(bx1), qS.x, qC.x);
              y0 = RT_FMA(__uint_as_float(by0), qS.y, qC.y); y1 = RT_FMA(__uint_as_float(by1), qS.y, qC.y);
              z0 = RT_FMA(__uint_as_float(bz0), qS.z, qC.z); z1 = RT_FMA(__uint_as_float(bz1), qS.z, qC.z);
            }
            // x0 / y0 / z0 are the entry planes: pre-selected per octant copy (SMEM == 2) or per lane (one copy)
            const float tn = RT_FMAX(RT_FMAX(x0, y0), RT_FMAX(z0, t_min));
            const float tf = RT_FMIN(RT_FMIN(x1, y1), RT_FMIN(z1, h.t));
            // hit inner -> next record; hit leaf -> its payload; missed inner -> escape; missed leaf -> next record.
            // The record after the last node is the sentinel (always missed, link 0 = finished).
            const int cont = at + 16;
            const bool leaf = link < 0;
            const int nxt = ((tn <= tf) == leaf) ? link : cont;
            if (searching) { node = nxt; resume = cont; }
          } else {
            const int at = searching ? node : 0;
            const float4 lo = __ldg((const float4 *)(nodes_q + at));
            const float4 hi = __ldg((const float4 *)(nodes_q + at) + 1);
            if (COUNT) cnt.box_tests += searching ? 1u : 0u;
            const int esc = RT_F2I(lo.w), pay = RT_F2I(hi.w);
            const bool bh = hit_box(lo, hi, pre, t_min, h.t);
            const int nxt = bh ? pay : esc; // inner: child; leaf: payload < 0
            if (searching) { node = nxt; resume = esc; }
          }
        }
      }
    }
    // ---- phase 2: pending primitive tests (one per lane per round)
    const unsigned m_leaf = __ballot_sync(FULL, node < 0 || (INST && node == node_pop));
    if (m_leaf && (__popc(m_leaf) >= P.leaf_min || !__ballot_sync(FULL, RT_SEARCHING(node)))) {
      if (INST && node == node_pop) { // the object's tree is done: back to the top level with the world ray
        r.o = w_o; r.d = w_d;
        pre = ray_precompute_fast(r);
        enter_tree(0, v3(P.qbase[0], P.qbase[1], P.qbase[2]), v3(P.qscale[0], P.qscale[1], P.qscale[2]));
        node = resume_top; // (the same octant copy as before the instance: the world ray has not changed)
        cur_inst = -1;
      } else if (INST && node < 0 && RT_PRIM_TYPE_OF(S.leaf_prims[(~node) >> 3]) == RT_PRIM_INSTANCE) {
        // enter the instance named by this top-level leaf (instance leaves hold exactly one instance)
        const int k = RT_PRIM_INDEX_OF(S.leaf_prims[(~node) >> 3]);
        const float4 *rec = S.inst + RT_INST_STRIDE * k;
        w_o = r.o; w_d = r.d;
        resume_top = resume;
        r.o = inst_point_to_object(rec, w_o);
        r.d = inst_vector_to_object(rec, w_d);
        pre = ray_precompute_fast(r);
        enter_tree(RT_F2I(rec[3].x) >> RT_NODE_SHIFT, xyz(rec[4]), xyz(rec[5]));
        cur_inst = k;
      } else if (node < 0) {
        const int enc = ~node;
        if (COUNT) cnt.prim_tests++;
        const HitAcc h_before = h;
        if (!PRIMS && SMEM != 0 && P.direct_leaf) {
          // single-sphere leaf named by the payload itself: one LDS.128 from a 32-bit shared address
          const int id = enc >> 3;
          float4 s4;
          asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(s4.x), "=f"(s4.y), "=f"(s4.z), "=f"(s4.w)
              : "r"(nodes_s + (unsigned)P.off_sph + ((unsigned)id << 4)));
          V3f c = xyz(s4);
          bool closed = PROFILE == 0;
          if (GENERAL && S.any_moving) { // moving_sphere::center (moving_sphere.h:34-36); its interval is closed
            const float4 mv = S.sph_mv[id];
            if (mv.w != 0.0f) { c = sphere_center_at(s4, mv, S.sph_t0[id], r.tm); closed = true; }
          }
          hit_sphere(s4, c, closed, r, pre, t_min, h, id);
          node = resume;
        } else {
          hit_prim<PROFILE, GENERAL, false, SPH>(S, S.leaf_prims[enc >> 3], r, pre, t_min, h);
          node = (enc & 7) ? node - 7 : resume; // next primitive of the leaf (first+1, count-1) or go on
        }
        if (INST && (h.id != h_before.id || h.t != h_before.t)) hit_inst = cur_inst;
      }
    }
    // ---- phase 3: shade + regenerate once enough lanes are out of the traversal
    const unsigned m_out = __ballot_sync(FULL, node == node_end);
    const int n_out = __popc(m_out);
    // Long paths (closed, lit rooms: config 3 runs 20 segments per path) regenerate rarely - most lanes that leave the
    // traversal only need the cheap hit shading - and do better with an earlier shading round; short paths end often
    // and the expensive regeneration code wants a fuller batch. Each warp keeps a running score of (continuing hits)
    // against (new paths) and switches thresholds: config 3 185 -> 172 ms per 150 spp, short-path scenes unchanged.
    const int batch_w = (GENERAL && !EXT && path_score > 2048) ? P.batch_long : P.batch; // (the extended kernel: -3.5 % with it)
    if (n_out < batch_w && n_out < 32) continue;
    // -- work-item management (warp-uniform)
    if (pool_next >= pool_end && !old_valid) {
      if (have_cur) { // the current item's pool is drained: it becomes the old item
        const unsigned infl = __ballot_sync(FULL, alive);
        old_x0 = tile_x0; old_y0 = tile_y0; old_chunk = chunk; old_chunk_n = chunk_n; old_buf = cur_buf;
        old_inflight = __popc(infl);
        old_valid = true; // flushed below as soon as old_inflight == 0
        have_cur = false;
      }
      if (more_work) {
        int w = 0;
        if (lane == 0) w = atomicAdd(P.work_counter, 1);
        w = __shfl_sync(FULL, w, 0);
        if (w >= P.n_work) {
          more_work = false;
        } else {
          const int tile = w % P.n_tiles;
          chunk = w / P.n_tiles;
          tile_x0 = (tile % P.tiles_x) * RT_TILE_W;
          tile_y0 = P.y0 + (tile / P.tiles_x) * RT_TILE_H;
          s0 = P.spp_begin + P.chunk_begin[chunk];
          chunk_n = P.chunk_begin[chunk + 1] - P.chunk_begin[chunk];
          pool_next = 0;
          pool_end = chunk_n * 32;
          if (TRACE) { // item w = the rays [w * trace_item, (w + 1) * trace_item)
            s0 = w * P.trace_item;
            pool_end = min(P.trace_item, P.n_rays - s0);
          }
          cur_buf = old_buf ^ 1;
          have_cur = true;
          unsigned *a = acc + cur_buf * 192 + lane;
#pragma unroll
          for (int k = 0; k < 6; k++) a[k * 32] = 0u;
          __syncwarp();
        }
      }
    }
    const bool pool_has = pool_next < pool_end;
    const bool done = alive && node == node_end;
    bool ended = false; // this lane's path ended in this round
    if (__ballot_sync(FULL, done || (!alive && pool_has))) {
      // (a) rays that left the scene end their path now
      const bool hit = !TRACE && done && h.id >= 0;
      if (done) {
        n_seg++;
        if (TRACE) { // the traversal's result IS the output
          P.trace_id[pixel_index] = h.id;
          P.trace_t[pixel_index] = h.id >= 0 ? h.t : 0.f;
          if (INST && P.trace_inst) P.trace_inst[pixel_index] = h.id >= 0 ? hit_inst : -1;
          alive = false;
          ended = true;
        } else if (!hit) {
          L = L + beta * miss_radiance(P.sp, r.d);
          acc_add(pix, L); // (one accumulation site per way of ending: a merged site after the shading measured 2 % slower)
          alive = false;
          ended = true;
        }
      }
      // (b) one Philox call for the bounce event of every hit lane
      Philox4 q;
      q.x = q.y = q.z = q.w = 0u;
      // dead lanes take the next (pixel, sample) items of the current pool by ballot rank
      bool fresh_path = false;
      int old_pix = pix;
      const unsigned dead = __ballot_sync(FULL, !alive);
      if (pool_has && dead) {
        if (!alive) {
          const int item = pool_next + __popc(dead & lt_mask);
          if (TRACE) {
            if (item < pool_end) {
              pix = cur_buf << 5;
              pixel_index = s0 + item; // the ray's index
              fresh_path = true;
            }
          } else if (item < pool_end) {
            const int px = item & 31;
            const int i = tile_x0 + (px & 7), j = tile_y0 + (px >> 3);
            if (i < P.W && j < P.y1) {
              pix = px | (cur_buf << 5);
              smp = s0 + (item >> 5);
              pixel_index = j * P.W + i;
              bounce = -1; // event index 0 = the camera ray
              fresh_path = true;
            }
          }
        }
        pool_next = min(pool_end, pool_next + __popc(dead));
      }
      // (c) ONE Philox call per lane: the bounce event of a hit, or the camera event of a new path
      if (!TRACE && (hit || fresh_path))
        q = philox_for_kernel<GENERAL>((uint32_t)pixel_index, (uint32_t)smp, (uint32_t)(1 + bounce), 0u, P.seed_lo, P.seed_hi, P.philox_rk);
      bool fresh_ray = false;
      if (hit) {
        const bool cont = shade_hit<PROFILE, GENERAL, EXT, INST, SPH>(S, P.sp, r, h, beta, L, q, hit_inst);
        bounce++;
        if (cont && bounce < P.sp.max_depth) {
          fresh_ray = true;
        } else {
          if (cont) { // depth exhausted
            if (PROFILE == 2) {
              if (P.sp.flags & RT_FLAG_DEPTH_BACKGROUND) L = L + beta * P.sp.background; // obj_render.cu:78-83
              else L = P.sp.background;                                                    // main.cu:104
            } else {
              L = v3(0, 0, 0); // main.cpp:58-60, final.cu:53
            }
          } // else: absorbed (L is 0 for profiles 0/1) or hit a light
          acc_add(pix, L);
          alive = false; // regenerated in the next round
          ended = true;
          node = node_end;
        }
      }
      // ONE accumulation site for every path that ended in this round - rays that left the scene in (a), possibly
      // already re-assigned to a new (pixel, sample) whose L is reset only below, and paths the shading terminated
      if (GENERAL && !EXT && !TRACE) { // hits that go on count 1, a new path counts 10 the other way, clamped
        path_score += __popc(__ballot_sync(FULL, hit)) - 10 * __popc(__ballot_sync(FULL, fresh_path));
        path_score = max(-4096, min(4096, path_score));
      }
      if (TRACE && fresh_path) { // (a hit lane keeps its path: it never takes a new item in the same round)
        const float4 ra = __ldg(P.trace_rays + 2 * (size_t)pixel_index), rb = __ldg(P.trace_rays + 2 * (size_t)pixel_index + 1);
        r.o = v3(ra.x, ra.y, ra.z); r.tm = ra.w; r.d = v3(rb.x, rb.y, rb.z);
        alive = true;
        n_paths++;
        fresh_ray = true;
      } else if (fresh_path) {
        const int px = pix & 31;
        const int i = tile_x0 + (px & 7), j = tile_y0 + (px >> 3);
        float x5 = 0.f;
        if (PROFILE == 2 && P.cam.time1 != P.cam.time0)
          x5 = u01(philox_for_kernel<GENERAL>((uint32_t)pixel_index, (uint32_t)smp, 0u, 1u, P.seed_lo, P.seed_hi, P.philox_rk).x);
        r = gen_camera_ray<PROFILE>(P.cam, P.W, P.H, i, j, u01(q.x), u01(q.y), u01(q.z), u01(q.w), x5);
        beta = v3(1, 1, 1);
        L = v3(0, 0, 0);
        bounce = 0;
        alive = true;
        n_paths++;
        fresh_ray = true;
      }
      // (d) every lane with a new ray starts its segment together
      if (EXT) media_seed = q.w;
      if (fresh_ray) begin_segment();
      // paths of the old item that ended in this round (their pix was captured before (b))
      if (old_valid) old_inflight -= __popc(__ballot_sync(FULL, ended && ((old_pix >> 5) == old_buf)));
    }
    // -- retire the old item once its last path has ended: lane <-> pixel, coalesced float4
    if (old_valid && old_inflight == 0) {
      __syncwarp();
      const int i = old_x0 + (lane & 7), j = old_y0 + (lane >> 3);
      if (!TRACE && i < P.W && j < P.y1) {
        const unsigned *a = acc + old_buf * 192 + lane;
        float c3[3];
#pragma unroll
        for (int c = 0; c < 3; c++) // one rounding: the exact 64-bit sum -> float
#ifdef RT_ACC_FLOAT
          c3[c] = __uint_as_float(a[c * 64]);
#else
          c3[c] = __ull2float_rn(((unsigned long long)a[c * 64 + 32] << 32) | a[c * 64]) * 2.3283064365386963e-10f;
#endif
        P.partial[(size_t)old_chunk * ((size_t)P.W * P.H) + (size_t)j * P.W + i] =
            make_float4(c3[0], c3[1], c3[2], (float)old_chunk_n);
      }
      old_valid = false;
    }
    if (!have_cur && !old_valid && !more_work) break;
  }
  // statistics: one atomic per warp
  unsigned long long a = n_paths, b = n_seg, c = cnt.box_tests, d = cnt.prim_tests;
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_down_sync(FULL, a, o);
    b += __shfl_down_sync(FULL, b, o);
    if (COUNT) { c += __shfl_down_sync(FULL, c, o); d += __shfl_down_sync(FULL, d, o); }
  }
  if (lane == 0) {
    atomicAdd(&P.stats[0], a);
    atomicAdd(&P.stats[1], b);
    if (COUNT) { atomicAdd(&P.stats[2], c); atomicAdd(&P.stats[3], d); }
  }
}

// accum[p] += sum over chunks (fixed order) of partial[c][p] for the pixels [first, first + count)
__global__ void k_combine(float4 *__restrict__ accum, const float4 *__restrict__ partial, int n_pix, int n_chunks,
                          int first, int count) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= count) return;
  p += first;
  float4 a = accum[p];
  for (int c = 0; c < n_chunks; c++) {
    float4 v = __ldcs(partial + (size_t)c * n_pix + p);
    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
  }
  accum[p] = a;
}

// Resolve. linear: [H][W][3] bottom row first (may be null). rgb8: [H][W][3] TOP row
// first (PPM order; may be null).
__global__ void k_resolve(const float4 *__restrict__ accum, int W, int H, int profile, float *__restrict__ linear,
                          uint8_t *__restrict__ rgb8) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= W * H) return;
  float4 a = accum[p];
  float inv = a.w > 0.f ? 1.0f / a.w : 0.f;
  float c[3] = {a.x * inv, a.y * inv, a.z * inv};
  if (linear) { linear[3 * p + 0] = c[0]; linear[3 * p + 1] = c[1]; linear[3 * p + 2] = c[2]; }
  if (rgb8) {
    int i = p % W, j = p / W;
    size_t o = ((size_t)(H - 1 - j) * W + i) * 3;
    for (int k = 0; k < 3; k++) {
      float g = sqrtf(fmaxf(c[k], 0.f));
      int q;
      if (profile == 0) q = (int)(256.0f * fminf(fmaxf(g, 0.0f), 0.999f)); // color.h:21-27
      else q = min(255, (int)(255.99f * g));                                // final.cu:227-229 (clamped)
      rgb8[o + k] = (uint8_t)q;
    }
  }
}

// Denoiser feature buffers: one thread per pixel, `spp` camera samples with the same Philox camera
// streams as k_render (event 0, streams 0 and 1), first SURFACE hit only (media are ignored).
// out: float[H*W][8] = albedo.rgb, normal.xyz, t, hit fraction.
template <int PROFILE, bool GENERAL, bool INST = false>
RT_HD void aov_pixel(const DevScene &S, const DevCamera &cam, const ShadeParams &sp, int W, int H, int spp, uint32_t seed_lo,
                     uint32_t seed_hi, int p, float *out) {
  const int i = p % W, j = p / W;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int smp = 0; smp < spp; smp++) {
    const Philox4 q = philox4x32_10((uint32_t)p, (uint32_t)smp, 0u, 0u, seed_lo, seed_hi);
    float x5 = 0.f;
    if (PROFILE == 2 && cam.time1 != cam.time0) x5 = u01(philox4x32_10((uint32_t)p, (uint32_t)smp, 0u, 1u, seed_lo, seed_hi).x);
    const Ray r = gen_camera_ray<PROFILE>(cam, W, H, i, j, u01(q.x), u01(q.y), u01(q.z), u01(q.w), x5);
    int inst = -1;
    const HitAcc h = trace_closest<PROFILE, GENERAL, false, INST>(S, r, sp.t_min, INFINITY, nullptr, &inst);
    V3f albedo, normal = v3(0, 0, 0);
    if (h.id >= 0) {
      first_hit_features<PROFILE, GENERAL, INST>(S, sp, r, h, albedo, normal, inst);
      acc[6] += h.t; acc[7] += 1.0f;
    } else {
      albedo = miss_radiance(sp, r.d);
    }
    acc[0] += albedo.x; acc[1] += albedo.y; acc[2] += albedo.z;
    acc[3] += normal.x; acc[4] += normal.y; acc[5] += normal.z;
  }
  const float inv = 1.0f / (float)spp;
  for (int k = 0; k < 8; k++) out[8 * (size_t)p + k] = acc[k] * inv;
}

template <int PROFILE, bool GENERAL, bool INST = false>
__global__ void k_aov(const DevScene S, const DevCamera cam, const ShadeParams sp, int W, int H, int spp, uint32_t seed_lo,
                      uint32_t seed_hi, float *__restrict__ out) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < W * H) aov_pixel<PROFILE, GENERAL, INST>(S, cam, sp, W, H, spp, seed_lo, seed_hi, p, out);
}

template <int PROFILE, bool GENERAL, bool INST = false>
__global__ void k_trace_closest(const DevScene S, const uint8_t *__restrict__ sphere_is_big,
                                const float4 *__restrict__ rays, int n, float t_min, float t_max, int use_accel,
                                int32_t *__restrict__ out_id, float *__restrict__ out_t,
                                const uint8_t *__restrict__ grouped = nullptr, int32_t *__restrict__ out_inst = nullptr) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  float4 a = rays[2 * k], b = rays[2 * k + 1];
  Ray r;
  r.o = v3(a.x, a.y, a.z); r.tm = a.w; r.d = v3(b.x, b.y, b.z);
  HitAcc h;
  int inst = -1;
  if (use_accel) h = trace_closest<PROFILE, GENERAL, false, INST>(S, r, t_min, t_max, nullptr, &inst);
  else h = trace_brute<PROFILE, GENERAL, INST>(S, sphere_is_big, r, t_min, t_max, grouped, &inst);
  out_id[k] = h.id;
  out_t[k] = h.id >= 0 ? h.t : 0.f;
  if (INST && out_inst) out_inst[k] = h.id >= 0 ? inst : -1;
}
