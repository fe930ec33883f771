// philox.cuh — counter-based Philox4x32-10 (Salmon et al., SC'11) replacing the
// reference's per-pixel curandState (XORWOW, 48 B/pixel; render_init
// accelerated-rt-cuda/final.cu:62-73) and the CPU's global rand()
// (rt_in_one_weekend/rtweekend.h:21-24). Stateless: the four 32-bit words
// drawn for an event are a pure function of
//   counter = (pixel index, global sample index, event index, stream)
//   key     = 64-bit seed
// so an image does not depend on which GPU / launch rendered which sample.
#pragma once
#include "rt_common.cuh"

struct Philox4 {
  uint32_t x, y, z, w;
};

RT_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t hi0 = RT_MULHI(M0, c0), lo0 = M0 * c0;
    uint32_t hi1 = RT_MULHI(M1, c2), lo1 = M1 * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += W0; k1 += W1;
  }
  Philox4 o = {c0, c1, c2, c3};
  return o;
}

// The same function with the ten round keys given (k + r * W, r = 0..9: {k0, k1} pairs). The key schedule is the same for
// every lane, warp and call of a launch: precomputed on the host and read as kernel-parameter constants it costs nothing,
// where the bump per round is two (uniform-datapath) instructions that still take issue slots - 20 of a call's ~80.
RT_HD Philox4 philox4x32_10_keys(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const uint32_t *rk) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t hi0 = RT_MULHI(M0, c0), lo0 = M0 * c0;
    uint32_t hi1 = RT_MULHI(M1, c2), lo1 = M1 * c2;
    uint32_t n0 = hi1 ^ c1 ^ rk[2 * r], n2 = hi0 ^ c3 ^ rk[2 * r + 1];
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
  }
  Philox4 o = {c0, c1, c2, c3};
  return o;
}
static inline void philox_round_keys(uint32_t k0, uint32_t k1, uint32_t rk[20]) {
  for (int r = 0; r < 10; r++) { rk[2 * r] = k0 + (uint32_t)r * 0x9E3779B9u; rk[2 * r + 1] = k1 + (uint32_t)r * 0xBB67AE85u; }
}

// The render kernel's call sites (camera event, bounce event, shutter time). OUTLINE: calls of ONE out-of-line copy
// instead of three inlined ones - the general kernels are bound by instruction fetch (rt_next_week final scene
// 106.9 -> 95.9 ms per 200 spp; configs 3 and 4 unchanged); the compact sphere-only kernels keep it inline.
#ifdef __CUDA_ARCH__
static __device__ __noinline__ Philox4 philox_call(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
  return philox4x32_10(c0, c1, c2, c3, k0, k1);
}
#endif
template <bool OUTLINE>
RT_HD Philox4 philox_for_kernel(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                const uint32_t *rk = nullptr) {
#ifdef __CUDA_ARCH__
#ifdef RT_PHILOX_OUTLINE_ALL
  return philox_call(c0, c1, c2, c3, k0, k1);
#else
  if (OUTLINE) return philox_call(c0, c1, c2, c3, k0, k1);
#endif
  if (rk) return philox4x32_10_keys(c0, c1, c2, c3, rk); // inlined copies: round keys from the kernel parameters
#endif
  return philox4x32_10(c0, c1, c2, c3, k0, k1);
}

// uniform in [0,1): top 24 bits (the reference's CPU generator is [0,1) too,
// rtweekend.h:21-24; curand_uniform is (0,1] — immaterial for the estimators).
RT_HD float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

// ---- direct (rejection-free) samplers with the SAME distributions as the
// reference's rejection loops, so that every event consumes a fixed number of
// random words (one Philox call per event, no divergent retry loops):
//   random_in_unit_disk   vec3.h:123-130            -> sqrt(u)-radius polar
//   random_unit_vector    vec3.h:112                -> uniform sphere (z, phi)
//   random_in_unit_sphere vec3.h:103-110, material.h:15-21 (CUDA)
//                                                   -> unit vector * cbrt(u)
RT_HD void rt_sincos_2pi(float u, float *s, float *c) {
  const float phi = RT_FMA(u, 6.283185307179586f, -3.14159265358979f); // [-pi, pi)
#ifdef __CUDA_ARCH__
  __sincosf(phi, s, c); // MUFU path: abs error ~2^-21 on [-pi, pi]
#else
  *s = sinf(phi); *c = cosf(phi);
#endif
}
RT_HD void sample_unit_disk(float u1, float u2, float *x, float *y) {
  float r = RT_SQRT(u1), s, c;
  rt_sincos_2pi(u2, &s, &c);
  *x = r * c; *y = r * s;
}
RT_HD V3f sample_unit_vector(float u1, float u2) {
  float z = RT_FMA(-2.0f, u1, 1.0f);
  float r = RT_SQRT(RT_FMAX(0.0f, RT_FMA(-z, z, 1.0f))), s, c;
  rt_sincos_2pi(u2, &s, &c);
  return v3(r * c, r * s, z);
}
RT_HD V3f sample_unit_ball(float u1, float u2, float u3) {
#ifdef __CUDA_ARCH__
  float rad = exp2f(__log2f(u3) * (1.0f / 3.0f)); // u3 = 0 -> 0
#else
  float rad = cbrtf(u3);
#endif
  return rad * sample_unit_vector(u1, u2);
}
