/* oracle/rt_oracle_impl.h — ORACLE L1 body (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the reference's path-tracing algorithm, included twice
 * by rt_oracle.c: REAL=double (FN = orc64_*, the CPU renderer's arithmetic) and
 * REAL=float (FN = orc32_*, the CUDA renderers' arithmetic). Every function
 * cites the reference file:line it follows. It consumes the same flattened
 * scene (include/rt_capi.h structs) as the GPU core so both see identical
 * inputs. Nothing in the product path may call this.
 *
 * Profiles (rt_profile): 0 = rt_in_one_weekend (CPU), 1 = accelerated-rt-cuda/
 * final.cu, 2 = rt_next_week/cuda and triangles/cuda.
 */

typedef struct { REAL x, y, z; } V3;

static inline V3 FN(v)(REAL x, REAL y, REAL z) { V3 r = {x, y, z}; return r; }
static inline V3 FN(add)(V3 a, V3 b) { return FN(v)(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline V3 FN(sub)(V3 a, V3 b) { return FN(v)(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline V3 FN(mul)(V3 a, V3 b) { return FN(v)(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline V3 FN(scale)(REAL t, V3 a) { return FN(v)(t * a.x, t * a.y, t * a.z); }
/* vec3.h:96 `operator/(v,t)` is (1/t)*v in every tree */
static inline V3 FN(divs)(V3 a, REAL t) { return FN(scale)((REAL)1 / t, a); }
static inline V3 FN(neg)(V3 a) { return FN(v)(-a.x, -a.y, -a.z); }
static inline REAL FN(dot)(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline V3 FN(cross)(V3 u, V3 w) {
  return FN(v)(u.y * w.z - u.z * w.y, u.z * w.x - u.x * w.z, u.x * w.y - u.y * w.x);
}
static inline REAL FN(len2)(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
static inline REAL FN(len)(V3 a) { return SQRT(FN(len2)(a)); }
static inline V3 FN(unit)(V3 a) { return FN(divs)(a, FN(len)(a)); }
static inline V3 FN(from3f)(const float *p) { return FN(v)((REAL)p[0], (REAL)p[1], (REAL)p[2]); }
static inline REAL FN(comp)(V3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

/* ---- RNG: splitmix64-seeded xoshiro256** (the reference uses glibc rand() on
 * the CPU, rtweekend.h:21-24, and cuRAND XORWOW on the GPU; neither sequence is
 * part of the contract — parity of images is statistical). */
typedef struct { uint64_t s[4]; uint64_t draws; int libc; /* 1: glibc rand(); 2: cuRAND XORWOW replay */ uint32_t xv[5], xd; } FN(Rng);
static inline uint64_t FN(rotl)(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static void FN(rng_seed)(FN(Rng) *r, uint64_t seed) {
  uint64_t z = seed;
  for (int i = 0; i < 4; i++) {
    z += 0x9e3779b97f4a7c15ull;
    uint64_t w = z;
    w = (w ^ (w >> 30)) * 0xbf58476d1ce4e5b9ull;
    w = (w ^ (w >> 27)) * 0x94d049bb133111ebull;
    r->s[i] = w ^ (w >> 31);
  }
  r->draws = 0;
  r->libc = 0;
}
/* cuRAND's default generator as the CUDA trees use it: curand_init(seed, 0, 0, &state) and
 * curand_uniform (restated from the published header curand_kernel.h, CUDA 12.9:
 * _curand_init_scratch :772-798 with subsequence = offset = 0, curand :863-874, _curand_uniform =
 * x * 2^-32 + 2^-33, i.e. (0, 1]). Pinned by the known answers the harness records
 * (tests/golden/cuda_ref_*.npz "xorwow"). */
static void FN(rng_seed_xorwow)(FN(Rng) *r, uint64_t seed) {
  uint32_t s0 = ((uint32_t)seed) ^ 0xaad26b49u, s1 = (uint32_t)(seed >> 32) ^ 0xf7dcefddu;
  uint32_t t0 = 1099087573u * s0, t1 = 2591861531u * s1;
  r->xd = 6615241u + t1 + t0;
  r->xv[0] = 123456789u + t0; r->xv[1] = 362436069u ^ t0; r->xv[2] = 521288629u + t1;
  r->xv[3] = 88675123u ^ t1; r->xv[4] = 5783321u + t0;
  r->draws = 0;
  r->libc = 2;
}
static inline float FN(xorwow_uniform)(FN(Rng) *r) {
  uint32_t t = r->xv[0] ^ (r->xv[0] >> 2);
  r->xv[0] = r->xv[1]; r->xv[1] = r->xv[2]; r->xv[2] = r->xv[3]; r->xv[3] = r->xv[4];
  r->xv[4] = (r->xv[4] ^ (r->xv[4] << 4)) ^ (t ^ (t << 1));
  r->xd += 362437u;
  r->draws++;
  return (float)(r->xv[4] + r->xd) * 2.3283064e-10f + (2.3283064e-10f / 2.0f);
}
static inline uint64_t FN(rng_next)(FN(Rng) *r) {
  uint64_t *s = r->s;
  uint64_t result = FN(rotl)(s[1] * 5, 7) * 9, t = s[1] << 17;
  s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = FN(rotl)(s[3], 45);
  r->draws++;
  return result;
}
/* random_double() in [0,1) — rtweekend.h:21-24 (31 random bits there; 24/53 here) */
static inline REAL FN(rnd)(FN(Rng) *r) {
  /* pinning mode: the reference's own generator, rtweekend.h:21-24 */
  if (r->libc == 2) return (REAL)FN(xorwow_uniform)(r);
  if (r->libc) { r->draws++; return (REAL)(rand() / (RAND_MAX + 1.0)); }
#if REAL_IS_FLOAT
  return (REAL)(FN(rng_next)(r) >> 40) * (REAL)(1.0 / 16777216.0);
#else
  return (REAL)(FN(rng_next)(r) >> 11) * (REAL)(1.0 / 9007199254740992.0);
#endif
}
static inline REAL FN(rnd_range)(FN(Rng) *r, REAL lo, REAL hi) { return lo + (hi - lo) * FN(rnd)(r); }

/* vec3.h:103-110 random_in_unit_sphere (CPU): p = vec3::random(-1,1), reject len2 >= 1.
 * CUDA: p = 2*RANDVEC3 - 1, reject len2 >= 1 (accelerated-rt-cuda/material.h:15-21). */
static V3 FN(random_in_unit_sphere)(FN(Rng) *r) {
  for (;;) {
    V3 p;
    if (r->libc == 2) { /* 2.0f * RANDVEC3 - vec3(1,1,1) (rt_next_week/cuda/material.h:11-21); nvcc evaluates left to right */
      p.x = (REAL)2 * FN(rnd)(r) - (REAL)1;
      p.y = (REAL)2 * FN(rnd)(r) - (REAL)1;
      p.z = (REAL)2 * FN(rnd)(r) - (REAL)1;
    } else if (r->libc) { /* g++ evaluates vec3(random_double(..) x3) (vec3.h:46-49) right to left */
      p.z = FN(rnd_range)(r, (REAL)-1, (REAL)1);
      p.y = FN(rnd_range)(r, (REAL)-1, (REAL)1);
      p.x = FN(rnd_range)(r, (REAL)-1, (REAL)1);
    } else {
      p.x = FN(rnd_range)(r, (REAL)-1, (REAL)1);
      p.y = FN(rnd_range)(r, (REAL)-1, (REAL)1);
      p.z = FN(rnd_range)(r, (REAL)-1, (REAL)1);
    }
    if (FN(len2)(p) >= (REAL)1) continue;
    return p;
  }
}
/* vec3.h:112 */
static V3 FN(random_unit_vector)(FN(Rng) *r) { return FN(unit)(FN(random_in_unit_sphere)(r)); }
/* vec3.h:123-130 / accelerated-rt-cuda/camera.h:8-16 */
static V3 FN(random_in_unit_disk)(FN(Rng) *r) {
  for (;;) {
    V3 p;
    if (r->libc) {
      p.y = FN(rnd_range)(r, (REAL)-1, (REAL)1);
      p.x = FN(rnd_range)(r, (REAL)-1, (REAL)1);
    } else {
      p.x = FN(rnd_range)(r, (REAL)-1, (REAL)1);
      p.y = FN(rnd_range)(r, (REAL)-1, (REAL)1);
    }
    p.z = 0;
    if (FN(len2)(p) >= (REAL)1) continue;
    return p;
  }
}
/* vec3.h:114 */
static inline V3 FN(reflect)(V3 a, V3 n) { return FN(sub)(a, FN(scale)((REAL)2 * FN(dot)(a, n), n)); }

#ifndef ORC_INST_DEFINED
#define ORC_INST_DEFINED
typedef struct orc_inst { float sin_y, cos_y, offset[3]; int32_t flag; } orc_inst;
#endif
typedef struct { V3 o, d; REAL tm; } FN(Ray);
static inline V3 FN(at)(const FN(Ray) *r, REAL t) { return FN(add)(r->o, FN(scale)(t, r->d)); }

typedef struct {
  REAL t;
  V3 p, normal;
  REAL u, v;
  int front_face;
  int32_t material;
  int32_t prim; /* RT_PRIM_ID */
} FN(Hit);

typedef struct {
  const rt_scene_desc *sc;
  int profile;
  uint32_t flags;
  /* statistics */
  uint64_t n_prim_tests, n_box_tests;
  /* optional per-primitive translate(rotate_y(object, angle), offset) wrappers in RAY-TRANSFORM form
   * (list order: spheres, triangles, quads), as the reference intersects instances */
  const orc_inst *inst;
  int ref_list_media; /* pinning mode: constant_medium::hit exactly as written, media last in list order */
} FN(World);

/* hittable.h:22-25 (CPU: flips), rt_next_week/cuda/hittable.h:22-30 (never flips),
 * triangles/cuda/include/hittable.h:29 (flips) */
static inline void FN(set_face_normal)(const FN(World) *w, FN(Hit) *h, const FN(Ray) *r, V3 outward) {
  h->front_face = FN(dot)(r->d, outward) < 0;
  int flips = (w->profile == 0) || (w->profile == 2 && (w->flags & RT_FLAG_FLIP_NORMALS));
  if (w->profile == 1) { h->normal = outward; return; } /* final.cu tree has no front_face at all */
  h->normal = (h->front_face || !flips) ? outward : FN(neg)(outward);
}

/* moving_sphere.h:34-36 */
static inline V3 FN(sphere_center)(const rt_sphere *s, REAL tm) {
  V3 c0 = FN(from3f)(s->center0);
  if (!s->moving) return c0;
  V3 c1 = FN(from3f)(s->center1);
  REAL f = (tm - (REAL)s->time0) / ((REAL)s->time1 - (REAL)s->time0);
  return FN(add)(c0, FN(scale)(f, FN(sub)(c1, c0)));
}

/* rt_next_week/cuda/sphere.h:28-40 */
static inline void FN(sphere_uv)(V3 p, REAL *u, REAL *v) {
  const REAL pi = (REAL)3.1415926535897932385;
  REAL theta = ACOS(-p.y);
  REAL phi = ATAN2(-p.z, p.x) + pi;
  *u = phi / ((REAL)2 * pi);
  *v = theta / pi;
}

/* sphere::hit. style A = rt_in_one_weekend/sphere.h:21-55 and
 * rt_next_week/cuda/moving_sphere.h:39-72 (disc<0 rejects, roots accepted on the
 * closed interval); style B = accelerated-rt-cuda/sphere.h:20-49 and
 * rt_next_week/cuda/sphere.h:43-77 (disc>0, open interval). */
static int FN(sphere_hit)(FN(World) *w, int idx, const FN(Ray) *r, REAL t_min, REAL t_max, FN(Hit) *h) {
  const rt_sphere *s = &w->sc->spheres[idx];
  w->n_prim_tests++;
  V3 center = FN(sphere_center)(s, r->tm);
  REAL radius = (REAL)s->radius;
  V3 oc = FN(sub)(r->o, center);
  REAL a = FN(len2)(r->d);
  REAL half_b = FN(dot)(oc, r->d);
  REAL c = FN(len2)(oc) - radius * radius;
  REAL disc = half_b * half_b - a * c;
  REAL root;
  int style_a = (w->profile == 0) || s->moving;
  if (style_a) {
    if (disc < 0) return 0;
    REAL sqrtd = SQRT(disc);
    root = (-half_b - sqrtd) / a;
    if (root < t_min || t_max < root) {
      root = (-half_b + sqrtd) / a;
      if (root < t_min || t_max < root) return 0;
    }
  } else {
    if (!(disc > 0)) return 0;
    root = (-half_b - SQRT(disc)) / a;
    if (!(root < t_max && root > t_min)) {
      root = (-half_b + SQRT(disc)) / a;
      if (!(root < t_max && root > t_min)) return 0;
    }
  }
  h->t = root;
  h->p = FN(at)(r, root);
  V3 outward = FN(divs)(FN(sub)(h->p, center), radius);
  FN(set_face_normal)(w, h, r, outward);
  h->u = h->v = 0;
  if (w->profile == 2) FN(sphere_uv)(outward, &h->u, &h->v);
  h->material = s->material;
  h->prim = RT_PRIM_ID(RT_PRIM_SPHERE, idx);
  return 1;
}

/* triangles/cuda/include/triangle.h:102-215 */
static int FN(triangle_hit)(FN(World) *w, int idx, const FN(Ray) *r, REAL t_min, REAL t_max, FN(Hit) *h) {
  const rt_triangle *tr = &w->sc->triangles[idx];
  w->n_prim_tests++;
  V3 v0 = FN(from3f)(tr->v0), v1 = FN(from3f)(tr->v1), v2 = FN(from3f)(tr->v2);
  V3 fn = FN(from3f)(tr->normal);
  V3 fnu = FN(unit)(fn); /* :49 face_normal_unit */
  REAL nd = FN(dot)(fn, r->d);
  if (FABS(nd) < (REAL)0.01) return 0; /* :126 (double literal in a float compare: promoted) */
  REAL t = FN(dot)(FN(sub)(v0, r->o), fn) / nd;
  if (t < 0) return 0;
  V3 p = FN(at)(r, t);
  if (t < t_min || t > t_max) return 0;
  V3 C = FN(cross)(FN(sub)(v1, v0), FN(sub)(p, v0));
  if (FN(dot)(fnu, C) < 0) return 0;
  C = FN(cross)(FN(sub)(v2, v1), FN(sub)(p, v1));
  REAL u = FN(dot)(fnu, C);
  if (u < 0) return 0;
  C = FN(cross)(FN(sub)(v0, v2), FN(sub)(p, v2));
  REAL vv = FN(dot)(fnu, C);
  if (vv < 0) return 0;
  h->t = t; h->p = p; h->u = u; h->v = vv;
  FN(set_face_normal)(w, h, r, fnu);
  h->material = tr->material;
  h->prim = RT_PRIM_ID(RT_PRIM_TRIANGLE, idx);
  return 1;
}

/* rt_next_week/cuda/aarect.h:38-65 (xy), :93-121 (xz), :149-176 (yz) */
static int FN(quad_hit)(FN(World) *w, int idx, const FN(Ray) *r, REAL t_min, REAL t_max, FN(Hit) *h) {
  const rt_quad *q = &w->sc->quads[idx];
  w->n_prim_tests++;
  int ax = q->axis, ia = (ax == 0) ? 1 : 0, ib = (ax == 2) ? 1 : 2;
  REAL t = ((REAL)q->k - FN(comp)(r->o, ax)) / FN(comp)(r->d, ax);
  if (t < t_min || t > t_max) return 0;
  REAL a = FN(comp)(r->o, ia) + t * FN(comp)(r->d, ia);
  REAL b = FN(comp)(r->o, ib) + t * FN(comp)(r->d, ib);
  if (a < (REAL)q->a0 || a > (REAL)q->a1 || b < (REAL)q->b0 || b > (REAL)q->b1) return 0;
  h->u = (a - (REAL)q->a0) / ((REAL)q->a1 - (REAL)q->a0);
  h->v = (b - (REAL)q->b0) / ((REAL)q->b1 - (REAL)q->b0);
  h->t = t;
  V3 outward = FN(v)(ax == 0, ax == 1, ax == 2);
  FN(set_face_normal)(w, h, r, outward);
  h->material = q->material;
  h->p = FN(at)(r, t);
  h->prim = RT_PRIM_ID(RT_PRIM_QUAD, idx);
  return 1;
}

static int FN(prim_hit_plain)(FN(World) *w, int32_t prim, const FN(Ray) *r, REAL t_min, REAL t_max, FN(Hit) *h) {
  int type = RT_PRIM_TYPE_OF(prim), idx = RT_PRIM_INDEX_OF(prim);
  if (type == RT_PRIM_SPHERE) return FN(sphere_hit)(w, idx, r, t_min, t_max, h);
  if (type == RT_PRIM_TRIANGLE) return FN(triangle_hit)(w, idx, r, t_min, t_max, h);
  return FN(quad_hit)(w, idx, r, t_min, t_max, h);
}
/* translate::hit (rt_next_week/cuda/hittable.h:66-79) around rotate_y::hit (:156-190): the RAY is moved
 * into the object's frame (t is preserved: neither rescales d), the hit point and normal are moved back,
 * and each wrapper re-applies set_face_normal with ITS ray (rotate_y: object-space ray against the
 * world-space normal, then translate: the un-rotated moved ray - which has the last word). */
static int FN(prim_hit)(FN(World) *w, int32_t prim, const FN(Ray) *r, REAL t_min, REAL t_max, FN(Hit) *h) {
  if (!w->inst) return FN(prim_hit_plain)(w, prim, r, t_min, t_max, h);
  int type = RT_PRIM_TYPE_OF(prim), idx = RT_PRIM_INDEX_OF(prim);
  int flat = idx + (type >= RT_PRIM_TRIANGLE ? w->sc->n_spheres : 0) + (type >= RT_PRIM_QUAD ? w->sc->n_triangles : 0);
  const orc_inst *in = &w->inst[flat];
  if (!in->flag) return FN(prim_hit_plain)(w, prim, r, t_min, t_max, h);
  const REAL sn = (REAL)in->sin_y, cs = (REAL)in->cos_y;
  FN(Ray) moved = *r;
  moved.o = FN(sub)(r->o, FN(from3f)(in->offset));
  FN(Ray) rot = moved;
  rot.o.x = cs * moved.o.x - sn * moved.o.z; rot.o.z = sn * moved.o.x + cs * moved.o.z;
  rot.d.x = cs * moved.d.x - sn * moved.d.z; rot.d.z = sn * moved.d.x + cs * moved.d.z;
  if (!FN(prim_hit_plain)(w, prim, &rot, t_min, t_max, h)) return 0;
  V3 p = h->p, n = h->normal;
  p.x = cs * h->p.x + sn * h->p.z; p.z = -sn * h->p.x + cs * h->p.z;
  n.x = cs * h->normal.x + sn * h->normal.z; n.z = -sn * h->normal.x + cs * h->normal.z;
  h->p = p;
  FN(set_face_normal)(w, h, &rot, n);
  h->p = FN(add)(h->p, FN(from3f)(in->offset));
  FN(set_face_normal)(w, h, &moved, h->normal);
  return 1;
}

/* hittable_list::hit — hittable_list.h:20-34 / accelerated-rt-cuda/hittable_list.h:22-37.
 * List order: spheres, triangles, quads. */
static int FN(list_hit)(FN(World) *w, const FN(Ray) *r, REAL t_min, REAL t_max, FN(Hit) *rec) {
  FN(Hit) tmp;
  int any = 0;
  REAL closest = t_max;
  const rt_scene_desc *sc = w->sc;
  for (int i = 0; i < sc->n_spheres; i++)
    if (FN(prim_hit)(w, RT_PRIM_ID(RT_PRIM_SPHERE, i), r, t_min, closest, &tmp)) { any = 1; closest = tmp.t; *rec = tmp; }
  for (int i = 0; i < sc->n_triangles; i++)
    if (FN(prim_hit)(w, RT_PRIM_ID(RT_PRIM_TRIANGLE, i), r, t_min, closest, &tmp)) { any = 1; closest = tmp.t; *rec = tmp; }
  for (int i = 0; i < sc->n_quads; i++)
    if (FN(prim_hit)(w, RT_PRIM_ID(RT_PRIM_QUAD, i), r, t_min, closest, &tmp)) { any = 1; closest = tmp.t; *rec = tmp; }
  return any;
}

/* ---- aabb + the reference's BVH (rt_next_week/cuda/aabb.h, bvh.h) ---- */
typedef struct { V3 mn, mx; } FN(Box);

/* aabb.h:33-50 */
static int FN(box_hit)(const FN(Box) *b, const FN(Ray) *r, REAL t_min, REAL t_max) {
  for (int a = 0; a < 3; a++) {
    REAL invD = (REAL)1 / FN(comp)(r->d, a);
    REAL t0 = (FN(comp)(b->mn, a) - FN(comp)(r->o, a)) * invD;
    REAL t1 = (FN(comp)(b->mx, a) - FN(comp)(r->o, a)) * invD;
    if (invD < 0) { REAL tmp = t1; t1 = t0; t0 = tmp; }
    t_min = t0 > t_min ? t0 : t_min;
    t_max = t1 < t_max ? t1 : t_max;
    if (t_max <= t_min) return 0;
  }
  return 1;
}
/* aabb.h:57-65 */
static FN(Box) FN(surrounding_box)(FN(Box) a, FN(Box) b) {
  FN(Box) r;
  r.mn = FN(v)(FMIN(a.mn.x, b.mn.x), FMIN(a.mn.y, b.mn.y), FMIN(a.mn.z, b.mn.z));
  r.mx = FN(v)(FMAX(a.mx.x, b.mx.x), FMAX(a.mx.y, b.mx.y), FMAX(a.mx.z, b.mx.z));
  return r;
}
/* sphere.h:79-84, moving_sphere.h:74-82, triangle.h:74-100, aarect.h:26-31,81-85,137-142 */
static FN(Box) FN(prim_box)(const rt_scene_desc *sc, int32_t prim, REAL thickness) {
  int type = RT_PRIM_TYPE_OF(prim), idx = RT_PRIM_INDEX_OF(prim);
  FN(Box) b;
  if (type == RT_PRIM_SPHERE) {
    const rt_sphere *s = &sc->spheres[idx];
    REAL rr = (REAL)s->radius;
    V3 rad = FN(v)(rr, rr, rr);
    V3 c0 = FN(sphere_center)(s, (REAL)s->time0);
    b.mn = FN(sub)(c0, rad); b.mx = FN(add)(c0, rad);
    if (s->moving) {
      V3 c1 = FN(sphere_center)(s, (REAL)s->time1);
      FN(Box) b1; b1.mn = FN(sub)(c1, rad); b1.mx = FN(add)(c1, rad);
      b = FN(surrounding_box)(b, b1);
    }
  } else if (type == RT_PRIM_TRIANGLE) {
    const rt_triangle *t = &sc->triangles[idx];
    V3 v0 = FN(from3f)(t->v0), v1 = FN(from3f)(t->v1), v2 = FN(from3f)(t->v2);
    b.mn = FN(v)(FMIN(FMIN(v0.x, v1.x), v2.x), FMIN(FMIN(v0.y, v1.y), v2.y), FMIN(FMIN(v0.z, v1.z), v2.z));
    b.mx = FN(v)(FMAX(FMAX(v0.x, v1.x), v2.x), FMAX(FMAX(v0.y, v1.y), v2.y), FMAX(FMAX(v0.z, v1.z), v2.z));
    REAL *mn = &b.mn.x, *mx = &b.mx.x;
    for (int i = 0; i < 3; i++)
      if (FABS(mn[i] - mx[i]) < thickness) { mn[i] -= thickness; mx[i] += thickness; }
  } else {
    const rt_quad *q = &sc->quads[idx];
    REAL lo[3], hi[3];
    int ax = q->axis, ia = (ax == 0) ? 1 : 0, ib = (ax == 2) ? 1 : 2;
    lo[ax] = (REAL)q->k - thickness; hi[ax] = (REAL)q->k + thickness;
    lo[ia] = (REAL)q->a0; hi[ia] = (REAL)q->a1;
    lo[ib] = (REAL)q->b0; hi[ib] = (REAL)q->b1;
    b.mn = FN(v)(lo[0], lo[1], lo[2]); b.mx = FN(v)(hi[0], hi[1], hi[2]);
  }
  return b;
}

/* bvh_node as the reference builds it: bvh.h:139-196 — random axis, sort by box
 * min, median split, span 1 -> left == right (same leaf twice), span 2 -> ordered
 * pair. Stored in arrays instead of heap objects. child >= 0: inner node index;
 * child < 0: ~RT_PRIM_ID leaf. */
typedef struct { int32_t left, right; FN(Box) box; } FN(RefNode);
typedef struct { FN(RefNode) *nodes; int n, cap; int32_t *prims; const rt_scene_desc *sc; REAL thickness; FN(Rng) rng; } FN(RefBvh);

typedef struct { const FN(RefBvh) *bvh; int axis; } FN(CmpCtx);
/* bvh.h:14-35 box_compare: box_a.min()[axis] < box_b.min()[axis] */
static int FN(box_compare)(const void *pa, const void *pb, void *vctx) {
  const FN(CmpCtx) *ctx = (const FN(CmpCtx) *)vctx;
  int32_t a = *(const int32_t *)pa, b = *(const int32_t *)pb;
  REAL ma = FN(comp)(FN(prim_box)(ctx->bvh->sc, a, ctx->bvh->thickness).mn, ctx->axis);
  REAL mb = FN(comp)(FN(prim_box)(ctx->bvh->sc, b, ctx->bvh->thickness).mn, ctx->axis);
  return (ma < mb) ? -1 : (ma > mb ? 1 : 0);
}
static FN(Box) FN(ref_child_box)(const FN(RefBvh) *b, int32_t child) {
  return child >= 0 ? b->nodes[child].box : FN(prim_box)(b->sc, ~child, b->thickness);
}
static int32_t FN(ref_bvh_build)(FN(RefBvh) *b, int start, int end) {
  int axis = (int)(FN(rnd)(&b->rng) * 3);
  if (axis > 2) axis = 2;
  int span = end - start;
  int me = b->n++;
  FN(CmpCtx) ctx = {b, axis};
  int32_t left, right;
  if (span == 1) {
    left = right = ~b->prims[start];
  } else if (span == 2) {
    if (FN(box_compare)(&b->prims[start], &b->prims[start + 1], &ctx) < 0) { left = ~b->prims[start]; right = ~b->prims[start + 1]; }
    else { left = ~b->prims[start + 1]; right = ~b->prims[start]; }
  } else {
    qsort_r(b->prims + start, span, sizeof(int32_t), FN(box_compare), &ctx);
    int mid = start + span / 2;
    left = FN(ref_bvh_build)(b, start, mid);
    right = FN(ref_bvh_build)(b, mid, end);
  }
  b->nodes[me].left = left; b->nodes[me].right = right;
  b->nodes[me].box = FN(surrounding_box)(FN(ref_child_box)(b, left), FN(ref_child_box)(b, right));
  return me;
}

/* bvh_node::hit — bvh.h:78-137 (iterative; leaf pair tested when EITHER child is a
 * leaf; otherwise both child boxes tested, left first, right pushed). */
static int FN(ref_bvh_hit)(FN(World) *w, const FN(RefBvh) *b, const FN(Ray) *r, REAL t_min, REAL t_max, FN(Hit) *rec) {
  w->n_box_tests++;
  if (!FN(box_hit)(&b->nodes[0].box, r, t_min, t_max)) return 0;
  int32_t stack[64];
  int sp = 0;
  stack[sp++] = -1;
  int32_t node = 0;
  int is_hit = 0;
  do {
    int32_t l = b->nodes[node].left, rr = b->nodes[node].right;
    if (l < 0 || rr < 0) {
      /* the reference assumes both are leaves here (true for its median splits) */
      int hl = (l < 0) ? FN(prim_hit)(w, ~l, r, t_min, t_max, rec) : 0;
      t_max = hl ? rec->t : t_max;
      int hr = (rr < 0) ? FN(prim_hit)(w, ~rr, r, t_min, hl ? rec->t : t_max, rec) : 0;
      t_max = hr ? rec->t : t_max;
      node = stack[--sp];
      if (hl || hr) is_hit = 1;
    } else {
      w->n_box_tests += 2;
      int hl = FN(box_hit)(&b->nodes[l].box, r, t_min, t_max);
      int hr = FN(box_hit)(&b->nodes[rr].box, r, t_min, t_max);
      if (!hl && !hr) node = stack[--sp];
      else {
        node = hl ? l : rr;
        if (hl && hr) stack[sp++] = rr;
      }
    }
  } while (node != -1);
  return is_hit;
}

/* ---- textures & materials ---- */
/* perlin::noise + trilinear_interp — rt_next_week/cuda/perlin.h:29-56,103-122 */
static REAL FN(perlin_noise)(const rt_perlin *pn, V3 p) {
  REAL fx = FLOOR(p.x), fy = FLOOR(p.y), fz = FLOOR(p.z);
  REAL u = p.x - fx, v = p.y - fy, w = p.z - fz;
  REAL uu = u * u * ((REAL)3 - (REAL)2 * u), vv = v * v * ((REAL)3 - (REAL)2 * v), ww = w * w * ((REAL)3 - (REAL)2 * w);
  int i = (int)fx, j = (int)fy, k = (int)fz;
  REAL accum = 0;
  for (int di = 0; di < 2; di++)
    for (int dj = 0; dj < 2; dj++)
      for (int dk = 0; dk < 2; dk++) {
        const float *c = pn->ranvec[pn->perm_x[(i + di) & 255] ^ pn->perm_y[(j + dj) & 255] ^ pn->perm_z[(k + dk) & 255]];
        /* perlin.h:41-43 smooths u, v, w IN PLACE before calling trilinear_interp, so the reference's
         * weight vector (perlin.h:111) is built from the SMOOTHED coordinates (the book uses the raw ones);
         * pinned by tests/golden/cuda_ref_nw_cornell_inst.npz */
        V3 weight_v = FN(v)(uu - (REAL)di, vv - (REAL)dj, ww - (REAL)dk);
        accum += ((REAL)di * uu + (REAL)(1 - di) * ((REAL)1 - uu)) * ((REAL)dj * vv + (REAL)(1 - dj) * ((REAL)1 - vv)) *
                 ((REAL)dk * ww + (REAL)(1 - dk) * ((REAL)1 - ww)) * FN(dot)(FN(from3f)(c), weight_v);
      }
  return accum;
}
/* perlin::turb — perlin.h:58-70 (depth 7) */
static REAL FN(perlin_turb)(const rt_perlin *pn, V3 p) {
  REAL accum = 0, weight = 1;
  for (int i = 0; i < 7; i++) {
    accum += weight * FN(perlin_noise)(pn, p);
    weight *= (REAL)0.5;
    p = FN(scale)((REAL)2, p);
  }
  return FABS(accum);
}
/* texture.h:13-53 (solid, checker), :55-75 (noise), :77-124 (image, with its map shift :110) */
static V3 FN(tex_value)(const rt_scene_desc *sc, const rt_material *m, REAL u, REAL v, V3 p) {
  if (m->texture == RT_TEX_CHECKER) {
    REAL sines = SIN((REAL)10 * p.x) * SIN((REAL)10 * p.y) * SIN((REAL)10 * p.z);
    if (sines < 0) return FN(from3f)(m->albedo2); /* odd */
    return FN(from3f)(m->albedo);                  /* even */
  }
  if (m->texture == RT_TEX_NOISE) {
    REAL sc_ = (REAL)m->albedo2[0];
    const rt_perlin *pn = &sc->perlin[(int)m->albedo2[1]];
    REAL val = (REAL)0.5 * ((REAL)1 + SIN(sc_ * p.z + (REAL)10 * FN(perlin_turb)(pn, FN(scale)(sc_, p))));
    return FN(scale)(val, FN(from3f)(m->albedo));
  }
  if (m->texture == RT_TEX_IMAGE) {
    const rt_image *im = &sc->images[(int)m->albedo2[0]];
    u = u < 0 ? 0 : (u > 1 ? 1 : u);
    v = (REAL)1 - (v < 0 ? 0 : (v > 1 ? 1 : v));
    int i = (int)(u * (REAL)im->width), j = (int)(v * (REAL)im->height);
    if (i >= im->width) i = im->width - 1;
    if (j >= im->height) j = im->height - 1;
    i = (i + im->width / 2 + im->width / 3) % im->width;
    const uint8_t *px = im->rgb + ((size_t)j * im->width + i) * 3;
    const REAL cs = (REAL)1 / (REAL)255;
    return FN(v)(cs * (REAL)px[0], cs * (REAL)px[1], cs * (REAL)px[2]);
  }
  return FN(from3f)(m->albedo);
}
/* material.h:91-96 / accelerated-rt-cuda/material.h:74-78 */
static REAL FN(schlick)(REAL cosine, REAL ref_idx) {
  REAL r0 = ((REAL)1 - ref_idx) / ((REAL)1 + ref_idx);
  r0 = r0 * r0;
  return r0 + ((REAL)1 - r0) * POW((REAL)1 - cosine, (REAL)5);
}
/* vec3.h:116-121 (CPU refract) */
static V3 FN(refract_cpu)(V3 uv, V3 n, REAL eta) {
  REAL cos_theta = FMIN(FN(dot)(FN(neg)(uv), n), (REAL)1);
  V3 perp = FN(scale)(eta, FN(add)(uv, FN(scale)(cos_theta, n)));
  V3 par = FN(scale)(-SQRT(FABS((REAL)1 - FN(len2)(perp))), n);
  return FN(add)(perp, par);
}
/* accelerated-rt-cuda/material.h:80-91 (CUDA refract) */
static int FN(refract_cuda)(V3 vin, V3 n, REAL ni_over_nt, V3 *out) {
  V3 uv = FN(unit)(vin);
  REAL dt = FN(dot)(uv, n);
  REAL disc = (REAL)1 - ni_over_nt * ni_over_nt * ((REAL)1 - dt * dt);
  if (disc > 0) {
    *out = FN(sub)(FN(scale)(ni_over_nt, FN(sub)(uv, FN(scale)(dt, n))), FN(scale)(SQRT(disc), n));
    return 1;
  }
  return 0;
}

/* material::scatter for the three profiles. Returns 1 if scattered. */
static int FN(scatter)(const FN(World) *w, const rt_material *m, const FN(Ray) *rin, const FN(Hit) *rec,
                       V3 *atten, FN(Ray) *out, FN(Rng) *rng) {
  out->tm = rin->tm;
  out->o = rec->p;
  switch (m->type) {
  case RT_MAT_LAMBERTIAN:
    if (w->profile == 0) {
      /* material.h:19-31: normal + random_unit_vector, near_zero guard (vec3.h:53-57,
       * reproduced with its misplaced parenthesis: fabs(e[0] < s)) */
      V3 dir = FN(add)(rec->normal, FN(random_unit_vector)(rng));
      const REAL s = (REAL)1e-8;
      if (FABS((REAL)(dir.x < s)) && (FABS(dir.y) < s) && (FABS(dir.z) < s)) dir = rec->normal;
      out->d = dir;
      *atten = FN(from3f)(m->albedo);
    } else {
      /* accelerated-rt-cuda/material.h:36-44; rt_next_week/cuda/material.h:43-52 */
      V3 target = FN(add)(FN(add)(rec->p, rec->normal), FN(random_in_unit_sphere)(rng));
      out->d = FN(sub)(target, rec->p);
      *atten = (w->profile == 2) ? FN(tex_value)(w->sc, m, rec->u, rec->v, rec->p) : FN(from3f)(m->albedo);
    }
    return 1;
  case RT_MAT_METAL: {
    /* material.h:40-49; accelerated-rt-cuda/material.h:58-71 */
    V3 reflected = FN(reflect)(FN(unit)(rin->d), rec->normal);
    out->d = FN(add)(reflected, FN(scale)((REAL)m->param, FN(random_in_unit_sphere)(rng)));
    *atten = (w->profile == 2) ? FN(tex_value)(w->sc, m, rec->u, rec->v, rec->p) : FN(from3f)(m->albedo);
    return FN(dot)(out->d, rec->normal) > 0;
  }
  case RT_MAT_DIELECTRIC: {
    *atten = FN(v)(1, 1, 1);
    REAL ir = (REAL)m->param;
    if (w->profile == 0) {
      /* material.h:60-85 */
      REAL ratio = rec->front_face ? ((REAL)1 / ir) : ir;
      V3 ud = FN(unit)(rin->d);
      REAL cos_theta = FMIN(FN(dot)(FN(neg)(ud), rec->normal), (REAL)1);
      REAL sin_theta = SQRT((REAL)1 - cos_theta * cos_theta);
      int cannot = ratio * sin_theta > (REAL)1;
      if (cannot || FN(schlick)(cos_theta, ratio) > FN(rnd)(rng)) out->d = FN(reflect)(ud, rec->normal);
      else out->d = FN(refract_cpu)(ud, rec->normal, ratio);
    } else {
      /* accelerated-rt-cuda/material.h:100-132 */
      V3 outward_normal, refracted = FN(v)(0, 0, 0);
      V3 reflected = FN(reflect)(rin->d, rec->normal);
      REAL ni_over_nt, reflect_prob, cosine;
      REAL dn = FN(dot)(rin->d, rec->normal);
      if (dn > 0) {
        outward_normal = FN(neg)(rec->normal);
        ni_over_nt = ir;
        cosine = dn / FN(len)(rin->d);
        cosine = SQRT((REAL)1 - ir * ir * ((REAL)1 - cosine * cosine));
      } else {
        outward_normal = rec->normal;
        ni_over_nt = (REAL)1 / ir;
        cosine = -dn / FN(len)(rin->d);
      }
      if (FN(refract_cuda)(rin->d, outward_normal, ni_over_nt, &refracted)) reflect_prob = FN(schlick)(cosine, ir);
      else reflect_prob = 1;
      if (FN(rnd)(rng) < reflect_prob) out->d = reflected;
      else out->d = refracted;
    }
    return 1;
  }
  case RT_MAT_ISOTROPIC: /* rt_next_week/cuda/material.h:183-191 */
    out->d = FN(random_in_unit_sphere)(rng);
    *atten = FN(tex_value)(w->sc, m, rec->u, rec->v, rec->p);
    return 1;
  default: /* diffuse_light: rt_next_week/cuda/material.h:163-167 */
    return 0;
  }
}
/* rt_next_week/cuda/material.h:34-36,169-172 */
static V3 FN(emitted)(const rt_scene_desc *sc, const rt_material *m, REAL u, REAL v, V3 p) {
  if (m->type == RT_MAT_DIFFUSE_LIGHT) return FN(tex_value)(sc, m, u, v, p);
  return FN(v)(0, 0, 0);
}

static int FN(world_hit)(FN(World) *w, const FN(RefBvh) *bvh, const FN(Ray) *r, REAL t_min, REAL t_max, FN(Hit) *rec) {
  if (bvh) return FN(ref_bvh_hit)(w, bvh, r, t_min, t_max, rec);
  return FN(list_hit)(w, r, t_min, t_max, rec);
}

/* constant_medium::hit — rt_next_week/cuda/constant_medium.h:36-73 over a convex boundary
 * (sphere, or box under rotate_y + translate: box.h, hittable.h). Entry/exit of the whole line,
 * entry clamped to 0 (:54-56), free-flight distance -1/density * log(rnd) (:60-61).
 * Two deliberate deviations, both following the book text the reference tree was written from
 * (RTNW "Volumes"): the scatter event must lie before the closest surface found (the reference
 * ignores t_max, so a medium visited after a nearer surface overrides it, depending on BVH
 * order), and the scattered ray leaves from the scatter point r.at(t) (the reference uses
 * r.at(rec1.t), the boundary entry point, :66). */
static int FN(medium_hit)(const FN(World) *w, int idx, const FN(Ray) *r, REAL t_max, FN(Hit) *h, FN(Rng) *rng) {
  const rt_medium *m = &w->sc->media[idx];
  REAL t1, t2;
  if (m->shape == 0) {
    V3 oc = FN(sub)(r->o, FN(from3f)(m->p0));
    REAL radius = (REAL)m->p1[0];
    REAL a = FN(len2)(r->d), hb = FN(dot)(oc, r->d), c = FN(len2)(oc) - radius * radius;
    REAL disc = hb * hb - a * c;
    if (!(disc > 0)) return 0;
    t1 = (-hb - SQRT(disc)) / a;
    t2 = (-hb + SQRT(disc)) / a;
  } else {
    /* translate::hit then rotate_y::hit move the ray into the box frame (hittable.h) */
    V3 o = FN(sub)(r->o, FN(from3f)(m->offset)), d = r->d;
    REAL sn = (REAL)m->sin_y, cs = (REAL)m->cos_y;
    V3 ol = FN(v)(cs * o.x - sn * o.z, o.y, sn * o.x + cs * o.z);
    V3 dl = FN(v)(cs * d.x - sn * d.z, d.y, sn * d.x + cs * d.z);
    t1 = -INFINITY_R; t2 = INFINITY_R;
    for (int a = 0; a < 3; a++) {
      REAL oa = FN(comp)(ol, a), da = FN(comp)(dl, a), lo = (REAL)m->p0[a], hi = (REAL)m->p1[a];
      if (da == 0) { if (oa < lo || oa > hi) return 0; continue; }
      REAL ta = (lo - oa) / da, tb = (hi - oa) / da;
      if (ta > tb) { REAL s_ = ta; ta = tb; tb = s_; }
      if (ta > t1) t1 = ta;
      if (tb < t2) t2 = tb;
    }
    if (!(t1 < t2)) return 0;
  }
  if (t1 < 0) t1 = 0;
  REAL ray_length = FN(len)(r->d);
  REAL inside = (t2 - t1) * ray_length;
  REAL hit_distance = ((REAL)-1 / (REAL)m->density) * LOG(FN(rnd)(rng));
  if (hit_distance > inside) return 0;
  REAL t = t1 + hit_distance / ray_length;
  if (!(t < t_max)) return 0;
  h->t = t;
  h->p = FN(at)(r, (w->flags & RT_FLAG_REFERENCE_MEDIUM) ? t1 : t); /* constant_medium.h:66 when asked for */
  h->normal = FN(v)(1, 0, 0); /* arbitrary (:69-70) */
  h->front_face = 1;
  h->u = h->v = 0;
  h->material = m->material;
  h->prim = RT_PRIM_ID(RT_PRIM_MEDIUM, idx);
  return 1;
}

/* constant_medium::hit EXACTLY as the reference wrote it (rt_next_week/cuda/constant_medium.h:36-85): two
 * boundary->hit calls (-inf..inf, then rec1.t + 0.00001..inf) through the boundary's own hit function -
 * sphere::hit (sphere.h:43-77) or translate(rotate_y(box)) with box::hit = hittable_list of six rects
 * (box.h:40-63, list order xy(z1), xy(z0), xz(y1), xz(y0), yz(x1), yz(x0)) -, t_min / t_max IGNORED,
 * rec.p = r.at(rec1.t). Used only to pin this file against the reference's device code
 * (tests/test_cuda_ref_pinning.py); the product semantics are medium_hit above. */
static int FN(ref_boundary_hit)(const rt_medium *m, const FN(Ray) *r, REAL t_min, REAL t_max, REAL *t_out) {
  if (m->shape == 0) {
    V3 oc = FN(sub)(r->o, FN(from3f)(m->p0));
    REAL radius = (REAL)m->p1[0];
    REAL a = FN(dot)(r->d, r->d), b = FN(dot)(oc, r->d), c = FN(dot)(oc, oc) - radius * radius;
    REAL disc = b * b - a * c;
    if (disc > 0) {
      REAL temp = (-b - SQRT(disc)) / a;
      if (temp < t_max && temp > t_min) { *t_out = temp; return 1; }
      temp = (-b + SQRT(disc)) / a;
      if (temp < t_max && temp > t_min) { *t_out = temp; return 1; }
    }
    return 0;
  }
  V3 o = FN(sub)(r->o, FN(from3f)(m->offset)), d = r->d;
  REAL sn = (REAL)m->sin_y, cs = (REAL)m->cos_y;
  REAL ol[3] = {cs * o.x - sn * o.z, o.y, sn * o.x + cs * o.z};
  REAL dl[3] = {cs * d.x - sn * d.z, d.y, sn * d.x + cs * d.z};
  int any = 0;
  REAL closest = t_max;
  const int axes[3] = {2, 1, 0};
  for (int f = 0; f < 6; f++) {
    const int ax = axes[f >> 1], ia = (ax == 0) ? 1 : 0, ib = (ax == 2) ? 1 : 2;
    const REAL k = (f & 1) ? (REAL)m->p0[ax] : (REAL)m->p1[ax];
    REAL t = (k - ol[ax]) / dl[ax];
    if (t < t_min || t > closest) continue;
    REAL a = ol[ia] + t * dl[ia], b = ol[ib] + t * dl[ib];
    if (a < (REAL)m->p0[ia] || a > (REAL)m->p1[ia] || b < (REAL)m->p0[ib] || b > (REAL)m->p1[ib]) continue;
    any = 1; closest = t;
  }
  *t_out = closest;
  return any;
}
static int FN(medium_hit_ref)(const FN(World) *w, int idx, const FN(Ray) *r, FN(Hit) *h, FN(Rng) *rng) {
  const rt_medium *m = &w->sc->media[idx];
  REAL t1, t2;
  if (!FN(ref_boundary_hit)(m, r, -INFINITY_R, INFINITY_R, &t1)) return 0;
  if (!FN(ref_boundary_hit)(m, r, (REAL)((double)t1 + 0.00001), INFINITY_R, &t2)) return 0;
  if (t1 < 0) t1 = 0;
  const REAL ray_length = FN(len)(r->d);
  const REAL inside = (t2 - t1) * ray_length;
  const REAL hit_distance = ((REAL)-1 / (REAL)m->density) * LOG(FN(rnd)(rng));
  if (hit_distance > inside) return 0;
  h->t = t1 + hit_distance / ray_length;
  h->p = FN(at)(r, t1);
  h->normal = FN(v)(1, 0, 0);
  h->front_face = 1;
  h->u = h->v = 0;
  h->material = m->material;
  h->prim = RT_PRIM_ID(RT_PRIM_MEDIUM, idx);
  return 1;
}

/* closest of the surfaces and the media (profile 2) */
static int FN(scene_hit)(FN(World) *w, const FN(RefBvh) *bvh, const FN(Ray) *r, REAL t_min, FN(Hit) *rec, FN(Rng) *rng) {
  int any = FN(world_hit)(w, bvh, r, t_min, INFINITY_R, rec);
  REAL closest = any ? rec->t : INFINITY_R;
  FN(Hit) tmp;
  if (w->ref_list_media) { /* pinning mode: the media as the LAST items of the reference's hittable_list */
    for (int i = 0; i < w->sc->n_media; i++)
      if (FN(medium_hit_ref)(w, i, r, &tmp, rng)) { any = 1; closest = tmp.t; *rec = tmp; }
    return any;
  }
  for (int i = 0; i < w->sc->n_media; i++)
    if (FN(medium_hit)(w, i, r, closest, &tmp, rng)) { any = 1; closest = tmp.t; *rec = tmp; }
  return any;
}

/* ray_color main.cpp:57-83 (profile 0, recursion unrolled forward: the product of
 * attenuations is associative up to rounding), get_color final.cu:30-54 (profile 1),
 * emissive get_color rt_next_week/cuda/main.cu:48-105 and
 * triangles/cuda/obj_render.cu:20-86 (profile 2; unwinding emitted_rec /
 * attenuation_rec arrays exactly as the reference does). */
static V3 FN(ray_color)(FN(World) *w, const FN(RefBvh) *bvh, FN(Ray) r, FN(Rng) *rng, uint64_t *nseg) {
  const rt_scene_desc *sc = w->sc;
  const REAL t_min = (REAL)sc->t_min;
  const int depth = sc->max_depth;
  if (w->profile != 2) {
    V3 cur = FN(v)(1, 1, 1);
    V3 att_rec[256];
    for (int i = 0; i < depth && i < 256; i++) {
      FN(Hit) rec;
      (*nseg)++;
      if (FN(world_hit)(w, bvh, &r, t_min, INFINITY_R, &rec)) {
        FN(Ray) sca; V3 att;
        if (FN(scatter)(w, &sc->materials[rec.material], &r, &rec, &att, &sca, rng)) {
          att_rec[i] = att;
          cur = FN(mul)(cur, att); /* final.cu:41 */
          r = sca;
        } else return FN(v)(0, 0, 0);
      } else {
        V3 ud = FN(unit)(r.d);
        REAL t = (REAL)0.5 * (ud.y + (REAL)1);
        V3 c = FN(add)(FN(scale)((REAL)1 - t, FN(v)(1, 1, 1)), FN(scale)(t, FN(v)((REAL)0.5, (REAL)0.7, (REAL)1.0)));
        if (w->profile == 0) { /* main.cpp:76: attenuation * ray_color(...) — the recursion multiplies on the way back */
          while (i-- > 0) c = FN(mul)(att_rec[i], c);
          return c;
        }
        return FN(mul)(cur, c); /* final.cu:49 */
      }
    }
    return FN(v)(0, 0, 0);
  }
  /* profile 2 */
  V3 bg = FN(from3f)(sc->background);
  V3 cur = FN(v)(1, 1, 1);
  V3 emitted_rec[256], atten_rec[256];
  int i;
  for (i = 0; i < depth && i < 256; i++) {
    FN(Hit) rec;
    (*nseg)++;
    if (FN(scene_hit)(w, bvh, &r, t_min, &rec, rng)) {
      const rt_material *m = &sc->materials[rec.material];
      FN(Ray) sca; V3 att;
      V3 em = FN(emitted)(sc, m, rec.u, rec.v, rec.p);
      if (FN(scatter)(w, m, &r, &rec, &att, &sca, rng)) {
        emitted_rec[i] = em; atten_rec[i] = att; r = sca;
      } else {
        cur = FN(mul)(cur, em);
        while (i-- > 0) cur = FN(add)(emitted_rec[i], FN(mul)(cur, atten_rec[i]));
        return cur;
      }
    } else {
      cur = FN(mul)(cur, bg);
      while (i-- > 0) cur = FN(add)(emitted_rec[i], FN(mul)(cur, atten_rec[i]));
      return cur;
    }
  }
  if (sc->flags & RT_FLAG_DEPTH_BACKGROUND) { /* obj_render.cu:78-83 */
    cur = FN(mul)(cur, bg);
    while (i-- > 0) cur = FN(add)(emitted_rec[i], FN(mul)(cur, atten_rec[i]));
    return cur;
  }
  return bg; /* main.cu:104 */
}

/* camera::get_ray camera.h:56-62; rt_next_week/cuda/camera.h:64-72 */
static FN(Ray) FN(get_ray)(const FN(World) *w, REAL s, REAL t, FN(Rng) *rng) {
  const rt_camera *c = &w->sc->camera;
  V3 rd = FN(scale)((REAL)c->lens_radius, FN(random_in_unit_disk)(rng));
  V3 offset = FN(add)(FN(scale)(rd.x, FN(from3f)(c->u)), FN(scale)(rd.y, FN(from3f)(c->v)));
  V3 origin = FN(from3f)(c->origin);
  FN(Ray) r;
  r.o = FN(add)(origin, offset);
  V3 d = FN(add)(FN(from3f)(c->lower_left_corner), FN(scale)(s, FN(from3f)(c->horizontal)));
  d = FN(add)(d, FN(scale)(t, FN(from3f)(c->vertical)));
  d = FN(sub)(d, origin);
  r.d = FN(sub)(d, offset);
  r.tm = 0;
  if (w->profile == 2) r.tm = FN(rnd)(rng) * ((REAL)c->time1 - (REAL)c->time0) + (REAL)c->time0;
  return r;
}

/* ------------------------------------------------------------------ exports */

/* Closest hit. mode 0: brute-force list (hittable_list order). mode 1: through a
 * reference-style BVH built here (bvh.h:139-196). rays [n][8] floats = o, tm, d, pad
 * (same layout as rt_trace_closest). */
int FN(closest_hit)(const rt_scene_desc *sc, int profile, const float *rays, int n, double t_min, double t_max,
                    int mode, uint64_t bvh_seed, int32_t *prim_id, double *t_out, uint64_t *counters /*[2] or NULL*/) {
  FN(World) w = {sc, profile, sc->flags, 0, 0};
  FN(RefBvh) bvh, *pb = NULL;
  int nprim = sc->n_spheres + sc->n_triangles + sc->n_quads;
  if (mode == 1 && nprim > 0) {
    bvh.nodes = (FN(RefNode) *)malloc(sizeof(FN(RefNode)) * (size_t)(2 * nprim + 1));
    bvh.prims = (int32_t *)malloc(sizeof(int32_t) * (size_t)nprim);
    bvh.n = 0; bvh.sc = sc;
    bvh.thickness = (sc->flags & RT_FLAG_FLIP_NORMALS) ? (REAL)0.01 : (REAL)0.1; /* rtweekend.h:60 of each tree */
    FN(rng_seed)(&bvh.rng, bvh_seed);
    int k = 0;
    for (int i = 0; i < sc->n_spheres; i++) bvh.prims[k++] = RT_PRIM_ID(RT_PRIM_SPHERE, i);
    for (int i = 0; i < sc->n_triangles; i++) bvh.prims[k++] = RT_PRIM_ID(RT_PRIM_TRIANGLE, i);
    for (int i = 0; i < sc->n_quads; i++) bvh.prims[k++] = RT_PRIM_ID(RT_PRIM_QUAD, i);
    FN(ref_bvh_build)(&bvh, 0, nprim);
    pb = &bvh;
  }
  for (int k = 0; k < n; k++) {
    const float *q = rays + 8 * k;
    FN(Ray) r;
    r.o = FN(from3f)(q); r.tm = (REAL)q[3]; r.d = FN(from3f)(q + 4);
    FN(Hit) rec;
    REAL tmax = (t_max >= 3.0e38) ? INFINITY_R : (REAL)t_max;
    if (FN(world_hit)(&w, pb, &r, (REAL)t_min, tmax, &rec)) { prim_id[k] = rec.prim; t_out[k] = (double)rec.t; }
    else { prim_id[k] = -1; t_out[k] = 0.0; }
  }
  if (counters) { counters[0] = w.n_box_tests; counters[1] = w.n_prim_tests; }
  if (pb) { free(bvh.nodes); free(bvh.prims); }
  return 0;
}

/* Closest hit through a PACKED threaded BVH downloaded from the GPU core
 * (rt_bvh_node, include/rt_capi.h), plus the always-tested big-primitive list,
 * with the reference's own primitive tests. Counts box and primitive tests —
 * the N_box / N_prim figures of the roofline model. */
int FN(closest_hit_packed)(const rt_scene_desc *sc, int profile, const rt_bvh_node *nodes, int n_nodes,
                           const int32_t *leaf_prims, const int32_t *big, int n_big, const float *rays, int n, double t_min, double t_max,
                           int32_t *prim_id, double *t_out, uint64_t *counters /*[2]*/) {
  FN(World) w = {sc, profile, sc->flags, 0, 0};
  for (int k = 0; k < n; k++) {
    const float *q = rays + 8 * k;
    FN(Ray) r;
    r.o = FN(from3f)(q); r.tm = (REAL)q[3]; r.d = FN(from3f)(q + 4);
    FN(Hit) rec, tmp;
    int any = 0;
    REAL closest = (t_max >= 3.0e38) ? INFINITY_R : (REAL)t_max;
    for (int i = 0; i < n_big; i++)
      if (FN(prim_hit)(&w, big[i], &r, (REAL)t_min, closest, &tmp)) { any = 1; closest = tmp.t; rec = tmp; }
    int node = 0;
    while (node < n_nodes) {
      const rt_bvh_node *nd = &nodes[node];
      FN(Box) b;
      b.mn = FN(from3f)(nd->bmin); b.mx = FN(from3f)(nd->bmax);
      w.n_box_tests++;
      /* conservative closed-interval slab test (a packed box may be flat) */
      REAL lo = (REAL)t_min, hi = closest;
      int ok = 1;
      for (int a = 0; a < 3 && ok; a++) {
        REAL invD = (REAL)1 / FN(comp)(r.d, a);
        REAL t0 = (FN(comp)(b.mn, a) - FN(comp)(r.o, a)) * invD;
        REAL t1 = (FN(comp)(b.mx, a) - FN(comp)(r.o, a)) * invD;
        if (invD < 0) { REAL s = t0; t0 = t1; t1 = s; }
        if (t0 > lo) lo = t0;
        if (t1 < hi) hi = t1;
        if (hi < lo) ok = 0;
      }
      if (ok) {
        if (nd->payload >= 0) { node = nd->payload; continue; }
        int enc = ~nd->payload, first = enc >> 3, count = (enc & 7) + 1;
        for (int q2 = 0; q2 < count; q2++)
          if (FN(prim_hit)(&w, leaf_prims[first + q2], &r, (REAL)t_min, closest, &tmp)) { any = 1; closest = tmp.t; rec = tmp; }
      }
      node = nd->escape;
    }
    if (any) { prim_id[k] = rec.prim; t_out[k] = (double)rec.t; } else { prim_id[k] = -1; t_out[k] = 0.0; }
  }
  if (counters) { counters[0] = w.n_box_tests; counters[1] = w.n_prim_tests; }
  return 0;
}

/* Render rows [j0,j1): the pixel/sample loop of worker() main.cpp:267-290 /
 * render final.cu:75-96 / main.cu:127-149. Per-pixel RNG streams (seed, pixel)
 * so that the result does not depend on how rows are split over callers.
 * sum/sumsq: [H*W*3] doubles, index j*W+i, j=0 bottom. Returns segments. */
uint64_t FN(render)(const rt_scene_desc *sc, int profile, int W, int H, int spp, uint64_t seed, int j0, int j1,
                    int use_ref_bvh, int libc_rand, double *sum, double *sumsq) {
  FN(World) w = {sc, profile, sc->flags, 0, 0};
  FN(RefBvh) bvh, *pb = NULL;
  int nprim = sc->n_spheres + sc->n_triangles + sc->n_quads;
  if (use_ref_bvh && nprim > 0) {
    bvh.nodes = (FN(RefNode) *)malloc(sizeof(FN(RefNode)) * (size_t)(2 * nprim + 1));
    bvh.prims = (int32_t *)malloc(sizeof(int32_t) * (size_t)nprim);
    bvh.n = 0; bvh.sc = sc;
    bvh.thickness = (sc->flags & RT_FLAG_FLIP_NORMALS) ? (REAL)0.01 : (REAL)0.1;
    FN(rng_seed)(&bvh.rng, 1984);
    int k = 0;
    for (int i = 0; i < sc->n_spheres; i++) bvh.prims[k++] = RT_PRIM_ID(RT_PRIM_SPHERE, i);
    for (int i = 0; i < sc->n_triangles; i++) bvh.prims[k++] = RT_PRIM_ID(RT_PRIM_TRIANGLE, i);
    for (int i = 0; i < sc->n_quads; i++) bvh.prims[k++] = RT_PRIM_ID(RT_PRIM_QUAD, i);
    FN(ref_bvh_build)(&bvh, 0, nprim);
    pb = &bvh;
  }
  uint64_t nseg = 0;
  for (int j = j0; j < j1; j++) {
    for (int i = 0; i < W; i++) {
      FN(Rng) rng;
      FN(rng_seed)(&rng, seed * 0x9E3779B97F4A7C15ull + (uint64_t)((size_t)j * W + i) * 0xD1B54A32D192ED03ull + 1);
      rng.libc = libc_rand; /* pinning mode: caller did srand(); single thread only */
      double s[3] = {0, 0, 0}, s2[3] = {0, 0, 0};
      for (int k = 0; k < spp; k++) {
        REAL u, v;
        if (profile == 0) { /* main.cpp:278-279 */
          u = ((REAL)i + FN(rnd)(&rng)) / (REAL)(W - 1);
          v = ((REAL)j + FN(rnd)(&rng)) / (REAL)(H - 1);
        } else { /* final.cu:85-86 */
          u = ((REAL)i + FN(rnd)(&rng)) / (REAL)W;
          v = ((REAL)j + FN(rnd)(&rng)) / (REAL)H;
        }
        FN(Ray) r = FN(get_ray)(&w, u, v, &rng);
        V3 c = FN(ray_color)(&w, pb, r, &rng, &nseg);
        double cc[3] = {(double)c.x, (double)c.y, (double)c.z};
        for (int a = 0; a < 3; a++) { s[a] += cc[a]; s2[a] += cc[a] * cc[a]; }
      }
      size_t idx = (size_t)j * W + i;
      for (int a = 0; a < 3; a++) { sum[3 * idx + a] = s[a]; if (sumsq) sumsq[3 * idx + a] = s2[a]; }
    }
  }
  if (pb) { free(bvh.nodes); free(bvh.prims); }
  return nseg;
}

/* One scatter event through this restatement (for BSDF distribution tests).
 * rng_seed selects the stream; returns 1 if scattered. */
int FN(scatter_one)(const rt_scene_desc *sc, int profile, int material, const double *d_in, const double *outward_n,
                    uint64_t rng_seed_, int libc_rand, double *dir_out, double *atten_out) {
  FN(World) w = {sc, profile, sc->flags, 0, 0};
  FN(Rng) rng;
  FN(rng_seed)(&rng, rng_seed_);
  rng.libc = libc_rand;
  FN(Ray) rin, out;
  rin.d = FN(v)((REAL)d_in[0], (REAL)d_in[1], (REAL)d_in[2]);
  rin.o = FN(neg)(rin.d);
  rin.tm = 0;
  FN(Hit) rec;
  rec.p = FN(v)(0, 0, 0); rec.t = 1; rec.u = rec.v = 0; rec.material = material; rec.prim = 0;
  FN(set_face_normal)(&w, &rec, &rin, FN(v)((REAL)outward_n[0], (REAL)outward_n[1], (REAL)outward_n[2]));
  V3 att = FN(v)(0, 0, 0);
  out.d = FN(v)(0, 0, 0);
  int ok = FN(scatter)(&w, &sc->materials[material], &rin, &rec, &att, &out, &rng);
  dir_out[0] = out.d.x; dir_out[1] = out.d.y; dir_out[2] = out.d.z;
  atten_out[0] = att.x; atten_out[1] = att.y; atten_out[2] = att.z;
  return ok;
}

/* Evidence for the medium deviation documented above medium_hit: constant_medium::hit AS THE REFERENCE
 * WROTE IT (rt_next_week/cuda/constant_medium.h:36-73: no t_max test) inside hittable_list::hit's loop
 * (hittable_list.h:20-34 pattern: every object is offered closest_so_far, and whatever returns true
 * replaces the record). Returns, over n_trials random numbers, the fraction of trials in which the
 * MEDIUM ends up as the closest hit for one ray, with the media visited before (order 0) or after
 * (order 1) the surfaces - the two differ, i.e. the reference's image depends on list / BVH order. */
double FN(medium_reference_order_demo)(const rt_scene_desc *sc, const float *ray8, int order, int n_trials, uint64_t seed) {
  FN(World) w = {sc, 2, sc->flags, 0, 0};
  FN(Rng) rng;
  FN(rng_seed)(&rng, seed);
  rng.libc = 0;
  FN(Ray) r;
  r.o = FN(from3f)(ray8); r.tm = (REAL)ray8[3]; r.d = FN(from3f)(ray8 + 4);
  int medium_wins = 0;
  for (int k = 0; k < n_trials; k++) {
    FN(Hit) rec, tmp;
    int any = 0, is_medium = 0;
    REAL closest = INFINITY_R;
    for (int pass = 0; pass < 2; pass++) {
      if ((pass == 0) == (order == 0)) { /* the media */
        for (int i = 0; i < sc->n_media; i++) {
          /* the reference's version: the scatter event is accepted wherever it falls inside the boundary */
          if (FN(medium_hit)(&w, i, &r, INFINITY_R, &tmp, &rng)) { any = 1; closest = tmp.t; rec = tmp; is_medium = 1; }
        }
      } else { /* the surfaces, each offered closest_so_far */
        if (FN(list_hit)(&w, &r, (REAL)sc->t_min, closest, &tmp)) { any = 1; closest = tmp.t; rec = tmp; is_medium = 0; }
      }
    }
    (void)rec;
    medium_wins += any && is_medium;
  }
  return (double)medium_wins / (double)n_trials;
}

/* texture value of a material at (u, v, p) through this restatement (texture parity tests) */
void FN(tex_value_at)(const rt_scene_desc *sc, int material, int n, const double *uvp /*[n][5]*/, double *rgb /*[n][3]*/) {
  for (int k = 0; k < n; k++) {
    const double *q = uvp + 5 * k;
    V3 c = FN(tex_value)(sc, &sc->materials[material], (REAL)q[0], (REAL)q[1], FN(v)((REAL)q[2], (REAL)q[3], (REAL)q[4]));
    rgb[3 * k] = c.x; rgb[3 * k + 1] = c.y; rgb[3 * k + 2] = c.z;
  }
}

/* sphere (u, v) of an outward unit normal — rt_next_week/cuda/sphere.h:28-40 */
void FN(sphere_uv_at)(int n, const double *normal /*[n][3]*/, double *uv /*[n][2]*/) {
  for (int k = 0; k < n; k++) {
    REAL u, v;
    FN(sphere_uv)(FN(v)((REAL)normal[3 * k], (REAL)normal[3 * k + 1], (REAL)normal[3 * k + 2]), &u, &v);
    uv[2 * k] = u; uv[2 * k + 1] = v;
  }
}

/* write_color color.h:14-28 (profile 0) / final.cu:91-95,227-229 (profiles 1, 2;
 * values above 255 are clamped here, the reference lets emitters overflow). */
void FN(quantise)(int profile, const double *sum3, int spp, int *rgb) {
  for (int a = 0; a < 3; a++) {
    if (profile == 0) {
      REAL scale = (REAL)1.0 / (REAL)spp;
      REAL x = SQRT(scale * (REAL)sum3[a]);
      x = x < 0 ? 0 : (x > (REAL)0.999 ? (REAL)0.999 : x);
      rgb[a] = (int)((REAL)256 * x);
    } else {
      REAL x = (REAL)sum3[a] / (REAL)spp;
      x = SQRT(x);
      int q = (int)(255.99 * (double)x);
      rgb[a] = q > 255 ? 255 : q;
    }
  }
}

/* ------------------------------------------------------------------ pinning against the reference's CUDA trees
 * The records below mirror oracle/ref_cuda_harness.cu (HHit, HScatter, HMediumHit), which runs the reference's
 * unmodified device code; tests/test_cuda_ref_pinning.py compares field by field. `obj` = flat list index
 * (spheres, triangles, quads, then media). Random numbers: XORWOW, seed + 1 + ray index per ray. */
#ifndef ORC_PIN_RECORDS_DEFINED
#define ORC_PIN_RECORDS_DEFINED
typedef struct { int32_t hit, obj; float t, p[3], n[3], u, v; int32_t front_face; } orc_pin_hit;
typedef struct { int32_t ok; float att[3], o[3], d[3], tm, emitted[3]; int32_t draws; } orc_pin_scatter;
typedef struct { int32_t hit; float t, p[3]; int32_t draws; } orc_pin_medium;
#endif
static int FN(flat_index)(const rt_scene_desc *sc, int32_t prim) {
  int type = RT_PRIM_TYPE_OF(prim), idx = RT_PRIM_INDEX_OF(prim);
  if (type == RT_PRIM_SPHERE) return idx;
  if (type == RT_PRIM_TRIANGLE) return sc->n_spheres + idx;
  if (type == RT_PRIM_QUAD) return sc->n_spheres + sc->n_triangles + idx;
  return sc->n_spheres + sc->n_triangles + sc->n_quads + idx;
}
static FN(Ray) FN(ray_from8)(const float *q) {
  FN(Ray) r;
  r.o = FN(from3f)(q); r.tm = (REAL)q[3]; r.d = FN(from3f)(q + 4);
  return r;
}
void FN(pin_hits)(const rt_scene_desc *sc, int profile, const orc_inst *inst, const float *rays, int n, double t_min,
                  double t_max, orc_pin_hit *out) {
  FN(World) w = {sc, profile, sc->flags, 0, 0, inst, 0};
  for (int k = 0; k < n; k++) {
    FN(Ray) r = FN(ray_from8)(rays + 8 * k);
    FN(Hit) rec;
    orc_pin_hit o;
    memset(&o, 0, sizeof o);
    o.obj = -1;
    if (FN(list_hit)(&w, &r, (REAL)t_min, (t_max >= 3.0e38) ? INFINITY_R : (REAL)t_max, &rec)) {
      o.hit = 1; o.obj = FN(flat_index)(sc, rec.prim); o.t = (float)rec.t;
      o.p[0] = (float)rec.p.x; o.p[1] = (float)rec.p.y; o.p[2] = (float)rec.p.z;
      o.n[0] = (float)rec.normal.x; o.n[1] = (float)rec.normal.y; o.n[2] = (float)rec.normal.z;
      o.u = (float)rec.u; o.v = (float)rec.v; o.front_face = rec.front_face;
    }
    out[k] = o;
  }
}
void FN(pin_scatter)(const rt_scene_desc *sc, int profile, const orc_inst *inst, const float *rays, int n, double t_min,
                     double t_max, uint64_t seed, orc_pin_scatter *out) {
  FN(World) w = {sc, profile, sc->flags, 0, 0, inst, 0};
  for (int k = 0; k < n; k++) {
    FN(Ray) r = FN(ray_from8)(rays + 8 * k);
    FN(Rng) rng;
    FN(rng_seed_xorwow)(&rng, seed + 1 + (uint64_t)k);
    FN(Hit) rec;
    orc_pin_scatter o;
    memset(&o, 0, sizeof o);
    o.ok = -1;
    if (FN(list_hit)(&w, &r, (REAL)t_min, (t_max >= 3.0e38) ? INFINITY_R : (REAL)t_max, &rec)) {
      const rt_material *m = &sc->materials[rec.material];
      FN(Ray) sca;
      sca.o = FN(v)(0, 0, 0); sca.d = FN(v)(0, 0, 0); sca.tm = 0;
      V3 att = FN(v)(0, 0, 0);
      V3 em = FN(emitted)(sc, m, rec.u, rec.v, rec.p);
      o.ok = FN(scatter)(&w, m, &r, &rec, &att, &sca, &rng);
      if (m->type == RT_MAT_DIFFUSE_LIGHT) { sca.o = FN(v)(0, 0, 0); sca.tm = 0; } /* the reference leaves `scattered` untouched */
      o.att[0] = (float)att.x; o.att[1] = (float)att.y; o.att[2] = (float)att.z;
      o.o[0] = (float)sca.o.x; o.o[1] = (float)sca.o.y; o.o[2] = (float)sca.o.z;
      o.d[0] = (float)sca.d.x; o.d[1] = (float)sca.d.y; o.d[2] = (float)sca.d.z;
      o.tm = (float)sca.tm;
      o.emitted[0] = (float)em.x; o.emitted[1] = (float)em.y; o.emitted[2] = (float)em.z;
    }
    o.draws = (int32_t)rng.draws;
    out[k] = o;
  }
}
void FN(pin_media)(const rt_scene_desc *sc, const float *rays, int n, uint64_t seed, orc_pin_medium *out) {
  FN(World) w = {sc, 2, sc->flags, 0, 0, NULL, 1};
  for (int k = 0; k < n; k++) {
    FN(Ray) r = FN(ray_from8)(rays + 8 * k);
    for (int m = 0; m < sc->n_media; m++) {
      FN(Rng) rng;
      FN(rng_seed_xorwow)(&rng, seed + 1 + (uint64_t)k);
      FN(Hit) rec;
      orc_pin_medium o;
      memset(&o, 0, sizeof o);
      if (FN(medium_hit_ref)(&w, m, &r, &rec, &rng)) {
        o.hit = 1; o.t = (float)rec.t;
        o.p[0] = (float)rec.p.x; o.p[1] = (float)rec.p.y; o.p[2] = (float)rec.p.z;
      }
      o.draws = (int32_t)rng.draws;
      out[(size_t)k * sc->n_media + m] = o;
    }
  }
}
/* get_color (main.cu:48-105 / obj_render.cu:20-86) per ray over the LIST world with the media last, each ray
 * from its own XORWOW stream. out: [n][4] = rgb, draws. */
void FN(pin_color)(const rt_scene_desc *sc, int profile, const orc_inst *inst, const float *rays, int n, uint64_t seed,
                   float *out) {
  FN(World) w = {sc, profile, sc->flags, 0, 0, inst, 1};
  for (int k = 0; k < n; k++) {
    FN(Ray) r = FN(ray_from8)(rays + 8 * k);
    FN(Rng) rng;
    FN(rng_seed_xorwow)(&rng, seed + 1 + (uint64_t)k);
    uint64_t nseg = 0;
    V3 c = FN(ray_color)(&w, NULL, r, &rng, &nseg);
    out[4 * k] = (float)c.x; out[4 * k + 1] = (float)c.y; out[4 * k + 2] = (float)c.z; out[4 * k + 3] = (float)rng.draws;
  }
}
void FN(xorwow_kat)(uint64_t seed, int n, float *out) {
  FN(Rng) rng;
  FN(rng_seed_xorwow)(&rng, seed);
  for (int k = 0; k < n; k++) out[k] = FN(xorwow_uniform)(&rng);
}
