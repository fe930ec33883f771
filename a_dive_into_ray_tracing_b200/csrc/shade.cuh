// shade.cuh — camera rays, surface interaction and the three BSDFs, per semantic
// profile (rt_profile). Replaces the reference's virtual `material::scatter` /
// `emitted`, `camera::get_ray` and the miss shader:
//   camera::get_ray      rt_in_one_weekend/camera.h:56-62; rt_next_week/cuda/camera.h:64-72
//   lambertian           material.h:19-31 (P0: normal + random_unit_vector);
//                        accelerated-rt-cuda/material.h:36-44, rt_next_week/cuda/material.h:43-52
//                        (P1/P2: normal + random_in_unit_sphere, un-normalised)
//   metal                material.h:40-49; accelerated-rt-cuda/material.h:58-71
//   dielectric           material.h:60-96 + vec3.h:116-121 (P0);
//                        accelerated-rt-cuda/material.h:74-132 (P1/P2)
//   diffuse_light        rt_next_week/cuda/material.h:157-176
//   checker_texture      rt_next_week/cuda/texture.h:33-53
//   sky                  main.cpp:80-82, final.cu:46-49
//   noise_texture/perlin rt_next_week/cuda/texture.h:55-75, perlin.h:29-72,103-122
//   image_texture        rt_next_week/cuda/texture.h:77-124; sphere uv sphere.h:28-40, rect uv aarect.h:52-53
//   isotropic            rt_next_week/cuda/material.h:178-195
//   constant_medium      rt_next_week/cuda/constant_medium.h:36-73 (apply_media below)
// Sampling is rejection-free with identical distributions (philox.cuh).
#pragma once
#include "intersect.cuh"
#include "philox.cuh"

struct ShadeParams {
  V3f background;
  int sky_gradient;
  uint32_t flags;
  float t_min;
  int max_depth;
};

template <int PROFILE>
RT_HD Ray gen_camera_ray(const DevCamera &cam, int W, int H, int i, int j, float x1, float x2, float x3, float x4,
                         float x5) {
  float s, t;
  if (PROFILE == 0) { // main.cpp:278-279
    s = RT_FDIV((float)i + x1, (float)(W - 1));
    t = RT_FDIV((float)j + x2, (float)(H - 1));
  } else { // final.cu:85-86
    s = RT_FDIV((float)i + x1, (float)W);
    t = RT_FDIV((float)j + x2, (float)H);
  }
  V3f offset = v3(0.f, 0.f, 0.f);
  if (cam.lens_radius > 0.0f) {
    float dx, dy;
    sample_unit_disk(x3, x4, &dx, &dy);
    dx *= cam.lens_radius; dy *= cam.lens_radius;
    offset = madd(dx * cam.u, dy, cam.v);
  }
  Ray r;
  r.o = cam.origin + offset;
  V3f d = madd(madd(cam.llc, s, cam.horizontal), t, cam.vertical);
  r.d = (d - cam.origin) - offset;
  r.tm = (PROFILE == 2) ? RT_FMA(x5, cam.time1 - cam.time0, cam.time0) : 0.0f;
  return r;
}

RT_HD V3f sky_color(V3f d) {
  float uy = d.y * RT_RSQRT(dot(d, d));
  float t = 0.5f * (uy + 1.0f);
  float w = 1.0f - t;
  return v3(RT_FMA(t, 0.5f, w), RT_FMA(t, 0.7f, w), RT_FMA(t, 1.0f, w));
}

RT_HD V3f reflect(V3f v, V3f n) { return madd(v, -2.0f * dot(v, n), n); }

RT_HD float schlick5(float cosine, float ref_idx) {
  float r0 = RT_FDIV(1.0f - ref_idx, 1.0f + ref_idx);
  r0 = r0 * r0;
  float m = 1.0f - cosine, m2 = m * m;
  return RT_FMA(1.0f - r0, m2 * m2 * m, r0);
}

RT_HD float rt_fast_sin(float x) {
#ifdef __CUDA_ARCH__
  return __sinf(x);
#else
  return sinf(x);
#endif
}

// ---- rarely used texture kinds live in out-of-line functions (RT_COLD, rt_common.cuh) so that the render
// kernel's register allocation is not shaped by them.

// perlin::noise with the trilinear Hermite interpolation of perlin.h:29-56,103-122 (as the reference
// computes it: see the note at the weight vector)
RT_COLD float perlin_noise(const float4 *__restrict__ vec, const uint8_t *__restrict__ perm, V3f p) {
  const float fx = floorf(p.x), fy = floorf(p.y), fz = floorf(p.z);
  const float u = p.x - fx, v = p.y - fy, w = p.z - fz;
  const float uu = u * u * (3.0f - 2.0f * u), vv = v * v * (3.0f - 2.0f * v), ww = w * w * (3.0f - 2.0f * w);
  const int i = (int)fx, j = (int)fy, k = (int)fz;
  float accum = 0.0f;
#pragma unroll
  for (int c = 0; c < 8; c++) {
    const int di = c >> 2, dj = (c >> 1) & 1, dk = c & 1;
    const int h = perm[(i + di) & 255] ^ perm[256 + ((j + dj) & 255)] ^ perm[512 + ((k + dk) & 255)];
    const float4 g = vec[h];
    // perlin.h:41-43 smooths u, v, w in place BEFORE trilinear_interp, so the reference's weight vector
    // (perlin.h:111) uses the smoothed coordinates (pinned: tests/golden/cuda_ref_nw_cornell_inst.npz)
    const float d = RT_FMA(g.z, ww - (float)dk, RT_FMA(g.y, vv - (float)dj, g.x * (uu - (float)di)));
    accum += (di ? uu : 1.0f - uu) * (dj ? vv : 1.0f - vv) * (dk ? ww : 1.0f - ww) * d;
  }
  return accum;
}

// Textures that need more than the hit point's checker parity: noise (perlin::turb, depth 7,
// perlin.h:58-70) and image lookups with the primitive's (u, v).
// (Everything is passed and returned BY VALUE: handing the kernel's DevScene / hit record to an
// out-of-line function by reference would pin them in local memory for the whole kernel.)
// p = the hit point the reference hands to texture::value (world space: translate / rotate_y move rec.p back,
// hittable.h:73,181); p_uv / outward = hit point and outward normal in the PRIMITIVE's space, where sphere::hit /
// xy_rect::hit compute (u, v) (the same as p / the world normal for primitives that are not instanced).
RT_COLD V3f texture_ext(const float4 *__restrict__ perlin_vec, const uint8_t *__restrict__ perlin_perm,
                        const DevImage *__restrict__ images, const float4 *__restrict__ quads, float4 m0, float4 m1,
                        V3f p, V3f p_uv, V3f outward, int32_t prim) {
  const int tex = (RT_F2I(m0.w) >> 8) & 0xff;
  if (tex == RT_TEX_NOISE) {
    const int table = (int)m1.y;
    const float4 *vec = perlin_vec + 256 * table;
    const uint8_t *perm = perlin_perm + 768 * table;
    const float scale = m1.x;
    V3f q = scale * p;
    float accum = 0.0f, weight = 1.0f;
    for (int i = 0; i < 7; i++) {
      accum = RT_FMA(weight, perlin_noise(vec, perm, q), accum);
      weight *= 0.5f;
      q = 2.0f * q;
    }
    const float val = 0.5f * (1.0f + sinf(RT_FMA(scale, p.z, 10.0f * fabsf(accum))));
    return val * xyz(m0);
  }
  // image: (u, v) of the primitive
  float u = 0.0f, v = 0.0f;
  const int type = RT_PRIM_TYPE_OF(prim), idx = RT_PRIM_INDEX_OF(prim);
  if (type == RT_PRIM_SPHERE) { // sphere.h:28-40
    const float pi = 3.1415926535897932385f;
    const float theta = acosf(RT_FMIN(1.0f, RT_FMAX(-1.0f, -outward.y)));
    const float phi = atan2f(-outward.z, outward.x) + pi;
    u = phi / (2.0f * pi);
    v = theta / pi;
  } else if (type == RT_PRIM_QUAD) { // aarect.h:52-53
    const float4 q0 = quads[2 * idx], q1 = quads[2 * idx + 1];
    const int ax = RT_F2I(q0.w);
    const float a = ax == 0 ? p_uv.y : p_uv.x, b = ax == 2 ? p_uv.y : p_uv.z;
    u = (a - q0.y) / (q0.z - q0.y);
    v = (b - q1.x) / (q1.y - q1.x);
  }
  const DevImage im = images[(int)m1.x];
  u = RT_FMIN(RT_FMAX(u, 0.0f), 1.0f);
  v = 1.0f - RT_FMIN(RT_FMAX(v, 0.0f), 1.0f);
  int i = (int)(u * (float)im.width), j = (int)(v * (float)im.height);
  if (i >= im.width) i = im.width - 1;
  if (j >= im.height) j = im.height - 1;
  i = (i + im.width / 2 + im.width / 3) % im.width; // texture.h:110 ("try to shift the map")
  const uint8_t *px = im.rgb + ((size_t)j * im.width + i) * 3;
  const float cs = 1.0f / 255.0f;
  return v3(cs * (float)px[0], cs * (float)px[1], cs * (float)px[2]);
}

// EXT: the scene uses noise / image textures or media (a separate kernel instantiation, so that
// scenes without them keep the leaner code).
template <bool EXT>
RT_HD V3f material_color(const DevScene &S, float4 m0, float4 m1, V3f p, V3f p_uv, V3f outward, int32_t prim) {
  int tex = (RT_F2I(m0.w) >> 8) & 0xff;
  if (tex == RT_TEX_CHECKER) {
    float sines = rt_fast_sin(10.0f * p.x) * rt_fast_sin(10.0f * p.y) * rt_fast_sin(10.0f * p.z);
    if (sines < 0.0f) return xyz(m1); // odd
  } else if (EXT && tex >= RT_TEX_NOISE) {
    return texture_ext(S.perlin_vec, S.perlin_perm, S.images, S.quad, m0, m1, p, p_uv, outward, prim);
  }
  return xyz(m0);
}

// constant_medium::hit for every medium of the scene, at the START of a segment: the
// free-flight distance is sampled up front and the scatter event, if it falls inside the
// boundary, becomes the segment's initial closest hit — surfaces found nearer by the traversal
// replace it, which is exactly "closest of surfaces and media" and also shortens the
// traversal. Entry/exit over the whole line, entry clamped to 0 (constant_medium.h:42-56);
// distance = -1/density * log(u) (:60-61). Deviations from the reference: the event
// must lie before the closest surface (the reference ignores t_max) and scattering continues
// from the scatter point (the reference restarts at the boundary entry point, :66).
// Random numbers: one word per medium derived from the Philox block of the event that produced the ray (medium_word).
// [t1, t2] = the whole line's intersection with medium m's convex boundary (constant_medium.h:42-52), unclamped.
// (Hardware division / square root: free-path sampling is checked statistically; the IEEE sequences made this the
// largest piece of straight-line code of every segment start of a foggy scene.)
RT_COLD bool medium_interval(const float4 *__restrict__ media, int m, V3f ro, V3f rd, float &t1, float &t2) {
  const float4 a0 = media[4 * m], a1 = media[4 * m + 1];
  if (RT_F2I(a0.w) == 0) {
    const V3f oc = ro - xyz(a0);
    const float a = dot(rd, rd), hb = dot(oc, rd), c = RT_FMA(-a1.x, a1.x, dot(oc, oc));
    const float disc = RT_FMA(hb, hb, -a * c);
    if (!(disc > 0.0f)) return false;
    const float sq = RT_SQRT(disc), ia = RT_FDIV(1.0f, a);
    t1 = (-hb - sq) * ia;
    t2 = (-hb + sq) * ia;
    return true;
  }
  const float4 a2 = media[4 * m + 2], a3 = media[4 * m + 3];
  const V3f o = ro - xyz(a2);
  const float sn = a3.x, cs = a3.y;
  const float ol[3] = {cs * o.x - sn * o.z, o.y, sn * o.x + cs * o.z};
  const float dl[3] = {cs * rd.x - sn * rd.z, rd.y, sn * rd.x + cs * rd.z};
  const float lo[3] = {a0.x, a0.y, a0.z}, hi[3] = {a1.x, a1.y, a1.z};
  t1 = -INFINITY; t2 = INFINITY;
  bool miss = false;
  for (int k = 0; k < 3; k++) {
    if (dl[k] == 0.0f) { miss = miss || ol[k] < lo[k] || ol[k] > hi[k]; continue; }
    const float id = RT_FDIV(1.0f, dl[k]);
    const float ta = (lo[k] - ol[k]) * id, tb = (hi[k] - ol[k]) * id;
    t1 = RT_FMAX(t1, RT_FMIN(ta, tb));
    t2 = RT_FMIN(t2, RT_FMAX(ta, tb));
  }
  return !miss && t1 < t2;
}

// One 32-bit word per medium from ONE word of the Philox block that produced this ray (the camera event's or the
// bounce's fourth word): two rounds of a multiply-xorshift finaliser over word + m * golden ratio. (A Philox call of
// its own per segment start was 100 straight-line instructions for one or two words.)
RT_HD uint32_t medium_word(uint32_t seed, int m) {
  uint32_t x = seed + 0x9E3779B9u * (uint32_t)(m + 1);
  x ^= x >> 16; x *= 0x85EBCA6Bu;
  x ^= x >> 13; x *= 0xC2B2AE35u;
  x ^= x >> 16;
  return x;
}

RT_COLD HitAcc apply_media(const float4 *__restrict__ media, int n_media, V3f ro, V3f rd, uint32_t seed, HitAcc h) {
  const float len = RT_SQRT(dot(rd, rd));
  for (int m = 0; m < n_media; m++) {
    float t1, t2;
    if (!medium_interval(media, m, ro, rd, t1, t2)) continue;
    t1 = RT_FMAX(t1, 0.0f);
    // u in (0, 1]: log(0) cannot occur
    const float u = (float)((medium_word(seed, m) >> 8) + 1u) * (1.0f / 16777216.0f);
#ifdef __CUDA_ARCH__
    const float hit_distance = media[4 * m + 1].w * __logf(u);
#else
    const float hit_distance = media[4 * m + 1].w * logf(u);
#endif
    if (hit_distance > (t2 - t1) * len) continue;
    const float t = t1 + RT_FDIV(hit_distance, len);
    if (t < h.t) { h.t = t; h.id = RT_PRIM_ID(RT_PRIM_MEDIUM, m); }
  }
  return h;
}

// Outward (nominal) normal and material of the primitive hit at p = r.at(h.t):
// sphere (p - c)/r (sphere.h:57), triangle unit face normal (triangle.h:49), rect +axis
// (aarect.h:54), medium arbitrary (constant_medium.h:69).
template <bool GENERAL, bool EXT, bool SPH = false>
RT_HD void surface_at(const DevScene &S, const Ray &r, const HitAcc &h, V3f p, V3f &outward, int &mat) {
  int type = (GENERAL && !SPH) ? RT_PRIM_TYPE_OF(h.id) : RT_PRIM_SPHERE;
  int idx = (GENERAL && !SPH) ? RT_PRIM_INDEX_OF(h.id) : h.id;
  if (type == RT_PRIM_SPHERE) {
    float4 s = S.sph[idx];
    V3f c = xyz(s);
    if (GENERAL && S.any_moving) {
      float4 mv = S.sph_mv[idx];
      if (mv.w != 0.0f) c = sphere_center_at(s, mv, S.sph_t0[idx], r.tm);
    }
    outward = RT_FDIV(1.0f, s.w) * (p - c); // (p - center) / radius
    mat = S.sph_mat[idx];
  } else if (type == RT_PRIM_TRIANGLE) {
    outward = xyz(S.tri_n[idx]);
    mat = S.tri_mat[idx];
  } else if (!EXT || type == RT_PRIM_QUAD) {
    int ax = RT_F2I(S.quad[2 * idx].w);
    outward = v3(ax == 0 ? 1.f : 0.f, ax == 1 ? 1.f : 0.f, ax == 2 ? 1.f : 0.f);
    mat = S.quad_mat[idx];
  } else { // scatter event inside a constant_medium: normal is arbitrary (constant_medium.h:69)
    outward = v3(1.f, 0.f, 0.f);
    mat = RT_F2I(S.media[4 * idx + 2].w);
  }
}

// The same for a hit inside instance `inst` of a two-level scene (inst < 0: a world-level primitive): the
// primitive lives in object space, so the hit point and the ray go there, and the normal comes back with the
// rotation (rotate_y::hit, hittable.h:181-187). p_uv / n_uv = object-space hit point and normal (texture (u, v)).
template <bool GENERAL, bool EXT, bool INST, bool SPH = false>
RT_HD void surface_at_inst(const DevScene &S, const Ray &r, const HitAcc &h, V3f p, int inst, V3f &outward, int &mat,
                           V3f &p_uv, V3f &n_uv) {
  if (INST && inst >= 0) {
    const float4 *rec = S.inst + RT_INST_STRIDE * inst;
    Ray ro = r;
    ro.o = inst_point_to_object(rec, r.o);
    ro.d = inst_vector_to_object(rec, r.d);
    p_uv = madd(ro.o, h.t, ro.d);
    surface_at<GENERAL, EXT>(S, ro, h, p_uv, n_uv, mat);
    outward = inst_vector_to_world(rec, n_uv);
  } else {
    surface_at<GENERAL, EXT, SPH>(S, r, h, p, outward, mat);
    p_uv = p;
    n_uv = outward;
  }
}

// First-hit features of a camera ray for denoisers (rt_render_aov): albedo = the texture colour the
// integrator would multiply by (1 for glass, emission clamped to 1 for lights), normal = the
// shading normal (facing the ray where the profile flips it), t = the ray parameter.
template <int PROFILE, bool GENERAL, bool INST = false>
RT_HD void first_hit_features(const DevScene &S, const ShadeParams &sp, const Ray &r, const HitAcc &h, V3f &albedo,
                              V3f &normal, int inst = -1) {
  const V3f p = madd(r.o, h.t, r.d);
  V3f outward, p_uv, n_uv;
  int mat;
  surface_at_inst<GENERAL, GENERAL, INST>(S, r, h, p, inst, outward, mat, p_uv, n_uv);
  normal = outward;
  if (PROFILE == 0 || (PROFILE == 2 && (sp.flags & RT_FLAG_FLIP_NORMALS)))
    if (!(dot(r.d, outward) < 0.0f)) normal = -outward;
  const float4 m0 = S.mats[2 * mat], m1 = S.mats[2 * mat + 1];
  const int mtype = RT_F2I(m0.w) & 0xff;
  if (mtype == RT_MAT_DIELECTRIC) albedo = v3(1.f, 1.f, 1.f);
  else albedo = (PROFILE == 2) ? material_color<GENERAL>(S, m0, m1, p, p_uv, n_uv, h.id) : xyz(m0);
  if (mtype == RT_MAT_DIFFUSE_LIGHT) albedo = v3(RT_FMIN(albedo.x, 1.f), RT_FMIN(albedo.y, 1.f), RT_FMIN(albedo.z, 1.f));
}

// dielectric::scatter of the CUDA trees (accelerated-rt-cuda/material.h:78-132): reflect / refract by Schlick's
// approximation with the un-normalised incoming direction.
RT_HD V3f dielectric_dir(V3f d, V3f n, float ir, float u1) {
  V3f reflected = reflect(d, n);
  float dn = dot(d, n);
  float inv_len = RT_RSQRT(dot(d, d));
  V3f on;
  float ni_over_nt, cosine;
  if (dn > 0.0f) {
    on = -n; ni_over_nt = ir;
    cosine = dn * inv_len;
    cosine = RT_SQRT(RT_FMAX(0.0f, 1.0f - ir * ir * (1.0f - cosine * cosine)));
  } else {
    on = n; ni_over_nt = RT_FDIV(1.0f, ir);
    cosine = -dn * inv_len;
  }
  V3f uv = inv_len * d;
  float dt = dot(uv, on);
  float disc = 1.0f - ni_over_nt * ni_over_nt * (1.0f - dt * dt);
  float reflect_prob = 1.0f;
  V3f refracted = reflected;
  if (disc > 0.0f) {
    refracted = ni_over_nt * madd(uv, -dt, on) - RT_SQRT(disc) * on;
    reflect_prob = schlick5(cosine, ir);
  }
  return (u1 < reflect_prob) ? reflected : refracted;
}
// (kept inline: an out-of-line copy for the extended kernel measured no gain, 97.9 vs 96.6 ms on the final scene)

// Surface interaction at an accepted hit. Updates the ray (origin = hit point,
// new direction), the throughput `beta` and (profile 2) the radiance `L`.
// Returns true when the path continues.
//   rnd: the four random words of this bounce.
template <int PROFILE, bool GENERAL, bool EXT = false, bool INST = false, bool SPH = false>
RT_HD bool shade_hit(const DevScene &S, const ShadeParams &sp, Ray &r, const HitAcc &h, V3f &beta, V3f &L,
                     Philox4 rnd, int inst = -1) {
  const V3f p = madd(r.o, h.t, r.d);
  V3f outward, p_uv, n_uv;
  int mat;
  surface_at_inst<GENERAL, EXT, INST, SPH>(S, r, h, p, inst, outward, mat, p_uv, n_uv);
  const float dn_out = dot(r.d, outward);
  const bool front_face = dn_out < 0.0f;
  V3f n = outward;
  if (PROFILE == 0 || (PROFILE == 2 && (sp.flags & RT_FLAG_FLIP_NORMALS)))
    if (!front_face) n = -outward; // hittable.h:22-25 / triangles include/hittable.h:29

  const float4 m0 = S.mats[2 * mat], m1 = S.mats[2 * mat + 1];
  const int mtype = RT_F2I(m0.w) & 0xff;
  // u3 scales a unit vector into the ball (radius cbrt(u3)): kept off 0 so that an isotropic scatter never
  // yields the zero direction (u01 returns [0, 1): 0 with probability 2^-24 per event)
  const float u1 = u01(rnd.x), u2 = u01(rnd.y), u3 = fmaxf(u01(rnd.z), 5.9604645e-8f);

  if (PROFILE == 2 && mtype == RT_MAT_DIFFUSE_LIGHT) {
    L = L + beta * material_color<EXT>(S, m0, m1, p, p_uv, n_uv, h.id); // emitted; never scatters
    return false;
  }
  V3f dir, att;
  if (mtype != RT_MAT_DIELECTRIC) {
    // lambertian and metal (94 % of the hits) share the random direction: one uniform unit
    // vector, scaled into the unit ball where the reference samples the ball — computed once,
    // outside the per-material branches, so the warp runs the transcendental code a single time
    V3f v = sample_unit_vector(u1, u2);
    if (!(PROFILE == 0 && mtype == RT_MAT_LAMBERTIAN)) {
#ifdef __CUDA_ARCH__
      {
        float rad; // cbrt(u3) = 2^(log2(u3)/3): radius of a uniform point in the ball (u3 = 0 -> 0)
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(rad) : "f"(__log2f(u3) * (1.0f / 3.0f)));
        v = rad * v;
      }
#else
      v = cbrtf(u3) * v;
#endif
    }
    att = (PROFILE == 2) ? material_color<EXT>(S, m0, m1, p, p_uv, n_uv, h.id) : xyz(m0);
    if (mtype == RT_MAT_LAMBERTIAN) {
      dir = n + v;
      if (PROFILE == 0 && dot(dir, dir) < 1e-16f) dir = n; // degenerate direction guard (material.h:24-26)
    } else if (EXT && mtype == RT_MAT_ISOTROPIC) {
      dir = v; // random_in_unit_sphere (material.h:187-188)
    } else {
      V3f refl = reflect(normalize(r.d), n);
      dir = madd(refl, m1.w, v);
      if (!(dot(dir, n) > 0.0f)) return false; // absorbed (adds nothing in any profile)
    }
  } else { // dielectric
    att = v3(1.f, 1.f, 1.f);
    const float ir = m1.w;
    if (PROFILE == 0) {
      float ratio = front_face ? RT_FDIV(1.0f, ir) : ir;
      V3f ud = normalize(r.d);
      float cos_theta = RT_FMIN(-dot(ud, n), 1.0f);
      float sin_theta = RT_SQRT(RT_FMAX(0.0f, RT_FMA(-cos_theta, cos_theta, 1.0f)));
      bool cannot = ratio * sin_theta > 1.0f;
      if (cannot || schlick5(cos_theta, ratio) > u1) {
        dir = reflect(ud, n);
      } else { // refract, vec3.h:116-121
        V3f perp = ratio * madd(ud, cos_theta, n);
        float par = -RT_SQRT(fabsf(1.0f - dot(perp, perp)));
        dir = madd(perp, par, n);
      }
    } else {
      dir = dielectric_dir(r.d, n, ir, u1);
    }
  }
  beta = beta * att;
  V3f o_next = p;
  if (EXT && (sp.flags & RT_FLAG_REFERENCE_MEDIUM) && RT_PRIM_TYPE_OF(h.id) == RT_PRIM_MEDIUM) {
    // the reference restarts the scattered ray at the boundary ENTRY point: rec.p = r.at(rec1.t), entry clamped
    // to the ray's origin (constant_medium.h:56-58,74)
    float t1 = 0.f, t2 = 0.f;
    medium_interval(S.media, RT_PRIM_INDEX_OF(h.id), r.o, r.d, t1, t2);
    o_next = madd(r.o, RT_FMAX(t1, 0.0f), r.d);
  }
  r.o = o_next;
  r.d = dir;
  return true;
}

// Radiance picked up by a ray that leaves the scene.
RT_HD V3f miss_radiance(const ShadeParams &sp, V3f d) {
  return sp.sky_gradient ? sky_color(d) : sp.background;
}
