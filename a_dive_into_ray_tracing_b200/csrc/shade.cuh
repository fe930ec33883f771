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
    s = ((float)i + x1) / (float)(W - 1);
    t = ((float)j + x2) / (float)(H - 1);
  } else { // final.cu:85-86
    s = ((float)i + x1) / (float)W;
    t = ((float)j + x2) / (float)H;
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
  float r0 = (1.0f - ref_idx) / (1.0f + ref_idx);
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

RT_HD V3f material_color(float4 m0, float4 m1, V3f p) {
  int tex = (RT_F2I(m0.w) >> 8) & 0xff;
  if (tex == RT_TEX_CHECKER) {
    float sines = rt_fast_sin(10.0f * p.x) * rt_fast_sin(10.0f * p.y) * rt_fast_sin(10.0f * p.z);
    if (sines < 0.0f) return xyz(m1); // odd
  }
  return xyz(m0);
}

// Surface interaction at an accepted hit. Updates the ray (origin = hit point,
// new direction), the throughput `beta` and (profile 2) the radiance `L`.
// Returns true when the path continues.
//   rnd: the four random words of this bounce.
template <int PROFILE, bool GENERAL>
RT_HD bool shade_hit(const DevScene &S, const ShadeParams &sp, Ray &r, const HitAcc &h, V3f &beta, V3f &L,
                     Philox4 rnd) {
  const V3f p = madd(r.o, h.t, r.d);
  V3f outward;
  int mat;
  int type = GENERAL ? RT_PRIM_TYPE_OF(h.id) : RT_PRIM_SPHERE;
  int idx = GENERAL ? RT_PRIM_INDEX_OF(h.id) : h.id;
  if (type == RT_PRIM_SPHERE) {
    float4 s = S.sph[idx];
    V3f c = xyz(s);
    if (GENERAL && S.any_moving) {
      float4 mv = S.sph_mv[idx];
      if (mv.w != 0.0f) c = sphere_center_at(s, mv, S.sph_t0[idx], r.tm);
    }
    outward = (1.0f / s.w) * (p - c); // (p - center) / radius
    mat = S.sph_mat[idx];
  } else if (type == RT_PRIM_TRIANGLE) {
    outward = xyz(S.tri_n[idx]);
    mat = S.tri_mat[idx];
  } else {
    int ax = RT_F2I(S.quad[2 * idx].w);
    outward = v3(ax == 0 ? 1.f : 0.f, ax == 1 ? 1.f : 0.f, ax == 2 ? 1.f : 0.f);
    mat = S.quad_mat[idx];
  }
  const float dn_out = dot(r.d, outward);
  const bool front_face = dn_out < 0.0f;
  V3f n = outward;
  if (PROFILE == 0 || (PROFILE == 2 && (sp.flags & RT_FLAG_FLIP_NORMALS)))
    if (!front_face) n = -outward; // hittable.h:22-25 / triangles include/hittable.h:29

  const float4 m0 = S.mats[2 * mat], m1 = S.mats[2 * mat + 1];
  const int mtype = RT_F2I(m0.w) & 0xff;
  const float u1 = u01(rnd.x), u2 = u01(rnd.y), u3 = u01(rnd.z);

  if (PROFILE == 2 && mtype == RT_MAT_DIFFUSE_LIGHT) {
    L = L + beta * material_color(m0, m1, p); // emitted; never scatters
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
      v = exp2f(__log2f(u3) * (1.0f / 3.0f)) * v; // cbrt(u3): radius of a uniform point in the ball
#else
      v = cbrtf(u3) * v;
#endif
    }
    att = (PROFILE == 2) ? material_color(m0, m1, p) : xyz(m0);
    if (mtype == RT_MAT_LAMBERTIAN) {
      dir = n + v;
      if (PROFILE == 0 && dot(dir, dir) < 1e-16f) dir = n; // degenerate direction guard (material.h:24-26)
    } else {
      V3f refl = reflect(normalize(r.d), n);
      dir = madd(refl, m1.w, v);
      if (!(dot(dir, n) > 0.0f)) return false; // absorbed (adds nothing in any profile)
    }
  } else { // dielectric
    att = v3(1.f, 1.f, 1.f);
    const float ir = m1.w;
    if (PROFILE == 0) {
      float ratio = front_face ? (1.0f / ir) : ir;
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
      V3f reflected = reflect(r.d, n);
      float dn = dot(r.d, n);
      float inv_len = RT_RSQRT(dot(r.d, r.d));
      V3f on;
      float ni_over_nt, cosine;
      if (dn > 0.0f) {
        on = -n; ni_over_nt = ir;
        cosine = dn * inv_len;
        cosine = RT_SQRT(RT_FMAX(0.0f, 1.0f - ir * ir * (1.0f - cosine * cosine)));
      } else {
        on = n; ni_over_nt = 1.0f / ir;
        cosine = -dn * inv_len;
      }
      V3f uv = inv_len * r.d;
      float dt = dot(uv, on);
      float disc = 1.0f - ni_over_nt * ni_over_nt * (1.0f - dt * dt);
      float reflect_prob = 1.0f;
      V3f refracted = reflected;
      if (disc > 0.0f) {
        refracted = ni_over_nt * madd(uv, -dt, on) - RT_SQRT(disc) * on;
        reflect_prob = schlick5(cosine, ir);
      }
      dir = (u1 < reflect_prob) ? reflected : refracted;
    }
  }
  beta = beta * att;
  r.o = p;
  r.d = dir;
  return true;
}

// Radiance picked up by a ray that leaves the scene.
RT_HD V3f miss_radiance(const ShadeParams &sp, V3f d) {
  return sp.sky_gradient ? sky_color(d) : sp.background;
}
