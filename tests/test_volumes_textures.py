"""rt_next_week scenes 3-8 (SURVEY.md §8f "next" row): noise_texture / perlin, image_texture,
isotropic + constant_medium. CPU part: the oracle restatement against analytic results, and the
device code (host emulation, tests/emu) against the oracle. GPU part: the kernels through the C ABI
against the same oracle."""
import numpy as np
import pytest

from tests import stats_util as SU
from a_dive_into_ray_tracing_b200 import ctypes_defs as D
from a_dive_into_ray_tracing_b200 import scenes
from a_dive_into_ray_tracing_b200.scenes import _mat, _medium_box, _medium_sphere, _quad, _sphere
from tests.emu.pyemu import Emu

NEW_SCENES = ["two_perlin_spheres", "earth", "simple_light", "cornell_smoke", "next_week_final"]


def transmittance_scene(shape, density, W=16, H=16):
    """White background seen through a purely absorbing medium (black isotropic albedo): the
    expected pixel value is exp(-density * chord)."""
    mats = np.array([_mat(D.RT_MAT_ISOTROPIC, (0, 0, 0))], D.MATERIAL_DT)
    if shape == "sphere":
        media = [_medium_sphere((0, 0, 0), 1.0, density, 0)]
    else:  # unit-ish box turned by 30 degrees: the chord through its centre along z
        media = [_medium_box((-0.5, -0.5, -0.5), (0.5, 0.5, 0.5), 30.0, (0, 0, 0), density, 0)]
    sc = D.Scene(materials=mats, media=np.array(media, D.MEDIUM_DT), profile=D.RT_PROFILE_NEXT_WEEK,
                 background=(1, 1, 1), sky_gradient=0, name="transmittance_" + shape)
    sc.camera = D.camera_from_lookat((0, 0, 40), (0, 0, 0), (0, 1, 0), 0.2, 1.0, 0.0, 40.0, dtype=np.float32)
    return sc


def expected_transmittance(shape, density):
    if shape == "sphere":
        return np.exp(-density * 2.0)
    return np.exp(-density * 1.0 / np.cos(np.deg2rad(30.0)))  # slab of thickness 1 crossed at 30 degrees


@pytest.mark.parametrize("shape", ["sphere", "box"])
def test_medium_transmittance_oracle_and_device_code(l1_32, l1_64, shape):
    density, W, H, spp = 0.6, 16, 16, 256
    sc = transmittance_scene(shape, density, W, H)
    want = expected_transmittance(shape, density)
    se = np.sqrt(want * (1 - want) / (W * H * spp))
    for orc in (l1_32, l1_64):
        r, _, _ = orc.render_parallel(sc, 2, W, H, spp, seed=3)
        assert abs(r.mean() / spp - want) < 4 * se + 2e-3, (shape, r.mean() / spp, want)
    s, _, st = Emu(sc).render(W, H, spp, seed=11)
    assert abs(s.mean() / spp - want) < 4 * se + 2e-3, (shape, s.mean() / spp, want)
    # an absorbed path goes on with zero throughput (the reference does the same): > 1 segment/path
    assert st[1] > st[0]


def test_single_scatter_albedo_energy(l1_32):
    """A dense white medium ball under a uniform white sky is a white furnace: every path ends in
    the sky with throughput 1, so the image is exactly 1 (any energy loss/gain shows here)."""
    mats = np.array([_mat(D.RT_MAT_ISOTROPIC, (1, 1, 1))], D.MATERIAL_DT)
    sc = D.Scene(materials=mats, media=np.array([_medium_sphere((0, 0, 0), 1.0, 3.0, 0)], D.MEDIUM_DT),
                 profile=D.RT_PROFILE_NEXT_WEEK, background=(1, 1, 1), sky_gradient=0, max_depth=500)
    sc.camera = D.camera_from_lookat((0, 0, 5), (0, 0, 0), (0, 1, 0), 30.0, 1.0, 0.0, 5.0, dtype=np.float32)
    r, _, _ = l1_32.render_parallel(sc, 2, 12, 12, 32, seed=1)
    np.testing.assert_allclose(r / 32, 1.0, atol=1e-5)
    s, _, _ = Emu(sc).render(12, 12, 32)
    np.testing.assert_allclose(s / 32, 1.0, atol=1e-5)


def test_perlin_tables_and_noise_properties(l1_64):
    t = scenes.make_perlin(5)
    for name in ("perm_x", "perm_y", "perm_z"):
        assert sorted(t[name].tolist()) == list(range(256))
    assert np.abs(t["ranvec"]).max() <= 1.0
    sc = scenes.two_perlin_spheres(8, 8)
    rng = np.random.default_rng(0)
    p = rng.uniform(-20, 20, (4000, 3))
    uvp = np.concatenate([np.zeros((len(p), 2)), p], 1)
    v = l1_64.tex_value(sc, 0, uvp)
    assert np.all(v >= 0) and np.all(v <= 1) and np.allclose(v[:, 0], v[:, 1]) and 0.2 < v.mean() < 0.8
    # continuity: the noise is smooth across lattice cell borders
    eps = 1e-6
    q = np.round(p) + rng.uniform(-0.3, 0.3, p.shape) * np.array([0, 1, 1])  # x on a lattice plane
    a = l1_64.tex_value(sc, 0, np.concatenate([np.zeros((len(q), 2)), q - [eps, 0, 0]], 1))
    b = l1_64.tex_value(sc, 0, np.concatenate([np.zeros((len(q), 2)), q + [eps, 0, 0]], 1))
    assert np.abs(a - b).max() < 1e-3
    # lattice period 256 / scale in every axis (perm index & 255)
    c = l1_64.tex_value(sc, 0, np.concatenate([np.zeros((len(p), 2)), p + [64.0, 0, 0]], 1))
    np.testing.assert_allclose(c, v, atol=1e-9)


def test_device_textures_match_oracle(l1_32, l1_64):
    """noise and image lookups of the device code (host emulation) against the restatement"""
    rng = np.random.default_rng(1)
    # noise: random points in the scene's range
    sc = scenes.two_perlin_spheres(8, 8)
    e = Emu(sc)
    p = rng.uniform(-8, 8, (5000, 3)).astype(np.float32)
    got = e.texture(0, D.prim_id(D.RT_PRIM_SPHERE, 1), p, np.tile([0, 1, 0], (len(p), 1)))
    want = l1_64.tex_value(sc, 0, np.concatenate([np.zeros((len(p), 2)), p.astype(np.float64)], 1))
    # sin(scale z + 10 turb): fp32 turbulence error ~1e-6 * 10 -> a few 1e-5 after the sine
    assert np.abs(got - want).max() < 2e-4, np.abs(got - want).max()
    assert np.abs(got - want).mean() < 1e-5
    # image on a sphere: same texel except within rounding of a texel border
    sc = scenes.earth(8, 8)
    e = Emu(sc)
    n = rng.normal(size=(5000, 3))
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    got = e.texture(0, D.prim_id(D.RT_PRIM_SPHERE, 0), (2 * n).astype(np.float32), n.astype(np.float32))
    uv = l1_64.sphere_uv(n.astype(np.float32).astype(np.float64))
    want = l1_64.tex_value(sc, 0, np.concatenate([uv, 2 * n], 1))
    same = np.all(np.abs(got - want) < 1e-6, axis=1)
    assert same.mean() > 0.995, same.mean()
    # image on a rect: (u, v) from the in-plane coordinates (aarect.h:52-53)
    mats = np.array([_mat(D.RT_MAT_DIFFUSE_LIGHT, (1, 1, 1), texture=D.RT_TEX_IMAGE)], D.MATERIAL_DT)
    scq = D.Scene(quads=np.array([_quad(2, -1, 3, 2, 4, 0.5, 0)], D.QUAD_DT), materials=mats,
                  images=[scenes.procedural_earth(64, 32)], profile=D.RT_PROFILE_NEXT_WEEK)
    e = Emu(scq)
    a, b = rng.uniform(-1, 3, 3000), rng.uniform(2, 4, 3000)
    pts = np.stack([a, b, np.full_like(a, 0.5)], 1).astype(np.float32)
    got = e.texture(0, D.prim_id(D.RT_PRIM_QUAD, 0), pts, np.tile([0, 0, 1], (len(a), 1)))
    uvp = np.concatenate([((pts[:, 0] + 1) / 4)[:, None], ((pts[:, 1] - 2) / 2)[:, None], pts], 1)
    want = l1_64.tex_value(scq, 0, uvp)
    assert np.all(np.abs(got - want) < 1e-6, axis=1).mean() > 0.995


@pytest.mark.parametrize("name", NEW_SCENES)
def test_emulated_next_week_scenes_vs_l1(l1_32, name):
    W, H = (40, 40) if name in ("cornell_smoke", "next_week_final") else (48, 32)
    spp = 128
    sc = getattr(scenes, name)(W, H)
    s, s2, st = Emu(sc, quality=1).render(W, H, spp, seed=7)
    r, r2, nseg = l1_32.render_parallel(sc, 2, W, H, spp, seed=5)
    mu_a, var_a = SU.mean_var(s, s2, spp)
    mu_b, var_b = SU.mean_var(r, r2, spp)
    ok, d, b = SU.three_sigma_check(mu_a, var_a, spp, mu_b, var_b, spp)
    assert ok, (name, d, b)
    z = SU.zscores(mu_a, var_a, spp, mu_b, var_b, spp)
    assert abs(z.mean()) < 0.15 and 0.8 < z.std() < 1.25, (name, z.mean(), z.std())
    assert abs(st[1] / st[0] - nseg / (W * H * spp)) < 0.05 * nseg / (W * H * spp), name


def test_scene_validation_of_new_fields():
    sc = scenes.cornell_smoke(8, 8)
    sc.media["material"][0] = 0  # a lambertian: not a phase function
    with pytest.raises(ValueError, match="isotropic"):
        Emu(sc)
    sc = scenes.two_perlin_spheres(8, 8)
    sc.materials["albedo2"][0][1] = 3  # perlin table index out of range
    with pytest.raises(ValueError, match="perlin"):
        Emu(sc)
    sc = scenes.cornell_smoke(8, 8)
    sc.profile = D.RT_PROFILE_FINAL_CU
    with pytest.raises(ValueError):
        Emu(sc)
    # non-finite geometry never reaches the builder
    for field, scene_fn, arr in (("center0", scenes.weekend, "spheres"), ("k", scenes.cornell_box, "quads"),
                                 ("v1", scenes.cornell_box, "triangles"), ("density", scenes.cornell_smoke, "media")):
        sc = scene_fn(8, 8)
        a = getattr(sc, arr)
        a[field][0] = np.nan if a[field][0].ndim == 0 else np.full_like(a[field][0], np.inf)
        with pytest.raises(ValueError, match="non-finite|density"):
            Emu(sc)


# ------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("shape", ["sphere", "box"])
def test_gpu_medium_transmittance(shape):
    from a_dive_into_ray_tracing_b200 import capi
    density, W, H, spp = 0.6, 32, 32, 1024
    sc = transmittance_scene(shape, density, W, H)
    want = expected_transmittance(shape, density)
    with capi.Context(profile=2, seed=5) as ctx:
        ctx.upload(sc).build_accel(1)
        ctx.render(W, H, spp)
        lin, _ = ctx.resolve()
    se = np.sqrt(want * (1 - want) / (W * H * spp))
    assert abs(lin.mean() - want) < 4 * se + 1e-3, (shape, lin.mean(), want)


@pytest.mark.gpu
def test_gpu_white_furnace_medium():
    from a_dive_into_ray_tracing_b200 import capi
    mats = np.array([_mat(D.RT_MAT_ISOTROPIC, (1, 1, 1))], D.MATERIAL_DT)
    sc = D.Scene(materials=mats, media=np.array([_medium_sphere((0, 0, 0), 1.0, 3.0, 0)], D.MEDIUM_DT),
                 profile=D.RT_PROFILE_NEXT_WEEK, background=(1, 1, 1), sky_gradient=0, max_depth=500)
    sc.camera = D.camera_from_lookat((0, 0, 5), (0, 0, 0), (0, 1, 0), 30.0, 1.0, 0.0, 5.0, dtype=np.float32)
    with capi.Context(profile=2, seed=5) as ctx:
        ctx.upload(sc).build_accel(1)
        ctx.render(64, 64, 64)
        lin, _ = ctx.resolve()
    np.testing.assert_allclose(lin, 1.0, atol=1e-5)


@pytest.mark.gpu
def test_gpu_textured_emitters_equal_device_code_emulation():
    """Camera looking straight at noise- and image-textured light sources: no scattering, so the
    kernel and the host emulation of the same device code (same Philox streams) must agree to
    rounding — the textures' tables, images and (u, v) reach the kernel intact."""
    from a_dive_into_ray_tracing_b200 import capi
    mats = np.array([_mat(D.RT_MAT_DIFFUSE_LIGHT, (1, 0.8, 0.6), texture=D.RT_TEX_NOISE, albedo2=(3.0, 0, 0)),
                     _mat(D.RT_MAT_DIFFUSE_LIGHT, (1, 1, 1), texture=D.RT_TEX_IMAGE, albedo2=(0, 0, 0))], D.MATERIAL_DT)
    sc = D.Scene(spheres=np.array([_sphere((1.2, 0, 0), 1.0, 1)], D.SPHERE_DT),
                 quads=np.array([_quad(2, -2.5, -0.2, -1, 1, 0.0, 0), _quad(2, -2.5, 2.5, -2.4, -1.2, 0.0, 1)], D.QUAD_DT),
                 materials=mats, perlin=np.array([scenes.make_perlin(3)], D.PERLIN_DT),
                 images=[scenes.procedural_earth(128, 64)], profile=D.RT_PROFILE_NEXT_WEEK, background=(0, 0, 0),
                 sky_gradient=0)
    sc.camera = D.camera_from_lookat((0, 0, 8), (0, -0.5, 0), (0, 1, 0), 40.0, 1.5, 0.0, 8.0, dtype=np.float32)
    W, H, spp = 96, 64, 16
    with capi.Context(profile=2, seed=1984) as ctx:
        ctx.upload(sc).build_accel(1)
        ctx.render(W, H, spp)
        lin, _ = ctx.resolve()
    s, _, _ = Emu(sc).render(W, H, spp, seed=1984)
    ref = (s / spp).astype(np.float32)
    assert ref.max() > 0.5  # the emitters are in view
    # texel-border / lattice-border samples may flip with the last ulp of the hit point
    close = np.abs(lin - ref).max(axis=-1) < 2e-3
    assert close.mean() > 0.99, close.mean()
    assert abs(lin.mean() - ref.mean()) < 2e-4


@pytest.mark.gpu
@pytest.mark.parametrize("name", NEW_SCENES)
def test_gpu_next_week_scenes_vs_oracle(l1_32, name):
    from a_dive_into_ray_tracing_b200 import capi
    W, H = (40, 40) if name in ("cornell_smoke", "next_week_final") else (48, 32)
    spp, K = 256, 8
    sc = getattr(scenes, name)(W, H)
    with capi.Context(profile=2, seed=7) as ctx:
        ctx.upload(sc).build_accel(1)
        batches = []
        for k in range(K):
            ctx.clear()
            ctx.render(W, H, spp // K, spp_begin=k * (spp // K))
            a = ctx.accum().astype(np.float64)
            batches.append(a[..., :3] / a[..., 3:4])
        st = ctx.stats()
    mu_a, var_mean_a = SU.batch_variance(batches)
    r, r2, nseg = l1_32.render_parallel(sc, 2, W, H, spp, seed=5)
    mu_b, var_b = SU.mean_var(r, r2, spp)
    ok, d, b = SU.three_sigma_check(mu_a, var_mean_a, 1, mu_b, var_b, spp)
    assert ok, (name, d, b)
    assert abs(st["segments"] / st["paths"] - nseg / (W * H * spp)) < 0.05 * nseg / (W * H * spp), name
    se = np.sqrt((var_mean_a + var_b / spp).reshape(-1, 3).sum(0)) / (W * H)
    dm = np.abs(mu_a.reshape(-1, 3).mean(0) - mu_b.reshape(-1, 3).mean(0))
    assert np.all(dm < 4 * se + 1e-4), (name, dm, se)


def test_reference_medium_semantics_depend_on_object_order(l1_64):
    """Why the medium deviates from the reference: constant_medium::hit as written
    ignores t_max, so inside hittable_list / bvh_node it overrides a nearer surface whenever it is
    visited after it. One ray, a wall 10 units away inside fog of density 0.05: physically the fog
    scatters first with probability 1 - exp(-0.5) = 39 %; the reference's code gives 39 % if the fog is
    listed before the wall and ~100 % if it is listed after it."""
    mats = np.array([_mat(D.RT_MAT_LAMBERTIAN, (0.5, 0.5, 0.5)), _mat(D.RT_MAT_ISOTROPIC, (1, 1, 1))], D.MATERIAL_DT)
    sc = D.Scene(quads=np.array([_quad(2, -50, 50, -50, 50, -10.0, 0)], D.QUAD_DT), materials=mats,
                 media=np.array([_medium_sphere((0, 0, 0), 100.0, 0.05, 1)], D.MEDIUM_DT),
                 profile=D.RT_PROFILE_NEXT_WEEK, background=(0, 0, 0), sky_gradient=0)
    ray = np.array([0, 0, 0, 0, 0, 0, -1, 0], np.float32)
    physical = 1.0 - np.exp(-0.05 * 10.0)
    before = l1_64.medium_reference_order_demo(sc, ray, order=0)
    after = l1_64.medium_reference_order_demo(sc, ray, order=1)
    assert abs(before - physical) < 0.015, before
    assert after > 0.98, after  # the wall all but disappears: 1 - exp(-0.05 * 100)
    # the semantics used for parity (the scatter event must lie before the closest surface) give the
    # physical answer: an emitting wall seen through absorbing fog keeps exp(-0.5) of its radiance
    mats = np.array([_mat(D.RT_MAT_DIFFUSE_LIGHT, (1, 1, 1)), _mat(D.RT_MAT_ISOTROPIC, (0, 0, 0))], D.MATERIAL_DT)
    sc2 = D.Scene(quads=sc.quads, materials=mats, media=sc.media, profile=D.RT_PROFILE_NEXT_WEEK, background=(0, 0, 0),
                  sky_gradient=0)
    sc2.camera = D.camera_from_lookat((0, 0, 0), (0, 0, -1), (0, 1, 0), 1.0, 1.0, 0.0, 1.0, dtype=np.float32)
    W, H, spp = 8, 8, 512
    want = np.exp(-0.5)
    se = np.sqrt(want * (1 - want) / (W * H * spp))
    s, _, _ = Emu(sc2).render(W, H, spp)
    assert abs(s.mean() / spp - want) < 4 * se + 2e-3, s.mean() / spp
    r, _, _ = l1_64.render_parallel(sc2, 2, W, H, spp, seed=9)
    assert abs(r.mean() / spp - want) < 4 * se + 2e-3, r.mean() / spp


def _reference_medium_scene(W, H, flags):
    """a dense, bright scattering ball lit by a white sky over a dark floor: where the scattered ray restarts
    (scatter point vs boundary entry point) changes how much light leaves the ball on which side"""
    mats = np.array([_mat(D.RT_MAT_ISOTROPIC, (0.95, 0.9, 0.8)), _mat(D.RT_MAT_LAMBERTIAN, (0.1, 0.1, 0.1))], D.MATERIAL_DT)
    sc = D.Scene(quads=np.array([_quad(1, -20, 20, -20, 20, -1.2, 1)], D.QUAD_DT), materials=mats,
                 media=np.array([_medium_sphere((0, 0, 0), 1.0, 2.5, 0)], D.MEDIUM_DT), profile=D.RT_PROFILE_NEXT_WEEK,
                 background=(1, 1, 1), sky_gradient=0, max_depth=50, flags=flags)
    sc.camera = D.camera_from_lookat((0, 0.5, 5), (0, 0, 0), (0, 1, 0), 30.0, W / H, 0.0, 5.0, dtype=np.float32)
    return sc


def test_reference_medium_flag_emulated_device_code_vs_oracle(l1_32):
    """RT_FLAG_REFERENCE_MEDIUM (constant_medium.h:74, rec.p = r.at(rec1.t)): the device code's flag path against the
    oracle's (which is pinned to the reference's own constant_medium::hit, tests/test_cuda_ref_pinning.py)."""
    W, H, spp = 24, 16, 256
    imgs = {}
    for flags in (0, D.RT_FLAG_REFERENCE_MEDIUM):
        sc = _reference_medium_scene(W, H, flags)
        r, r2, _ = l1_32.render_parallel(sc, 2, W, H, spp, seed=3)
        s, s2, _ = Emu(sc).render(W, H, spp, seed=11)
        mu_o, var_o = SU.mean_var(r, r2, spp)
        mu_e, var_e = SU.mean_var(s, s2, spp)
        ok, d, b = SU.three_sigma_check(mu_e, var_e, spp, mu_o, var_o, spp)
        assert ok, (flags, d, b)
        imgs[flags] = (mu_e, var_e)
    # the two semantics are different pictures (many sigma apart on the ball), so the flag is observable
    ball = np.s_[4:12, 8:16]
    diff = np.abs(imgs[0][0][ball].mean() - imgs[D.RT_FLAG_REFERENCE_MEDIUM][0][ball].mean())
    se = np.sqrt((imgs[0][1][ball].mean() + imgs[D.RT_FLAG_REFERENCE_MEDIUM][1][ball].mean()) / (spp * 64 * 3))
    assert diff > 6 * se, (diff, se)


@pytest.mark.gpu
def test_gpu_reference_medium_flag_vs_oracle(l1_32):
    from a_dive_into_ray_tracing_b200 import capi
    W, H, spp, K = 48, 32, 256, 8
    for flags in (0, D.RT_FLAG_REFERENCE_MEDIUM):
        sc = _reference_medium_scene(W, H, flags)
        with capi.Context(profile=2, seed=9) as ctx:
            ctx.upload(sc).build_accel(1)
            batches = []
            for k in range(K):
                ctx.clear()
                ctx.render(W, H, spp // K, spp_begin=k * (spp // K))
                a = ctx.accum().astype(np.float64)
                batches.append(a[..., :3] / a[..., 3:4])
        mu_a, var_mean_a = SU.batch_variance(batches)
        r, r2, _ = l1_32.render_parallel(sc, 2, W, H, spp, seed=3)
        mu_b, var_b = SU.mean_var(r, r2, spp)
        ok, d, b = SU.three_sigma_check(mu_a, var_mean_a, 1, mu_b, var_b, spp)
        assert ok, (flags, d, b)
        dm = np.abs(mu_a.mean() - mu_b.mean())
        se = np.sqrt(var_mean_a.mean() / (W * H * 3) + var_b.mean() / (spp * W * H * 3))
        assert dm < 4 * se + 1e-4, (flags, dm, se)
