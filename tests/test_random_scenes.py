"""Property tests on random scenes (hypothesis): the LBVH builder's invariants and
BVH == brute force closest hit, for the device code stepped on the CPU (always) and through
the C ABI on the GPU (-m gpu). Covers duplicates, coincident and degenerate primitives,
clustered Morton codes, mixed primitive types and every leaf size."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from a_dive_into_ray_tracing_b200 import ctypes_defs as D
from a_dive_into_ray_tracing_b200 import scenes
from a_dive_into_ray_tracing_b200.ctypes_defs import MATERIAL_DT, QUAD_DT, SPHERE_DT, TRIANGLE_DT, Scene
from tests.bvh_checks import check_packed_bvh


def random_scene(seed, n_sph, n_tri, n_quad, cluster, dup):
    rng = np.random.Generator(np.random.Philox(seed))
    mats = np.zeros(3, MATERIAL_DT)
    mats["type"] = [D.RT_MAT_LAMBERTIAN, D.RT_MAT_METAL, D.RT_MAT_DIFFUSE_LIGHT]
    mats["albedo"] = rng.random((3, 3))
    mats["param"] = [0, 0.3, 0]
    spread = 0.05 if cluster else 8.0

    def pos(k):
        return (rng.normal(size=(k, 3)) * spread).astype(np.float32)

    sph = np.zeros(n_sph, SPHERE_DT)
    sph["center0"] = pos(n_sph)
    sph["center1"] = sph["center0"] + (rng.random((n_sph, 3)) * 0.3).astype(np.float32)
    sph["radius"] = (0.05 + rng.random(n_sph) * (0.02 if cluster else 0.8)).astype(np.float32)
    sph["moving"] = rng.integers(0, 2, n_sph)
    sph["time1"] = 1.0
    sph["material"] = rng.integers(0, 3, n_sph)
    if dup and n_sph > 1:
        sph[1:1 + n_sph // 2] = sph[0]  # exact duplicates: identical Morton codes and boxes
    tris = np.zeros(n_tri, TRIANGLE_DT)
    for i in range(n_tri):
        c = pos(1)[0]
        v = c + (rng.normal(size=(3, 3)) * (0.02 if cluster else 0.7)).astype(np.float32)
        n = np.cross(v[1] - v[0], v[2] - v[0])
        if not np.isfinite(n).all() or np.linalg.norm(n) < 1e-12:
            v[2] = v[0] + np.float32([0.1, 0.2, 0.05])
        tris[i] = scenes.triangle_record(v[0], v[1], v[2], -n, -n, -n, int(rng.integers(0, 3)))
    quads = np.zeros(n_quad, QUAD_DT)
    for i in range(n_quad):
        c = pos(1)[0]
        w, h = rng.random(2) * (0.05 if cluster else 3.0) + 0.01
        quads[i] = (int(rng.integers(0, 3)), c[0], c[0] + w, c[1], c[1] + h, c[2], int(rng.integers(0, 3)))
    sc = Scene(spheres=sph, triangles=tris, quads=quads, materials=mats, profile=D.RT_PROFILE_NEXT_WEEK,
               sky_gradient=0, background=(0.5, 0.6, 0.7), t_min=1e-3, name="random")
    sc.camera = D.camera_from_lookat((0, 0, 3 * spread + 3), (0, 0, 0), (0, 1, 0), 60.0, 1.0, 0.0, 1.0, 0.0, 1.0,
                                     dtype=np.float32)
    o = (rng.normal(size=(400, 3)) * (spread * 2 + 1)).astype(np.float32)
    d = rng.normal(size=(400, 3)).astype(np.float32)
    d[:40] = np.float32([1, 0, 0])       # axis-parallel rays (zero direction components)
    d[40:60] = np.float32([0, -1, 0])
    rays = np.zeros((400, 8), np.float32)
    rays[:, 0:3], rays[:, 3], rays[:, 4:7] = o, rng.random(400), d
    return sc, rays


scene_args = dict(seed=st.integers(0, 10 ** 6), n_sph=st.integers(0, 60), n_tri=st.integers(0, 40),
                  n_quad=st.integers(0, 12), cluster=st.booleans(), dup=st.booleans(),
                  max_leaf=st.sampled_from([1, 2, 4, 8]), quality=st.sampled_from([0, 1, 2]))


@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(**scene_args)
def test_emulated_builder_and_traversal(seed, n_sph, n_tri, n_quad, cluster, dup, max_leaf, quality):
    from tests.emu.pyemu import Emu
    sc, rays = random_scene(seed, n_sph, n_tri, n_quad, cluster, dup)
    e = Emu(sc, quality=quality, shuffle=seed & 1, max_leaf=max_leaf)
    nodes, leaf, big = e.accel()
    check_packed_bvh(sc, nodes, leaf, big, max_leaf=max_leaf)
    ia, ta, _ = e.trace(rays, use_accel=1)
    ib, tb, _ = e.trace(rays, use_accel=0)
    np.testing.assert_array_equal(ia, ib)
    np.testing.assert_array_equal(ta, tb)


@pytest.mark.gpu
@settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(**scene_args)
def test_gpu_builder_and_traversal(seed, n_sph, n_tri, n_quad, cluster, dup, max_leaf, quality):
    import os
    from a_dive_into_ray_tracing_b200 import capi
    sc, rays = random_scene(seed, n_sph, n_tri, n_quad, cluster, dup)
    os.environ["B200RT_MAX_LEAF"] = str(max_leaf)
    try:
        with capi.Context(profile=2, seed=seed) as ctx:
            ctx.upload(sc).build_accel(quality)
            nodes, leaf, big = ctx.accel()
            check_packed_bvh(sc, nodes, leaf, big, max_leaf=max_leaf)
            ia, ta = ctx.trace_closest(rays, use_accel=True)
            ib, tb = ctx.trace_closest(rays, use_accel=False)
            np.testing.assert_array_equal(ia, ib)
            np.testing.assert_array_equal(ta, tb)
            ctx.render(32, 32, 4)
            a = ctx.accum()
            assert np.all(a[..., 3] == 4) and np.all(np.isfinite(a))
    finally:
        os.environ.pop("B200RT_MAX_LEAF", None)


# ---------------------------------------------------------------- white furnace with media
def furnace_scene(seed, n_sph, n_tri, n_quad, n_media):
    """Random geometry whose every interaction conserves energy exactly (white lambertian, glass,
    white isotropic media) under a uniform white background: every pixel must come out as 1 -
    any lost or invented path (NaN direction, missed surface, wrong medium interval) shows."""
    from a_dive_into_ray_tracing_b200.scenes import _medium_box, _medium_sphere
    sc, _ = random_scene(seed, n_sph, n_tri, n_quad, cluster=False, dup=False)
    rng = np.random.Generator(np.random.Philox(seed + 77))
    mats = np.zeros(4, MATERIAL_DT)
    mats["type"] = [D.RT_MAT_LAMBERTIAN, D.RT_MAT_DIELECTRIC, D.RT_MAT_LAMBERTIAN, D.RT_MAT_ISOTROPIC]
    mats["albedo"] = 1.0
    mats["param"] = [0, 1.5, 0, 0]
    media = []
    for _ in range(n_media):
        c = rng.normal(size=3) * 3.0
        if rng.random() < 0.5:
            media.append(_medium_sphere(tuple(c), 0.3 + 2.0 * rng.random(), 0.05 + 3.0 * rng.random(), 3))
        else:
            ext = 0.3 + 2.5 * rng.random(3)
            media.append(_medium_box(tuple(-ext), tuple(ext), float(rng.uniform(-180, 180)), tuple(c),
                                     0.05 + 3.0 * rng.random(), 3))
    sc2 = Scene(spheres=sc.spheres, triangles=sc.triangles, quads=sc.quads, materials=mats,
                media=np.array(media, D.MEDIUM_DT) if media else None, profile=D.RT_PROFILE_NEXT_WEEK, sky_gradient=0,
                background=(1.0, 1.0, 1.0), t_min=1e-3, max_depth=50, name="furnace",
                flags=D.RT_FLAG_FLIP_NORMALS if seed & 1 else 0)
    sc2.spheres["moving"] = 0  # (a moving sphere can swallow a ray origin between two bounces: still energy-neutral, but keep the scene static)
    sc2.camera = sc.camera
    return sc2


furnace_args = dict(seed=st.integers(0, 10 ** 6), n_sph=st.integers(0, 30), n_tri=st.integers(0, 20),
                    n_quad=st.integers(0, 8), n_media=st.integers(0, 5))


@settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(**furnace_args)
def test_emulated_white_furnace(seed, n_sph, n_tri, n_quad, n_media):
    from tests.emu.pyemu import Emu
    sc = furnace_scene(seed, n_sph, n_tri, n_quad, n_media)
    s, _, st_ = Emu(sc, quality=1).render(24, 24, 4, seed=seed)
    np.testing.assert_allclose(s / 4, 1.0, atol=2e-5)


@pytest.mark.gpu
@settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(**furnace_args)
def test_gpu_white_furnace(seed, n_sph, n_tri, n_quad, n_media):
    from a_dive_into_ray_tracing_b200 import capi
    sc = furnace_scene(seed, n_sph, n_tri, n_quad, n_media)
    with capi.Context(profile=2, seed=seed) as ctx:
        ctx.upload(sc).build_accel(1)
        ctx.render(48, 48, 8)
        lin, _ = ctx.resolve()
    np.testing.assert_allclose(lin, 1.0, atol=2e-5)
