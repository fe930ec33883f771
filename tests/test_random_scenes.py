"""Property tests on random scenes (hypothesis): the LBVH builder's invariants and
BVH == brute force closest hit, for the device code stepped on the CPU (always) and through
the C ABI on the GPU (-m gpu). Covers duplicates, coincident and degenerate primitives,
clustered Morton codes, mixed primitive types and every leaf size."""
import numpy as np
import pytest
from hypothesis import HealthCheck, example, given, settings
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


# ---------------------------------------------------------------- random two-level scenes (groups, instances, boxes)
def random_rigid(rng, spread):
    """a random rotation (QR of a Gaussian matrix, det +1) and translation as the 3x4 rows of rt_instance.m"""
    q, r = np.linalg.qr(rng.normal(size=(3, 3)))
    q = q * np.sign(np.diag(r))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    m = np.zeros((3, 4))
    m[:, :3] = q
    m[:, 3] = rng.normal(size=3) * spread
    return m.astype(np.float32).ravel()


def random_two_level_scene(seed, n_sph, n_tri, n_quad, n_box, n_groups, n_inst, cluster):
    """random_scene + boxes (six rects in box.h's order, at the world level and inside groups) + groups cut from the
    tails of the primitive arrays + instances with arbitrary rigid matrices (several may place the same group)"""
    from a_dive_into_ray_tracing_b200.scenes import box_as_quads
    sc, rays = random_scene(seed, n_sph, n_tri, n_quad, cluster, False)
    rng = np.random.Generator(np.random.Philox(seed + 4242))
    spread = 0.05 if cluster else 8.0
    quads = list(sc.quads)
    box_first = []
    for _ in range(n_box):
        c = rng.normal(size=3) * spread
        e = rng.random(3) * (0.05 if cluster else 2.0) + 0.01
        box_first.append(len(quads))
        quads += box_as_quads(tuple(np.float32(c)), tuple(np.float32(c + e)), int(rng.integers(0, 3)))
    quads = np.array(quads, QUAD_DT) if quads else np.zeros(0, QUAD_DT)
    ns, nt, nq = len(sc.spheres), len(sc.triangles), len(quads)
    groups = np.zeros(n_groups, D.GROUP_DT)
    # each group takes a slice of the tail of every array (slices of different groups may overlap: shared geometry)
    for g in range(n_groups):
        for f, c, n in (("first_sphere", "n_spheres", ns), ("first_triangle", "n_triangles", nt), ("first_quad", "n_quads", nq)):
            k = int(rng.integers(0, n // 2 + 1))
            lo = int(rng.integers(n // 2, n - k + 1)) if n else 0
            groups[f][g], groups[c][g] = lo, k
    nonempty = [g for g in range(n_groups) if groups["n_spheres"][g] + groups["n_triangles"][g] + groups["n_quads"][g] > 0]
    inst = np.zeros(n_inst if nonempty else 0, D.INSTANCE_DT)
    for i in range(len(inst)):
        inst["m"][i] = random_rigid(rng, spread)
        inst["group"][i] = nonempty[int(rng.integers(0, len(nonempty)))]
    out = Scene(spheres=sc.spheres, triangles=sc.triangles, quads=quads, materials=sc.materials, camera=sc.camera,
                background=sc.background, sky_gradient=0, t_min=1e-3, profile=D.RT_PROFILE_NEXT_WEEK, name="random2",
                groups=groups, instances=inst)
    return out, rays


two_level_args = dict(seed=st.integers(0, 10 ** 6), n_sph=st.integers(0, 40), n_tri=st.integers(0, 30),
                      n_quad=st.integers(0, 8), n_box=st.integers(0, 6), n_groups=st.integers(0, 4),
                      n_inst=st.integers(0, 6), cluster=st.booleans())


@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(**two_level_args)
@example(seed=186, n_sph=0, n_tri=4, n_quad=0, n_box=4, n_groups=3, n_inst=0, cluster=False)  # a group owns rects 2..4 of a box
def test_emulated_two_level_builder_and_traversal(seed, n_sph, n_tri, n_quad, n_box, n_groups, n_inst, cluster):
    """tree walk == list-order brute force, bit for bit (primitive, instance, t), on random two-level scenes"""
    from tests.emu.pyemu import Emu
    sc, rays = random_two_level_scene(seed, n_sph, n_tri, n_quad, n_box, n_groups, n_inst, cluster)
    e = Emu(sc, quality=seed % 3, shuffle=seed & 1, max_leaf=1)
    if len(sc.groups) or len(sc.instances):
        a, b = e.trace_inst(rays, use_accel=1), e.trace_inst(rays, use_accel=0)
    else:
        a, b = e.trace(rays, use_accel=1)[:2], e.trace(rays, use_accel=0)[:2]
    for x, y in zip(a, b):
        np.testing.assert_array_equal(x, y)


@pytest.mark.gpu
@settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(plan=st.sampled_from(["2", "1", "3", "0"]), **two_level_args)
@example(plan="2", seed=186, n_sph=0, n_tri=4, n_quad=0, n_box=4, n_groups=3, n_inst=0, cluster=False)
def test_gpu_two_level_builder_and_traversal(seed, n_sph, n_tri, n_quad, n_box, n_groups, n_inst, cluster, plan):
    """the same through the C ABI; the render kernel's own traversal (every residency plan) agrees with both except
    for exact ties, and a frame renders finite"""
    import os
    from a_dive_into_ray_tracing_b200 import capi
    sc, rays = random_two_level_scene(seed, n_sph, n_tri, n_quad, n_box, n_groups, n_inst, cluster)
    os.environ["B200RT_SMEM"] = plan
    try:
        with capi.Context(profile=2, seed=seed) as ctx:
            ctx.upload(sc).build_accel(seed % 3)
            r0 = ctx.trace_closest(rays, use_accel=0, with_instances=True)
            r1 = ctx.trace_closest(rays, use_accel=1, with_instances=True)
            r2 = ctx.trace_closest(rays, use_accel=2, with_instances=True)
            ctx.render(32, 32, 4)
            a = ctx.accum()
    finally:
        os.environ.pop("B200RT_SMEM", None)
    for x, y in zip(r0, r1):
        np.testing.assert_array_equal(x, y)
    same = (r2[0] == r1[0]) & (r2[1] == r1[1])
    both = (r2[0] >= 0) & (r1[0] >= 0)
    # where the kernel's traversal names another primitive it is a tie: both hit, t equal within the hardware
    # reciprocal's 2 ulp (coincident / duplicated geometry is part of the generator)
    assert np.all(both[~same])
    tol = 1e-5 * np.abs(r1[2]) + 1e-6 * np.linalg.norm(rays[:, :3], axis=1) / np.linalg.norm(rays[:, 4:7], axis=1)
    assert np.all(np.abs(r2[2] - r1[2])[both] <= tol[both])
    assert (~same).mean() <= 0.05
    assert np.all(a[..., 3] == 4) and np.all(np.isfinite(a))


@settings(max_examples=15, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(seed=st.integers(0, 10 ** 6), n_sph=st.integers(1, 20), n_tri=st.integers(0, 12), n_box=st.integers(0, 4),
       n_groups=st.integers(1, 3), n_inst=st.integers(1, 4))
def test_emulated_white_furnace_with_instances(seed, n_sph, n_tri, n_box, n_groups, n_inst):
    """energy conservation through instances: white lambertian / glass objects placed by random rigid matrices under a
    white sky give exactly 1 everywhere - a normal that came back in the wrong space (a lost or doubled ray) shows"""
    from tests.emu.pyemu import Emu
    sc, _ = random_two_level_scene(seed, n_sph, n_tri, 2, n_box, n_groups, n_inst, False)
    mats = np.zeros(3, MATERIAL_DT)
    mats["type"] = [D.RT_MAT_LAMBERTIAN, D.RT_MAT_DIELECTRIC, D.RT_MAT_LAMBERTIAN]
    mats["albedo"] = 1.0
    mats["param"] = [0, 1.5, 0]
    sc.materials = mats
    sc.spheres["moving"] = 0
    sc.background = (1.0, 1.0, 1.0)
    sc.flags = D.RT_FLAG_FLIP_NORMALS if seed & 1 else 0
    s, _, _ = Emu(sc, quality=1, max_leaf=1).render(20, 20, 4, seed=seed)
    np.testing.assert_allclose(s / 4, 1.0, atol=2e-5)
