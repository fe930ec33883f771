"""First-class instancing (rt_group / rt_instance, the two-level BVH) - SURVEY.md 8f rank 1. Replaces the
reference's translate / rotate_y wrappers (rt_next_week/cuda/hittable.h:49-190; per triangle in
triangles/cuda/obj_render.cu:498-511; around the 1000-sphere cluster in rt_next_week/cuda/main.cu:373-381).

CPU half (-m "not gpu"): the device code emulated on the host (tests/emu: the builder's per-thread bodies, the
instance-aware closest-hit walk, shading with instance normals) against
  * what the reference's OWN device code returned for per-object wrapped scenes on a B200
    (tests/golden/cuda_ref_*_inst.npz, oracle/ref_cuda_harness.cu) - ids on robust rays, t within 1e-5;
  * oracle L1's ray-transform form (the reference's definition of an instance);
  * the baked form of the same scenes;
and scene validation, the C++ host layer's flattening.
GPU half (-m gpu): the same through the C ABI - brute force, the plain tree walk and the render kernel's own
traversal -, rendered frames, AOVs, and the transform-only update rt_instances_update."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from a_dive_into_ray_tracing_b200 import capi, scenes
from a_dive_into_ray_tracing_b200 import ctypes_defs as D
from tests.cuda_ref_util import RefCase
from tests.emu.pyemu import Emu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCENES = {
    "cornell_box": lambda inst, w=96, h=96: scenes.cornell_box(w, h, instanced=inst),
    "obj_room": lambda inst, w=96, h=96: scenes.obj_room(width=w, height=h, instanced=inst),
    "next_week_final": lambda inst, w=96, h=96: scenes.next_week_final(w, h, instanced=inst),
}


def per_primitive_wrappers(sc):
    """The scene's instances as the per-object translate(rotate_y()) wrappers the oracle's ray-transform form
    takes (flat list order: spheres, triangles, quads) + the owning instance of every primitive."""
    from oracle.pyoracle import L1
    ns, nt = len(sc.spheres), len(sc.triangles)
    inst = np.zeros(sc.n_prims, L1.INST_DT)
    inst["cos_y"] = 1
    owner = np.full(sc.n_prims, -1, np.int32)
    for k, I in enumerate(sc.instances):
        g = sc.groups[I["group"]]
        for lo, f, c in ((0, "first_sphere", "n_spheres"), (ns, "first_triangle", "n_triangles"),
                         (ns + nt, "first_quad", "n_quads")):
            sl = slice(lo + g[f], lo + g[f] + g[c])
            inst["flag"][sl] = 1
            inst["sin_y"][sl], inst["cos_y"][sl] = I["m"][2], I["m"][0]
            inst["offset"][sl] = I["m"][[3, 7, 11]]
            owner[sl] = k
    return inst, owner


def t_tolerance(rays, t):
    """1e-5 relative in t plus what the float rounding of the ray origin moves the hit point: 2e-6 |o| / |d|"""
    r64 = rays.astype(np.float64)
    return 1e-5 * np.abs(t) + 2e-6 * np.linalg.norm(r64[:, :3], axis=1) / np.linalg.norm(r64[:, 4:7], axis=1)


def flat_to_ids(sc, obj, hit):
    ns, nt, nq = len(sc.spheres), len(sc.triangles), len(sc.quads)
    ids = np.full(len(obj), -1, np.int32)
    for lo, hi, ty in ((0, ns, 0), (ns, ns + nt, 1), (ns + nt, ns + nt + nq, 2)):
        m = (hit == 1) & (obj >= lo) & (obj < hi)
        ids[m] = (ty << 28) | (obj[m] - lo)
    return ids


def probe_rays(sc, W=96, H=96, seed=3):
    """primary rays + as many rays leaving the first hit points in random directions (every octant)"""
    rays = D.primary_rays(sc.camera, W, H, sc.profile)
    rng = np.random.default_rng(seed)
    far = rays.copy()
    t = rng.uniform(0.2, 1.2, len(rays)).astype(np.float32)
    far[:, :3] = rays[:, :3] + t[:, None] * rays[:, 4:7]
    far[:, 4:7] = rng.normal(size=(len(rays), 3)).astype(np.float32)
    far[:, 3] = rng.uniform(0, 1, len(rays)).astype(np.float32)
    return np.concatenate([rays, far])


# ----------------------------------------------------------------------------- CPU: emulated device code
@pytest.mark.parametrize("name", ["nw_cornell_inst", "tri_room_inst"])
def test_two_level_walk_vs_reference_cuda_vectors(name, l1_64):
    """Every object of these cases sits in its own translate(rotate_y()) in the REFERENCE's device code
    (golden vectors). The same scene as groups + instances through the emulated two-level walk: the ids the
    reference returned on every robust ray, t within 1e-5 of the double restatement."""
    c = RefCase(name)
    sc = scenes.wrap_per_primitive(c.scene, c.inst)
    assert len(sc.instances) >= 1 and sc.groups["n_spheres"].sum() + sc.groups["n_triangles"].sum() + sc.groups["n_quads"].sum() >= 2
    ref = c.out("list")
    want = c.prim_ids(ref["obj"])
    h64 = l1_64.pin_hits(c.scene, c.scene.profile, c.rays, c.t_min, inst=c.inst)
    robust = h64["obj"] == ref["obj"]
    assert robust.mean() > 0.99
    e = Emu(sc, max_leaf=1)
    for mode in (1, 0):
        ids, inst, ts = e.trace_inst(c.rays, t_min=c.t_min, use_accel=mode)
        mism = (ids != want) & robust
        assert mism.mean() <= 1e-3, (mode, int(mism.sum()))
        k = robust & (ref["hit"] == 1) & (ids == want)
        r64 = c.rays.astype(np.float64)
        tol = 1e-5 * np.abs(h64["t"]) + 2e-6 * np.linalg.norm(r64[:, :3], axis=1) / np.linalg.norm(r64[:, 4:7], axis=1)
        err = np.abs(ts[k].astype(np.float64) - h64["t"][k])
        assert np.all(err <= tol[k]), (mode, (err / tol[k]).max())
        # the instance reported for a hit is the wrapper of the object that was hit
        _, owner = per_primitive_wrappers(sc)
        np.testing.assert_array_equal(inst[k], owner[ref["obj"][k]])


@pytest.mark.parametrize("name", sorted(SCENES))
def test_two_level_walk_vs_ray_transform_oracle_and_baked(name, l1_64):
    a, b = SCENES[name](False), SCENES[name](True)
    assert len(b.instances) >= 1 and len(a.instances) == 0 and a.n_prims >= b.n_prims
    rays = probe_rays(b)
    inst, owner = per_primitive_wrappers(b)
    h = l1_64.pin_hits(b, b.profile, rays, b.t_min, inst=inst)
    want = flat_to_ids(b, h["obj"], h["hit"])
    eb = Emu(b, max_leaf=1)
    got = {}
    for mode in (1, 0):
        ids, ins, ts = eb.trace_inst(rays, t_min=b.t_min, use_accel=mode)
        got[mode] = (ids, ins, ts)
        mism = ids != want
        assert mism.mean() < 2e-3, (mode, int(mism.sum()))
        k = (want >= 0) & ~mism
        err = np.abs(ts[k].astype(np.float64) - h["t"][k]) / t_tolerance(rays, h["t"])[k]
        assert err.max() <= 1.0, (mode, err.max())
        np.testing.assert_array_equal(ins[k], owner[h["obj"][k]])
        assert (ins[k] >= 0).sum() > 100  # the instanced objects are actually seen
    # the tree is transparent: identical bits with and without it
    np.testing.assert_array_equal(got[0][0], got[1][0])
    np.testing.assert_array_equal(got[0][2], got[1][2])
    np.testing.assert_array_equal(got[0][1], got[1][1])
    # baked form of the same scene: same surfaces (t of the primary rays; ids are numbered differently)
    ea = Emu(a)
    n0 = len(rays) // 2
    ia, ta, _ = ea.trace(rays[:n0], t_min=a.t_min)
    ib, _, tb = got[1]
    assert ((ia >= 0) != (ib[:n0] >= 0)).mean() < 2e-3
    k = (ia >= 0) & (ib[:n0] >= 0)
    err = np.abs(ta[k].astype(np.float64) - tb[:n0][k]) / t_tolerance(rays[:n0], ta)[k]
    assert np.quantile(err, 0.999) <= 2.0


def test_instanced_render_equals_baked_render_emulated():
    """shading of instance hits (normal rotated back to world, world-space texture point): the instanced and the
    baked Cornell box converge to the same image (emulated device code, 3 sigma of the per-pixel estimator)"""
    W = H = 24
    a, b = scenes.cornell_box(W, H), scenes.cornell_box(W, H, instanced=True)
    sa, s2a, _ = Emu(a).render(W, H, 48, seed=5)
    sb, s2b, _ = Emu(b, max_leaf=1).render(W, H, 48, seed=6)
    ma, mb = sa / 48, sb / 48
    va = np.maximum(s2a / 48 - ma ** 2, 0) / 48
    vb = np.maximum(s2b / 48 - mb ** 2, 0) / 48
    d = np.abs(ma - mb).reshape(-1, 3).mean(0)
    bound = 3.0 * np.sqrt(va + vb).reshape(-1, 3).mean(0)
    assert np.all(d <= bound), (d, bound)
    assert ma.mean() > 0.01


def test_scene_validation_of_groups_and_instances():
    sc = scenes.cornell_box(32, 32, instanced=True)

    def broken(edit):
        s = sc.with_camera(sc.camera)
        s.groups, s.instances = sc.groups.copy(), sc.instances.copy()
        edit(s)
        with pytest.raises(ValueError) as e:
            Emu(s)
        return str(e.value)

    def scale(s):
        s.instances["m"][0][0] *= 2.0
    assert "rigid" in broken(scale)

    def mirror(s):
        s.instances["m"][0][[4, 5, 6]] *= -1.0
    assert "mirrored" in broken(mirror)

    def bad_group(s):
        s.instances["group"][1] = 7
    assert "group index" in broken(bad_group)

    def bad_range(s):
        s.groups["n_quads"][0] = 10 ** 6
    assert "primitive range" in broken(bad_range)

    def empty(s):
        s.groups["n_quads"][1] = 0
    assert "empty group" in broken(empty)

    def nan(s):
        s.instances["m"][1][3] = np.nan
    assert "non-finite" in broken(nan)

    w = scenes.weekend(32, 32)
    w.groups = np.zeros(1, D.GROUP_DT)
    with pytest.raises(ValueError) as e:
        Emu(w)
    assert "profile 2" in str(e.value)


def test_cpp_host_layer_flattens_wrappers_as_instances():
    """include/rtx: translate(rotate_y(object)) becomes an rt_group + rt_instance (flat_scene::instancing); the
    Cornell box of apps/scenes.h gives the groups / matrices of scenes.cornell_box(instanced=True); the mesh of
    obj_model - every triangle wrapped on its own, obj_render.cu:498-511 - merges into ONE group."""
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "apps"), "-s", "../build/librtx_host.so"])
    L = C.CDLL(os.path.join(ROOT, "build", "librtx_host.so"))
    L.rtx_host_build.argtypes = [C.c_int, C.c_uint, C.c_double, C.c_char_p]
    vp = C.c_void_p
    L.rtx_host_counts.argtypes = [vp]
    L.rtx_host_counts3.argtypes = [vp]
    L.rtx_host_get.argtypes = [vp, vp, vp, vp, vp]
    L.rtx_host_get3.argtypes = [vp, vp]

    def build(which, obj=b""):
        assert L.rtx_host_build(which, 1, 1.0, obj) == 0
        n, n3 = np.zeros(5, np.int32), np.zeros(2, np.int32)
        L.rtx_host_counts(n.ctypes.data)
        L.rtx_host_counts3(n3.ctypes.data)
        s, t, q = np.zeros(n[0], D.SPHERE_DT), np.zeros(n[1], D.TRIANGLE_DT), np.zeros(n[2], D.QUAD_DT)
        g, i = np.zeros(n3[0], D.GROUP_DT), np.zeros(n3[1], D.INSTANCE_DT)
        L.rtx_host_get(s.ctypes.data, t.ctypes.data, q.ctypes.data, None, None)
        L.rtx_host_get3(g.ctypes.data, i.ctypes.data)
        return s, t, q, g, i

    L.rtx_host_set_instancing(1)
    try:
        s, t, q, g, i = build(3)  # cornell_box
        ref = scenes.cornell_box(instanced=True)
        assert len(t) == 0 and len(q) == len(ref.quads) == 18
        for f in ("axis", "a0", "a1", "b0", "b1", "k"):
            np.testing.assert_array_equal(q[f], ref.quads[f])
        np.testing.assert_array_equal(g, ref.groups)
        np.testing.assert_array_equal(i["group"], ref.instances["group"])
        np.testing.assert_allclose(i["m"], ref.instances["m"], rtol=0, atol=1e-6)
        s, t, q, g, i = build(8)  # the final scene: the sphere cluster is the one instance
        ref = scenes.next_week_final(instanced=True)
        assert len(g) == 1 and g["n_spheres"][0] == 1000 and g["first_sphere"][0] == len(s) - 1000
        np.testing.assert_allclose(i["m"], ref.instances["m"], rtol=0, atol=1e-5)
        assert s["center0"][-1000:].min() >= 0 and s["center0"][-1000:].max() < 165  # object space
        mesh = os.path.join(scenes.DATA_DIR, "blob_2.obj")
        if not os.path.exists(mesh):
            scenes.make_blob_mesh(mesh, 2)
        s, t, q, g, i = build(2, os.fsencode(mesh))  # obj_model: per-triangle wrappers -> one group
        ref = scenes.obj_room(mesh, instanced=True)
        assert len(g) == 1 and len(i) == 1 and g["n_triangles"][0] == len(t) == len(ref.triangles)
        np.testing.assert_allclose(i["m"], ref.instances["m"], rtol=0, atol=1e-6)
        np.testing.assert_allclose(t["v0"], ref.triangles["v0"], rtol=0, atol=5e-7)  # (v * scale in double vs float)
    finally:
        L.rtx_host_set_instancing(0)
    s, t, q, g, i = build(3)
    assert len(g) == 0 and len(i) == 0 and len(t) == 24  # baked: the rotated boxes are triangles


# ----------------------------------------------------------------------------- GPU: the C ABI
gpu = pytest.mark.gpu


@gpu
@pytest.mark.parametrize("name", ["nw_cornell_inst", "tri_room_inst"])
def test_gpu_instances_vs_reference_cuda_vectors(name, l1_64):
    c = RefCase(name)
    sc = scenes.wrap_per_primitive(c.scene, c.inst)
    ref = c.out("list")
    want = c.prim_ids(ref["obj"])
    h64 = l1_64.pin_hits(c.scene, c.scene.profile, c.rays, c.t_min, inst=c.inst)
    robust = h64["obj"] == ref["obj"]
    _, owner = per_primitive_wrappers(sc)
    r64 = c.rays.astype(np.float64)
    tol = 1e-5 * np.abs(h64["t"]) + 2e-6 * np.linalg.norm(r64[:, :3], axis=1) / np.linalg.norm(r64[:, 4:7], axis=1)
    with capi.Context(profile=sc.profile, seed=7) as ctx:
        ctx.upload(sc).build_accel(1)
        for mode in (0, 1, 2):
            ids, inst, ts = ctx.trace_closest(c.rays, t_min=c.t_min, use_accel=mode, with_instances=True)
            mism = (ids != want) & robust
            assert mism.mean() <= 1e-3, (mode, int(mism.sum()))
            k = robust & (ref["hit"] == 1) & (ids == want)
            err = np.abs(ts[k].astype(np.float64) - h64["t"][k])
            assert np.all(err <= tol[k]), (mode, (err / tol[k]).max())
            np.testing.assert_array_equal(inst[k], owner[ref["obj"][k]])


@gpu
@pytest.mark.parametrize("plan", ["2", "1", "3", "0"])
@pytest.mark.parametrize("name", sorted(SCENES))
def test_gpu_two_level_hits(name, plan, l1_64, monkeypatch):
    """instanced scene through the C ABI: brute force == plain walk == render kernel's traversal (ids, instances; t
    bit-identical between brute force and the plain walk), all equal to the oracle's ray-transform form, and the
    same surfaces as the baked scene - for every shared-memory residency plan of the render kernel (each tree
    quantised in its own frame, B200RT_SMEM caps the plan)"""
    monkeypatch.setenv("B200RT_SMEM", plan)
    a, b = SCENES[name](False), SCENES[name](True)
    rays = probe_rays(b)
    inst, owner = per_primitive_wrappers(b)
    h = l1_64.pin_hits(b, b.profile, rays, b.t_min, inst=inst)
    want = flat_to_ids(b, h["obj"], h["hit"])
    with capi.Context(profile=2, seed=1) as ctx:
        ctx.upload(b).build_accel(1)
        res = {m: ctx.trace_closest(rays, t_min=b.t_min, use_accel=m, with_instances=True) for m in (0, 1, 2)}
        got_plan = ctx.stats()["smem_plan"]
        assert got_plan == int(plan) or name == "next_week_final"  # (the final scene's primitives do not fit)
    for m, (ids, ins, ts) in res.items():
        mism = ids != want
        assert mism.mean() < 2e-3, (m, int(mism.sum()))
        k = (want >= 0) & ~mism
        err = np.abs(ts[k].astype(np.float64) - h["t"][k]) / t_tolerance(rays, h["t"])[k]
        assert err.max() <= 1.0, (m, err.max())
        np.testing.assert_array_equal(ins[k], owner[h["obj"][k]])
    for f in range(3):
        np.testing.assert_array_equal(res[0][f], res[1][f])
    same = res[2][0] == res[1][0]
    assert same.mean() > 0.999
    np.testing.assert_array_equal(res[2][1][same], res[1][1][same])
    np.testing.assert_allclose(res[2][2][same], res[1][2][same], rtol=2e-6, atol=0)
    with capi.Context(profile=2, seed=1) as ctx:
        ctx.upload(a).build_accel(1)
        n0 = len(rays) // 2
        ia, ta = ctx.trace_closest(rays[:n0], t_min=a.t_min, use_accel=2)
    ib, tb = res[2][0][:n0], res[2][2][:n0]
    assert ((ia >= 0) != (ib >= 0)).mean() < 2e-3
    k = (ia >= 0) & (ib >= 0)
    assert np.quantile(np.abs(ta[k].astype(np.float64) - tb[k]) / t_tolerance(rays[:n0], ta)[k], 0.999) <= 2.0


@gpu
@pytest.mark.parametrize("name,spp", [("cornell_box", 256), ("obj_room", 128), ("next_week_final", 128)])
def test_gpu_instanced_frame_equals_baked_frame(name, spp):
    """rendered through k_render's INST instantiation (incl. the extended variant for the final scene): the same
    image as the baked scene within 3 sigma of the two estimators, same segments per path"""
    W = H = 96
    a, b = SCENES[name](False, W, H), SCENES[name](True, W, H)
    out = []
    for sc, seed in ((a, 11), (b, 12)):
        with capi.Context(profile=2, seed=seed) as ctx:
            ctx.upload(sc).build_accel(1)
            ctx.render(W, H, spp)
            acc = ctx.accum().astype(np.float64)
            st = ctx.stats()
        out.append((acc[..., :3] / spp, st["segments"] / st["paths"]))
    (ma, sa), (mb, sb) = out
    assert abs(sa - sb) < 0.03 * sa, (sa, sb)
    # per-pixel variance is not returned by the accumulation buffer: compare block means against the spread of the
    # two independent estimates over the blocks (8x8 blocks of 64 pixels x spp samples)
    blk = lambda m: m.reshape(H // 8, 8, W // 8, 8, 3).mean((1, 3))
    da = blk(ma) - blk(mb)
    scale = 0.5 * (blk(ma) + blk(mb)).mean()
    assert np.abs(da.mean((0, 1))).max() < 0.03 * scale + 1e-3, (da.mean((0, 1)), scale)
    assert np.sqrt((da ** 2).mean()) < 0.25 * scale + 1e-3


@gpu
def test_gpu_transform_only_update():
    """rt_instances_update: new matrices, top level rebuilt, the objects' trees untouched - the frame and the hits
    equal a fresh upload + full build of the moved scene; a wrong count / group is rejected"""
    W = H = 64
    sc = scenes.next_week_final(W, H, instanced=True)
    moved = sc.instances.copy()
    moved["m"][0] = D.rigid_y(-40.0, (50.0, 200.0, 300.0))
    sc2 = sc.with_camera(sc.camera)
    sc2.groups, sc2.instances = sc.groups, moved
    rays = probe_rays(sc, W, H)
    with capi.Context(profile=2, seed=3) as ctx:
        ctx.upload(sc).build_accel(1)
        full_ms = ctx.stats()["ms_build"]
        i0, n0, t0 = ctx.trace_closest(rays, t_min=sc.t_min, use_accel=2, with_instances=True)
        ctx.update_instances(moved)
        upd_ms = ctx.stats()["ms_build"]
        i1, n1, t1 = ctx.trace_closest(rays, t_min=sc.t_min, use_accel=2, with_instances=True)
        ctx.render(W, H, 16)
        f1 = ctx.accum()
        with pytest.raises(capi.RtError):
            ctx.update_instances(moved[:0])
        bad = moved.copy()
        bad["group"][0] = 1
        with pytest.raises(capi.RtError):
            ctx.update_instances(bad)
        stretched = moved.copy()
        stretched["m"][0][0] = 3.0
        with pytest.raises(capi.RtError):
            ctx.update_instances(stretched)
        # the failed updates changed nothing
        i1b, n1b, t1b = ctx.trace_closest(rays, t_min=sc.t_min, use_accel=2, with_instances=True)
    with capi.Context(profile=2, seed=3) as ctx:
        ctx.upload(sc2).build_accel(1)
        i2, n2, t2 = ctx.trace_closest(rays, t_min=sc.t_min, use_accel=2, with_instances=True)
        ctx.render(W, H, 16)
        f2 = ctx.accum()
    assert (n0 >= 0).sum() > 50 and (n1 >= 0).sum() > 50
    assert (i0 != i1).mean() > 0.01  # the cluster really moved
    np.testing.assert_array_equal(i1, i2)
    np.testing.assert_array_equal(n1, n2)
    np.testing.assert_array_equal(t1, t2)
    np.testing.assert_array_equal(i1, i1b)
    np.testing.assert_array_equal(f1, f2)  # same seed, same scene, same tree: the same frame bit for bit
    assert upd_ms <= full_ms  # (only the top level is rebuilt)


@gpu
def test_gpu_aov_of_instanced_scene():
    """first-hit normals of instanced objects come back in WORLD space: equal to the baked scene's"""
    W = H = 48
    a, b = scenes.cornell_box(W, H), scenes.cornell_box(W, H, instanced=True)
    out = []
    for sc in (a, b):
        with capi.Context(profile=2, seed=2) as ctx:
            ctx.upload(sc).build_accel(1)
            out.append(ctx.render_aov(W, H, 4))
    fa, fb = out
    assert np.abs(fa[..., 3:6] - fb[..., 3:6]).mean() < 2e-3
    assert np.abs(fa[..., 6] - fb[..., 6]).mean() / fa[..., 6].mean() < 1e-3
    assert np.abs(fb[..., 3:6]).max() > 0.9


@gpu
def test_gpu_render_cli_instancing_and_cpp_update(tmp_path):
    """the C++ host layer end to end: render_cli --instancing renders the final scene with its cluster as an
    instance (same image statistics as the baked run), and a C++ program moves an instance with
    renderer::place_instance + update_instances and gets the frame of a freshly built moved scene"""
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "apps"), "-s"])
    imgs = []
    for extra in ([], ["--instancing"]):
        out = tmp_path / ("final%d.ppm" % len(extra))
        with open(out, "wb") as f:
            p = subprocess.run([os.path.join(ROOT, "build", "render_cli"), "--scene", "final", "--width", "96", "--height",
                                "96", "--spp", "128", "--binary"] + extra, stdout=f, stderr=subprocess.PIPE, text=True,
                               timeout=300)
        assert p.returncode == 0, p.stderr
        assert ("1 instances of 1 objects" if extra else "0 instances of 0 objects") in p.stderr
        raw = out.read_bytes()
        imgs.append(np.frombuffer(raw[len(b"P6\n96 96\n255\n"):], np.uint8).reshape(96, 96, 3).astype(np.float64))
    assert abs(imgs[0].mean() - imgs[1].mean()) < 2.0 and np.abs(imgs[0] - imgs[1]).mean() < 12.0
    src = tmp_path / "move.cpp"
    src.write_text(r"""
#include <cstdio>
#include <cstring>
#include "rtx.h"
using namespace rtx;
int main() {
  auto white = make_shared<lambertian>(color(.73, .73, .73));
  auto light = make_shared<diffuse_light>(color(4, 4, 4));
  auto scene_at = [&](double angle, vec3 off) {
    hittable_list w;
    w.add(make_shared<xz_rect>(-500, 500, -500, 500, 0, white));
    w.add(make_shared<xz_rect>(-200, 200, -200, 200, 500, light));
    w.add(make_shared<translate>(make_shared<rotate_y>(make_shared<box>(point3(0, 0, 0), point3(100, 200, 100), white), angle), off));
    return w;
  };
  camera cam(point3(0, 250, -700), point3(0, 150, 0), vec3(0, 1, 0), 40, 1.0, 0.0, 700.0, 0.0, 1.0, true);
  render_options o;
  o.profile = RT_PROFILE_NEXT_WEEK; o.sky_gradient = false; o.background = color(0.1, 0.1, 0.1); o.instancing = true; o.seed = 9;
  renderer a(o), b(o);
  a.set_scene(scene_at(10, vec3(-150, 0, 0)), cam);
  if (a.flat.instances.size() != 1 || a.flat.groups.size() != 1 || a.flat.quads.size() != 8) return 2;
  a.render(64, 64, 32);
  image8 before = a.resolve();
  a.place_instance(0, 55, vec3(120, 30, 40));
  a.update_instances();
  rt_accum_clear(a.handle());
  a.render(64, 64, 32);
  image8 moved = a.resolve();
  b.set_scene(scene_at(55, vec3(120, 30, 40)), cam);
  b.render(64, 64, 32);
  image8 fresh = b.resolve();
  if (moved.rgb != fresh.rgb) return 3;   // same seed, same scene, same trees: the same bytes
  if (moved.rgb == before.rgb) return 4;  // and the box really moved
  printf("ok\n");
  return 0;
}
""")
    exe = tmp_path / "move"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include", "rtx"), "-I", os.path.join(ROOT, "include"),
                           str(src), "-o", str(exe), "-L", os.path.join(ROOT, "a_dive_into_ray_tracing_b200"), "-lb200rt",
                           "-Wl,-rpath," + os.path.join(ROOT, "a_dive_into_ray_tracing_b200")])
    p = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and "ok" in p.stdout, (p.returncode, p.stdout, p.stderr)
