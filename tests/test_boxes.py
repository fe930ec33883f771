"""`box` objects as one BVH leaf (RT_PRIM_BOX): six consecutive rects that are the sides of an axis-aligned box
(rt_next_week/cuda/box.h:41-58) are recognised at flatten time (csrc/scene_flatten.h) and share one leaf; the leaf
test runs the six rect formulas on the same numbers, in rect order, with the rects' ids. So closest hits - rect id
AND t - are bit-identical to the six-leaf tree (B200RT_BOXES=0), the tree has a third of the nodes, and the
rt_next_week final scene (400 ground boxes = 2400 of its 3407 primitives) becomes resident in shared memory."""
import numpy as np
import pytest

from a_dive_into_ray_tracing_b200 import capi, scenes
from a_dive_into_ray_tracing_b200 import ctypes_defs as D
from tests.emu.pyemu import Emu
from tests.test_instancing import probe_rays

CASES = {
    "next_week_final": lambda: scenes.next_week_final(64, 64),
    "next_week_final_instanced": lambda: scenes.next_week_final(64, 64, instanced=True),
    "cornell_box_instanced": lambda: scenes.cornell_box(64, 64, instanced=True),
    "cornell_smoke": lambda: scenes.cornell_smoke(64, 64),
}


def _types(leaf):
    return set((np.asarray(leaf).astype(np.uint32) >> 28).tolist())


@pytest.mark.parametrize("name", sorted(CASES))
def test_box_leaves_give_bit_identical_hits_emulated(name, monkeypatch):
    sc = CASES[name]()
    rays = probe_rays(sc, 64, 64)
    out = {}
    for on in ("0", "1"):
        monkeypatch.setenv("B200RT_BOXES", on)
        e = Emu(sc, max_leaf=1)
        nodes, leaf, big = e.accel()
        if len(sc.instances):
            ids, inst, ts = e.trace_inst(rays, t_min=sc.t_min)
        else:
            ids, ts, _ = e.trace(rays, t_min=sc.t_min)
            inst = np.zeros_like(ids)
        out[on] = (ids, inst, ts, len(nodes), _types(leaf))
    a, b = out["0"], out["1"]
    np.testing.assert_array_equal(a[0], b[0])
    np.testing.assert_array_equal(a[1], b[1])
    np.testing.assert_array_equal(a[2], b[2])
    assert (a[0] >= 0).mean() > 0.3 and ((a[0] >> 28) == D.RT_PRIM_QUAD).sum() > 100
    if name == "cornell_smoke":  # its boxes are media boundaries, not primitives: nothing to cluster
        assert D.RT_PRIM_BOX not in b[4]
        return
    assert D.RT_PRIM_BOX in b[4] and D.RT_PRIM_BOX not in a[4]
    if not len(sc.instances):
        assert b[3] < 0.5 * a[3]  # 2400 rect leaves -> 400 box leaves


def test_box_recognition_is_exact():
    """a rect moved by one ulp, a wrong order or a group boundary inside the six rects: no box"""
    sc = scenes.next_week_final(32, 32)
    n_boxes = lambda s: int((np.asarray(Emu(s, max_leaf=1).accel()[1]).astype(np.uint32) >> 28 == D.RT_PRIM_BOX).sum())
    assert n_boxes(sc) == 400
    s2 = sc.with_camera(sc.camera)
    s2.quads = sc.quads.copy()
    s2.quads["a1"][0] = np.nextafter(s2.quads["a1"][0], np.float32(np.inf))
    assert n_boxes(s2) == 399
    s3 = sc.with_camera(sc.camera)
    s3.quads = sc.quads.copy()
    s3.quads[[6, 7]] = s3.quads[[7, 6]]
    assert n_boxes(s3) == 399
    s4 = sc.with_camera(sc.camera)
    s4.groups = np.zeros(1, D.GROUP_DT)
    s4.groups["first_quad"], s4.groups["n_quads"] = 3, 9  # cuts through box 0, holds box 1 entirely
    s4.instances = np.zeros(1, D.INSTANCE_DT)
    s4.instances["m"][0] = D.rigid_y(0.0, (0.0, 500.0, 0.0))
    assert n_boxes(s4) == 399


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_gpu_box_leaves_give_bit_identical_hits(name, monkeypatch):
    """through the C ABI: the plain walk and the render kernel's own traversal return the same (rect id, t) with and
    without box leaves, bit for bit; brute force agrees"""
    sc = CASES[name]()
    rays = probe_rays(sc, 64, 64)
    out = {}
    for on in ("0", "1"):
        monkeypatch.setenv("B200RT_BOXES", on)
        with capi.Context(profile=2, seed=1) as ctx:
            ctx.upload(sc).build_accel(1)
            out[on] = [ctx.trace_closest(rays, t_min=sc.t_min, use_accel=m, with_instances=True) for m in (0, 1, 2)]
            out[on].append(ctx.stats())
    for m in (0, 1, 2):
        a, b = out["0"][m], out["1"][m]
        np.testing.assert_array_equal(a[2], b[2], err_msg="t, mode %d" % m)
        # ids: identical, except that an EXACT tie between coplanar rects of different trees (the instanced Cornell
        # boxes stand on the floor rect: same t bit for bit) may be resolved for the other rect when the trees differ
        diff = a[0] != b[0]
        assert diff.mean() <= 1e-3 and (m != 0 or not diff.any()), (m, int(diff.sum()))
        np.testing.assert_array_equal(a[1][~diff], b[1][~diff])
    for f in range(3):
        np.testing.assert_array_equal(out["1"][0][f], out["1"][1][f])
    if name == "next_week_final":
        assert out["1"][3]["n_nodes"] < 0.5 * out["0"][3]["n_nodes"]
        assert out["1"][3]["smem_plan"] in (1, 2) and out["0"][3]["smem_plan"] == 3  # resident only with box leaves


@pytest.mark.gpu
def test_gpu_final_scene_frame_with_box_leaves(monkeypatch):
    """same seed, same hits -> the same frame bit for bit with and without box leaves (integer pixel sums)"""
    W = H = 64
    sc = scenes.next_week_final(W, H)
    frames = []
    for on in ("0", "1"):
        monkeypatch.setenv("B200RT_BOXES", on)
        with capi.Context(profile=2, seed=4) as ctx:
            ctx.upload(sc).build_accel(1)
            ctx.render(W, H, 16)
            frames.append(ctx.accum())
    np.testing.assert_array_equal(frames[0], frames[1])


def _lattice_scene():
    """a 4x3x4 lattice of boxes that TOUCH (coincident faces of neighbours: exact ties in t between rects of different
    boxes and between the exit side of one box and the entry side of the next), plus two boxes nested exactly"""
    from a_dive_into_ray_tracing_b200.ctypes_defs import MATERIAL_DT, QUAD_DT, Scene
    mats = np.zeros(2, MATERIAL_DT)
    mats["type"] = [D.RT_MAT_LAMBERTIAN, D.RT_MAT_METAL]
    mats["albedo"] = [[0.7, 0.7, 0.7], [0.8, 0.6, 0.2]]
    quads = []
    for i in range(4):
        for j in range(3):
            for k in range(4):
                p0 = (np.float32(i), np.float32(j), np.float32(k))
                p1 = (np.float32(i + 1), np.float32(j + 1), np.float32(k + 1))
                quads += scenes.box_as_quads(p0, p1, (i + j + k) & 1)
    quads += scenes.box_as_quads((np.float32(1), np.float32(1), np.float32(1)), (np.float32(3), np.float32(2), np.float32(3)), 1)
    quads += scenes.box_as_quads((np.float32(1), np.float32(1), np.float32(1)), (np.float32(3), np.float32(2), np.float32(3)), 0)
    sc = Scene(quads=np.array(quads, QUAD_DT), materials=mats, background=(1, 1, 1), sky_gradient=0, t_min=1e-3,
               profile=D.RT_PROFILE_NEXT_WEEK, name="lattice")
    sc.camera = D.camera_from_lookat((2, 1.5, -6), (2, 1.5, 2), (0, 1, 0), 40.0, 1.0, 0.0, 10.0, dtype=np.float32)
    return sc


def _lattice_rays():
    """axis-parallel rays ON the lattice planes and edges, diagonal rays through edges and corners, rays that start on a
    shared face, and random ones"""
    rng = np.random.default_rng(11)
    rays = []
    g = np.arange(0, 4.01, 0.5, dtype=np.float32)
    for ax in range(3):
        for u in g:
            for v in g[:7]:
                for s in (-1.0, 1.0):
                    o = np.zeros(3, np.float32)
                    d = np.zeros(3, np.float32)
                    o[ax] = -2.0 if s > 0 else 6.0
                    o[(ax + 1) % 3], o[(ax + 2) % 3] = u, v
                    d[ax] = s
                    rays.append(np.concatenate([o, [0.0], d, [0.0]]))
    for _ in range(1500):  # through lattice points, in lattice directions
        p = rng.integers(0, 5, 3).astype(np.float32)
        d = rng.integers(-2, 3, 3).astype(np.float32)
        if not d.any():
            d[0] = 1
        o = p - d * np.float32(rng.integers(1, 4))
        rays.append(np.concatenate([o, [0.0], d, [0.0]]))
    for _ in range(1500):  # starting on a shared face
        o = np.array([rng.integers(0, 5), rng.random() * 3, rng.random() * 4], np.float32)
        d = rng.normal(size=3).astype(np.float32)
        rays.append(np.concatenate([o, [0.0], d, [0.0]]))
    o = (rng.random((2000, 3)) * [6, 5, 6] - 1).astype(np.float32)
    d = rng.normal(size=(2000, 3)).astype(np.float32)
    rnd = np.concatenate([o, np.zeros((2000, 1), np.float32), d, np.zeros((2000, 1), np.float32)], axis=1)
    return np.concatenate([np.array(rays, np.float32), rnd]).astype(np.float32)


def test_box_sides_ties_and_edges_emulated(monkeypatch):
    """the box leaf lets its six sides compete among themselves before the winner meets the running hit: on exact ties
    (coincident faces, nested identical boxes) and on edges / corners that must still be the answer of six separate
    rect leaves AND of the list-order brute force, bit for bit"""
    sc, rays = _lattice_scene(), _lattice_rays()
    out = {}
    for on in ("0", "1"):
        monkeypatch.setenv("B200RT_BOXES", on)
        e = Emu(sc, max_leaf=1)
        _, leaf, _ = e.accel()
        ids, ts, _ = e.trace(rays, t_min=sc.t_min)
        ib, tb, _ = e.trace(rays, t_min=sc.t_min, use_accel=0)
        np.testing.assert_array_equal(ids, ib)
        np.testing.assert_array_equal(ts, tb)
        out[on] = (ids, ts, _types(leaf))
    np.testing.assert_array_equal(out["0"][0], out["1"][0])
    np.testing.assert_array_equal(out["0"][1], out["1"][1])
    assert D.RT_PRIM_BOX in out["1"][2] and D.RT_PRIM_BOX not in out["0"][2]
    assert (out["1"][0] >= 0).mean() > 0.5


@pytest.mark.gpu
def test_gpu_box_sides_ties_and_edges(monkeypatch):
    sc, rays = _lattice_scene(), _lattice_rays()
    res = []
    for on in ("0", "1"):
        monkeypatch.setenv("B200RT_BOXES", on)
        with capi.Context(profile=2) as ctx:
            ctx.upload(sc).build_accel(1)
            ia, ta = ctx.trace_closest(rays, use_accel=1)
            ib, tb = ctx.trace_closest(rays, use_accel=0)
        np.testing.assert_array_equal(ia, ib)
        np.testing.assert_array_equal(ta, tb)
        res.append((ia, ta))
    np.testing.assert_array_equal(res[0][0], res[1][0])
    np.testing.assert_array_equal(res[0][1], res[1][1])
