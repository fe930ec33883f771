"""GPU parity of the code that is TIMED: rt_trace_closest(use_accel=2) feeds a caller's ray batch
through the TRACE instantiation of k_render - the same shared-memory residency plans and quadrant
node orderings, warp-voted search bursts, hardware reciprocals in the ray setup, direct leaf
payloads and regeneration by ballot rank as rt_render - and returns (primitive id, t).

Reference semantics: hittable_list::hit (rt_in_one_weekend/hittable_list.h:20-34) via the
committed golden vectors of the compiled reference, and brute force over the flattened arrays for
the CUDA-tree scenes. Tolerances: ids identical on the golden rays; on random rays ids identical
except exact-tie silhouettes (both candidates within tolerance in t); t within 1e-5 relative, plus - for
rays that START next to a surface inside a scene with coordinates in the hundreds, where t is tiny
against the ray's own coordinates - 1e-6 of the origin's magnitude in hit-point position (fp32 can
place the origin itself only to 6e-8 of that)."""
import zlib

import numpy as np
import pytest

from a_dive_into_ray_tracing_b200 import capi, ctypes_defs as D, scenes

pytestmark = pytest.mark.gpu

PLANS = ("2", "1", "3", "0")


def _quadrant(rays):
    return (rays[:, 4] < 0).astype(int) | ((rays[:, 6] < 0).astype(int) << 1)


def _scene_rays(sc, n, seed, W=64, H=48):
    """camera rays + rays started inside the scene's bounds in uniformly random directions (all four
    (d.x, d.z) sign quadrants, every octant), with random times in the shutter interval."""
    rng = np.random.default_rng(seed)
    cam = D.primary_rays(sc.camera, W, H, sc.profile)
    lo = np.array([np.inf] * 3)
    hi = -lo
    for s in sc.spheres:
        if abs(s["radius"]) < 100:
            lo = np.minimum(lo, np.array(s["center0"]) - abs(s["radius"]))
            hi = np.maximum(hi, np.array(s["center0"]) + abs(s["radius"]))
    for t in sc.triangles:
        for k in ("v0", "v1", "v2"):
            lo = np.minimum(lo, t[k])
            hi = np.maximum(hi, t[k])
    if not np.all(np.isfinite(lo)):
        lo, hi = np.array([-5.0, 0.0, -5.0]), np.array([5.0, 5.0, 5.0])
    o = lo + (hi - lo) * rng.random((n, 3))
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d *= rng.uniform(0.5, 3.0, size=(n, 1))  # un-normalised directions, as scattered rays are
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:3] = o
    rays[:, 3] = rng.random(n)
    rays[:, 4:7] = d
    return np.concatenate([cam, rays]).astype(np.float32)


def _compare(ia, ta, ib, tb, rays, max_tie_frac=2e-4):
    both = (ia >= 0) & (ib >= 0)
    err = np.zeros(len(ia))
    err[both] = np.abs(ta[both].astype(np.float64) - tb[both])
    r64 = rays.astype(np.float64)
    tol = 1e-5 * np.abs(tb) + 1e-6 * np.linalg.norm(r64[:, 0:3], axis=1) / np.linalg.norm(r64[:, 4:7], axis=1)
    mism = ia != ib
    # a different id is only acceptable as an exact-tie silhouette: both hits, same t within tolerance
    assert np.all(both[mism]) and (err[mism] <= tol[mism]).all(), (int(mism.sum()), err[mism].max() if mism.any() else 0)
    assert len(ia) == 0 or mism.mean() <= max_tie_frac, mism.mean()
    assert np.all(err <= tol), (err / np.maximum(tol, 1e-30)).max()


@pytest.mark.parametrize("plan", PLANS)
def test_render_kernel_traversal_vs_reference_golden(plan, hits_primary, hits_bounce, monkeypatch):
    """ids identical to the compiled reference's on all 11 600 golden rays and t within 1e-5 relative, for
    every residency plan of k_render."""
    monkeypatch.setenv("B200RT_SMEM", plan)
    with capi.Context(profile=0, seed=1984) as ctx:
        ctx.upload(scenes.weekend(400, 225)).build_accel(1)
        for g, tol_abs in ((hits_primary, 0.0), (hits_bounce, 3e-5)):
            ia, ta = ctx.trace_closest(g["rays"], use_accel=2)
            assert ctx.stats()["smem_plan"] == int(plan)
            np.testing.assert_array_equal(ia, g["ids"])
            hit = g["ids"] >= 0
            err = np.abs(ta[hit].astype(np.float64) - g["t"][hit])
            assert np.all(err <= 1e-5 * g["t"][hit] + tol_abs), err.max()
            ib, tb = ctx.trace_closest(g["rays"], use_accel=0)
            np.testing.assert_array_equal(ia, ib)
    # the golden bounce rays cover all four node orderings
    assert set(_quadrant(hits_bounce["rays"]).tolist()) == {0, 1, 2, 3}


@pytest.mark.parametrize("plan", PLANS)
@pytest.mark.parametrize("name", ["weekend", "final_cu", "next_week", "obj_room", "cornell_box", "next_week_final"])
def test_render_kernel_traversal_vs_brute_force(name, plan, monkeypatch):
    """Sphere-only kernels (profiles 0, 1) and the general kernel (moving spheres, triangles, rects) against
    the list-order brute force, 20 000 rays in all quadrants, every residency plan."""
    monkeypatch.setenv("B200RT_SMEM", plan)
    sc = getattr(scenes, name)(width=64, height=48)
    rays = _scene_rays(sc, 20000, seed=zlib.crc32(name.encode()) & 0xffff)
    q = _quadrant(rays)
    assert min(np.bincount(q, minlength=4)) > 1000
    t_min = float(sc.t_min)
    with capi.Context(profile=sc.profile, seed=3) as ctx:
        ctx.upload(sc).build_accel(1)
        ia, ta = ctx.trace_closest(rays, t_min=t_min, use_accel=2)
        ib, tb = ctx.trace_closest(rays, t_min=t_min, use_accel=0)
        ic, tc = ctx.trace_closest(rays, t_min=t_min, use_accel=1)
    # cornell_box: the baked boxes stand ON the floor rect - their bottom faces are coplanar with it, so rays that
    # start inside a box and go down meet an exact tie between a triangle and the rect, and the rect's t comes from
    # the ray's hardware reciprocal in the render kernel (2 ulp): more (true) ties than silhouettes alone give
    _compare(ia, ta, ib, tb, rays, max_tie_frac=5e-3 if name == "cornell_box" else 2e-4)
    np.testing.assert_array_equal(ic, ib)
    assert (ia >= 0).mean() > 0.3


def test_render_kernel_traversal_ragged_batches(monkeypatch):
    """empty, single-ray, non-multiple-of-32 batches; item sizes from one ray per work item (every item
    overlaps the previous one's drain) to one item for the whole batch; t_max clipping."""
    sc = scenes.weekend(400, 225)
    rays = _scene_rays(sc, 3001, seed=9)
    with capi.Context(profile=0, seed=1) as ctx:
        ctx.upload(sc).build_accel(1)
        ref_i, ref_t = ctx.trace_closest(rays, use_accel=0)
        for n in (0, 1, 31, 33, 257, len(rays)):
            for item in ("1", "32", "100", "256", "100000"):
                monkeypatch.setenv("B200RT_TRACE_ITEM", item)
                ia, ta = ctx.trace_closest(rays[:n], use_accel=2)
                _compare(ia, ta, ref_i[:n], ref_t[:n], rays[:n], max_tie_frac=1.0 if n < 1000 else 2e-3)
        monkeypatch.delenv("B200RT_TRACE_ITEM")
        ia, ta = ctx.trace_closest(rays, t_max=6.0, use_accel=2)
        ib, tb = ctx.trace_closest(rays, t_max=6.0, use_accel=0)
        _compare(ia, ta, ib, tb, rays)
        assert np.all(ta[ia >= 0] < 6.0) and (ia < 0).sum() > (ref_i < 0).sum()


def test_render_kernel_traversal_tiny_scenes():
    base = scenes.weekend(64, 36)
    rays = D.primary_rays(base.camera, 64, 36, 0)
    for keep in ([], [5], [5, 9], [0], [0, 486]):
        sc = scenes.scene_from_rows(scenes.rows_from_scene(base)[keep].reshape(-1, 12))
        sc.camera = base.camera
        with capi.Context(profile=0) as ctx:
            ctx.upload(sc).build_accel(1)
            ia, ta = ctx.trace_closest(rays, use_accel=2)
            ib, tb = ctx.trace_closest(rays, use_accel=0)
            _compare(ia, ta, ib, tb, rays, max_tie_frac=1e-3)


@pytest.mark.parametrize("plan", PLANS)
def test_render_kernel_traversal_degenerate_directions(plan, monkeypatch):
    """The fp16 slope pair of the FHFMA step at its edges: directions with one or two components exactly zero (the
    slope saturates to max-finite / infinity and 0 * inf = NaN must be dropped by min / max), tiny components,
    origins exactly on box planes, and direction scales from 1e-4 to 1e4 (slopes near the fp16 denormals and beyond
    the fp16 range): ids as brute force, t within tolerance, in every residency plan and both kernel families."""
    monkeypatch.setenv("B200RT_SMEM", plan)
    rng = np.random.default_rng(77)
    for name in ("weekend", "obj_room"):
        sc = scenes.weekend(400, 225) if name == "weekend" else scenes.obj_room(width=64, height=64, mesh="blob968")
        base = _scene_rays(sc, 6000, seed=3)[-6000:]
        rays = []
        for k in range(6):  # zero out one component (k < 3) or two (k >= 3)
            r = base[k * 600:(k + 1) * 600].copy()
            if k < 3:
                r[:, 4 + k] = 0.0
            else:
                r[:, 4 + (k - 3)] = 0.0
                r[:, 4 + (k - 2) % 3] = -0.0
            rays.append(r)
        tiny = base[3600:4200].copy()
        idx = rng.integers(0, 3, 600)
        tiny[np.arange(600), 4 + idx] = rng.choice([1e-30, -1e-30, 1e-12, -1e-12, 3e-8, -3e-8], 600)
        rays.append(tiny)
        for scale in (1e-4, 1e-2, 1e2, 1e4):
            r = base[4200:4800].copy()
            r[:, 4:7] *= scale
            rays.append(r)
        snap = base[4800:5400].copy()  # origins snapped to a coarse grid: many lie exactly on box / primitive planes
        snap[:, 0:3] = np.round(snap[:, 0:3] * 2.0) / 2.0
        rays.append(snap)
        rays = np.concatenate(rays).astype(np.float32)
        with capi.Context(profile=sc.profile, seed=5) as ctx:
            ctx.upload(sc).build_accel(1)
            ia, ta = ctx.trace_closest(rays, use_accel=2)
            ib, tb = ctx.trace_closest(rays, use_accel=0)
        # (a different id only as an exact tie - both hit, same t within tolerance; a hit is never lost)
        _compare(ia, ta, ib, tb, rays, max_tie_frac=5e-3)
        assert (ib >= 0).mean() > 0.3
