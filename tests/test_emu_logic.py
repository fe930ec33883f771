"""CPU tests of the HOST LOGIC and of the per-thread device code stepped serially on
the CPU (tests/emu — a debugging aid, not a product path): scene flattening, the LBVH
builder's invariants, threaded traversal == brute force == reference golden hits, and
the sampling/shading code's statistics against the reference golden render."""
import numpy as np
import pytest

from a_dive_into_ray_tracing_b200 import ctypes_defs as D
from a_dive_into_ray_tracing_b200 import scenes
from tests import stats_util as SU
from tests.bvh_checks import check_packed_bvh
from tests.emu.pyemu import Emu


@pytest.fixture(scope="module")
def weekend_emu():
    return Emu(scenes.weekend(400, 225), quality=1, shuffle=1)


def test_lbvh_invariants_weekend(weekend_emu):
    sc = weekend_emu.scene
    nodes, leaf, big = weekend_emu.accel()
    assert big.tolist() == [0]  # the r=1000 ground sphere is split off
    assert len(leaf) == sc.n_prims - 1 and len(nodes) % 2 == 1 and len(nodes) <= 2 * (sc.n_prims - 1) - 1
    check_packed_bvh(sc, nodes, leaf, big)


@pytest.mark.parametrize("quality,shuffle", [(0, 0), (1, 0), (2, 1)])
def test_lbvh_invariants_other_scenes(quality, shuffle):
    for sc in (scenes.next_week(120, 80), scenes.obj_room(width=64, height=64, subdivisions=1)):
        e = Emu(sc, quality=quality, shuffle=shuffle)
        nodes, leaf, big = e.accel()
        check_packed_bvh(sc, nodes, leaf, big)


def test_emulated_closest_hit_matches_reference(weekend_emu, hits_primary, hits_bounce):
    for g, tol_abs in ((hits_primary, 0.0), (hits_bounce, 3e-5)):  # grazing self re-hits are ill-conditioned in fp32
        ib, tb, _ = weekend_emu.trace(g["rays"], use_accel=0)
        ia, ta, cnt = weekend_emu.trace(g["rays"], use_accel=1)
        np.testing.assert_array_equal(ia, ib)       # BVH-independent
        np.testing.assert_array_equal(ta, tb)
        np.testing.assert_array_equal(ia, g["ids"])  # == the reference's primitive ids
        hit = g["ids"] >= 0
        err = np.abs(ta[hit].astype(np.float64) - g["t"][hit])
        assert np.all(err <= 1e-5 * g["t"][hit] + tol_abs)
        assert cnt[0] / len(ia) < 40


def test_emulated_render_statistics(weekend_emu, render_c1):
    W, H, spp = int(render_c1["W"]), int(render_c1["H"]), int(render_c1["spp"])
    e = Emu(scenes.weekend(W, H), quality=1)
    s, s2, st = e.render(W, H, spp, seed=1984)
    assert st[0] == W * H * spp
    mu_a, var_a = SU.mean_var(s, s2, spp)
    mu_b, var_b = SU.mean_var(render_c1["sum"], render_c1["sumsq"], spp)
    ok, d, b = SU.three_sigma_check(mu_a, var_a, spp, mu_b, var_b, spp)
    assert ok, (d, b)
    z = SU.zscores(mu_a, var_a, spp, mu_b, var_b, spp)
    assert abs(z.mean()) < 0.1 and 0.85 < z.std() < 1.15
    assert abs(st[1] / st[0] - int(render_c1["segments"]) / (W * H * spp)) < 0.02


def test_emulated_profiles_vs_l1(l1_32):
    """profiles 1 (final.cu) and 2 (next-week, obj room): emulated device code vs the
    float restatement of the CUDA trees, statistically."""
    cases = [(scenes.final_cu(60, 40), 60, 40, 48), (scenes.next_week(60, 40), 60, 40, 48),
             (scenes.obj_room(width=40, height=40, subdivisions=1), 40, 40, 96),
             (scenes.cornell_box(40, 40), 40, 40, 96)]
    for sc, W, H, spp in cases:
        e = Emu(sc, quality=1)
        s, s2, st = e.render(W, H, spp, seed=7)
        r, r2, nseg = l1_32.render_parallel(sc, sc.profile, W, H, spp, seed=5)
        mu_a, var_a = SU.mean_var(s, s2, spp)
        mu_b, var_b = SU.mean_var(r, r2, spp)
        ok, d, b = SU.three_sigma_check(mu_a, var_a, spp, mu_b, var_b, spp)
        assert ok, (sc.name, d, b)
        z = SU.zscores(mu_a, var_a, spp, mu_b, var_b, spp)
        assert abs(z.mean()) < 0.15 and 0.8 < z.std() < 1.25, (sc.name, z.mean(), z.std())
        assert abs(st[1] / st[0] - nseg / (W * H * spp)) < 0.05 * nseg / (W * H * spp), sc.name


def test_emulated_closest_hit_general_scenes(l1_64):
    """triangles / rects / moving spheres: emulated device code vs the double
    restatement of the reference's tests (ids equal on robust rays, t within 1e-5)."""
    for sc, W, H in ((scenes.next_week(160, 100), 160, 100), (scenes.obj_room(width=96, height=96), 96, 96)):
        e = Emu(sc, quality=1)
        rays = D.primary_rays(sc.camera, W, H, sc.profile, time=0.37)
        ia, ta, _ = e.trace(rays, t_min=sc.t_min, use_accel=1)
        ib, tb, _ = e.trace(rays, t_min=sc.t_min, use_accel=0)
        np.testing.assert_array_equal(ia, ib)
        io, to, _ = l1_64.closest_hit(sc, sc.profile, rays, t_min=sc.t_min)
        mism = ia != io
        assert mism.mean() < 2e-3, (sc.name, mism.sum())
        same = (~mism) & (io >= 0)
        rel = np.abs(ta[same] - to[same]) / to[same]
        assert np.percentile(rel, 99.9) < 1e-5, (sc.name, rel.max())


@pytest.mark.parametrize("max_leaf", [1, 2, 4, 8])
def test_leaf_collapsing(max_leaf, hits_primary):
    """Subtrees with <= max_leaf primitives become one leaf: same hits, fewer nodes."""
    sc = scenes.weekend(400, 225)
    e = Emu(sc, quality=1, shuffle=1, max_leaf=max_leaf)
    nodes, leaf, big = e.accel()
    check_packed_bvh(sc, nodes, leaf, big, max_leaf=max_leaf)
    if max_leaf == 1:
        assert len(nodes) == 2 * 486 - 1
    rays = hits_primary["rays"][:3000]
    ia, ta, cnt = e.trace(rays, use_accel=1)
    np.testing.assert_array_equal(ia, hits_primary["ids"][:3000])
