"""CPU tests (-m "not gpu"): the L1 oracle against the committed golden vectors that
were produced by the UNMODIFIED reference (tools/make_golden.py), and — where the
compiled reference oracle/_ref/libref_l0.so is available — bit-exact pinning of the
restatement against it."""
import os

import numpy as np

from a_dive_into_ray_tracing_b200 import ctypes_defs as D
from a_dive_into_ray_tracing_b200 import scenes
from tests import stats_util as SU


def test_scene_fixture(golden_dir):
    rows = np.load(os.path.join(scenes.DATA_DIR, "weekend_scene.npy"))
    assert rows.shape == (487, 12)
    kinds = np.bincount(rows[:, 4].astype(int))
    assert kinds.tolist() == [394, 64, 29]  # lambertian / metal / dielectric
    assert rows[0, 3] == 1000.0 and tuple(rows[0, :3]) == (0.0, -1000.0, 0.0)
    # sphere[1] as probed in the survey (float-rounded here)
    assert abs(rows[1, 0] - (-10.295210698945448)) < 1e-6 and rows[1, 4] == 1
    assert abs(rows[1, 8] - 0.16761137777939439) < 1e-7
    sc = scenes.weekend(400, 225)
    assert len(sc.spheres) == 487 and len(sc.materials) == 487
    np.testing.assert_array_equal(scenes.rows_from_scene(sc)[:, :9], rows[:, :9])


def test_l1_closest_hit_matches_reference_golden(l1_64, hits_primary, hits_bounce):
    """orc64 closest-hit == the reference's hittable_list::hit, bit-exact (ids AND t)."""
    sc = scenes.weekend(400, 225)
    for g in (hits_primary, hits_bounce):
        ids, ts, _ = l1_64.closest_hit(sc, 0, g["rays"])
        np.testing.assert_array_equal(ids, g["ids"])
        np.testing.assert_array_equal(ts, g["t"])


def test_l1_reference_style_bvh_is_transparent(l1_64, hits_primary):
    """bvh_node (bvh.h:78-196 restated) returns exactly what the list returns."""
    sc = scenes.weekend(400, 225)
    rays = hits_primary["rays"][:4000]
    a, ta, _ = l1_64.closest_hit(sc, 0, rays, mode=0)
    for seed in (1, 1984):
        b, tb, cnt = l1_64.closest_hit(sc, 0, rays, mode=1, bvh_seed=seed)
        np.testing.assert_array_equal(a, b)
        np.testing.assert_array_equal(ta, tb)
        assert cnt[0] / len(rays) < 120  # far fewer than 487 tests per ray


def test_l1_float_arithmetic_is_what_loses_precision(l1_32, hits_primary):
    """The CUDA trees' float formulas (restated in orc32) miss the 1e-5 bound on t —
    the reason the GPU core re-derives the sphere test."""
    sc = scenes.weekend(400, 225)
    ids, ts, _ = l1_32.closest_hit(sc, 0, hits_primary["rays"])
    same = (ids == hits_primary["ids"]) & (ids >= 0)
    rel = np.abs(ts[same] - hits_primary["t"][same]) / hits_primary["t"][same]
    assert (ids != hits_primary["ids"]).mean() < 1e-3
    assert rel.max() > 1e-5  # documented defect of the naive float quadratic


def test_l1_render_statistics_match_reference_golden(l1_64, render_c1):
    """Independent RNG streams: image agrees with the reference within Monte-Carlo error."""
    W, H, spp = int(render_c1["W"]), int(render_c1["H"]), int(render_c1["spp"])
    sc = scenes.weekend(W, H)
    s, s2, nseg = l1_64.render_parallel(sc, 0, W, H, spp, seed=3)
    mu_a, var_a = SU.mean_var(s, s2, spp)
    mu_b, var_b = SU.mean_var(render_c1["sum"], render_c1["sumsq"], spp)
    ok, d, b = SU.three_sigma_check(mu_a, var_a, spp, mu_b, var_b, spp)
    assert ok, (d, b)
    z = SU.zscores(mu_a, var_a, spp, mu_b, var_b, spp)
    assert abs(z.mean()) < 0.1 and 0.8 < z.std() < 1.2
    assert abs(nseg / (W * H * spp) - int(render_c1["segments"]) / (W * H * spp)) < 0.03


def test_quantise_known_answers(l1_64, l1_32):
    # color.h:21-27: int(256*clamp(sqrt(sum/spp),0,0.999))
    assert l1_64.quantise(0, [10.0, 2.5, 0.0], 10).tolist() == [255, 128, 0]
    assert l1_64.quantise(0, [100.0, 0.1, 4.9], 10).tolist() == [255, 25, 179]
    # final.cu:227-229: int(255.99*sqrt(x))
    assert l1_32.quantise(1, [10.0, 2.5, 0.0], 10).tolist() == [255, 127, 0]


def test_bsdf_moments_profiles(l1_64, l1_32):
    """Sanity of the restated BSDFs: lambertian P0 is cosine-distributed about the normal
    (E[cos] = 2/3 for normal + unit vector), P1 uses the ball (E[dir.n] = 1, |dir-n|<=1)."""
    sc = scenes.weekend(400, 225)
    n = np.array([0.0, 1.0, 0.0])
    d_in = np.array([0.3, -1.0, 0.2])
    lam = int(np.where(sc.materials["type"] == D.RT_MAT_LAMBERTIAN)[0][0])
    cos0, dot1 = [], []
    for k in range(4000):
        ok, out, att = l1_64.scatter_one(sc, 0, lam, d_in, n, 100 + k)
        assert ok == 1
        cos0.append(out[1] / np.linalg.norm(out))
        ok, out1, _ = l1_32.scatter_one(sc, 1, lam, d_in, n, 100 + k)
        dot1.append(out1[1])
        assert np.linalg.norm(out1 - n) <= 1.0 + 1e-5
    assert abs(np.mean(cos0) - 2.0 / 3.0) < 0.02
    assert abs(np.mean(dot1) - 1.0) < 0.03
    np.testing.assert_allclose(att, sc.materials["albedo"][lam], rtol=1e-6)


# ---------------------------------------------------------------- pinning vs the compiled reference
def _cam22(c):
    return np.array(list(c.origin) + list(c.lower_left_corner) + list(c.horizontal) + list(c.vertical) + list(c.u) +
                    list(c.v) + list(c.w) + [c.lens_radius], np.float64)


def test_pin_closest_hit_against_compiled_reference(l0, l1_64, golden_dir):
    rows = np.load(os.path.join(scenes.DATA_DIR, "weekend_scene.npy"))
    l0.scene_set(rows)
    sc = scenes.weekend(400, 225)
    rng = np.random.Generator(np.random.Philox(11))
    rays = D.primary_rays(sc.camera, 400, 225, 0, s_jitter=rng.random(), t_jitter=rng.random(),
                          lens=(0.3, -0.2), pixels=rng.integers(0, 400 * 225, 5000))
    r6 = np.concatenate([rays[:, 0:3], rays[:, 4:7]], 1).astype(np.float64)
    i0, t0 = l0.closest_hit(r6)
    i1, t1, _ = l1_64.closest_hit(sc, 0, rays)
    np.testing.assert_array_equal(i0, i1)
    np.testing.assert_array_equal(t0, t1)


def test_pin_render_loop_against_compiled_reference(l0, l1_64, golden_dir):
    """Same rand() stream, same float-rounded scene and camera -> the restatement's whole
    pixel/sample/bounce loop reproduces the reference's sums BIT-EXACTLY."""
    from oracle import pyoracle
    rows = np.load(os.path.join(scenes.DATA_DIR, "weekend_scene.npy"))
    l0.scene_set(rows)
    W, H, spp = 40, 22, 4
    sc = scenes.weekend(W, H)
    l0.camera_override(_cam22(sc.camera))
    try:
        s0, s20, n0 = l0.render(W, H, spp, pyoracle.WEEKEND_CAM13(W / H), seed=5)
        l1_64.srand(5)
        s1, s21, n1 = l1_64.render(sc, 0, W, H, spp, seed=0, libc_rand=True)
    finally:
        l0.camera_override(None)
    assert n0 == n1
    np.testing.assert_array_equal(s0, s1)
    np.testing.assert_array_equal(s20, s21)


def test_pin_worker_and_write_color(l0, l1_64):
    """worker() (main.cpp:267-290) == the harness loop; write_color == quantise."""
    from oracle import pyoracle
    W, H, spp = 24, 12, 2
    cam13 = pyoracle.WEEKEND_CAM13(W / H)
    l0.lib.l0_scene_build(1)
    a, _, nseg = l0.render(W, H, spp, cam13, seed=9, want_sumsq=False)
    secs, nseg_w, b = l0.worker_timed(W, H, spp, cam13, 0, W * H, 1, seed=9, want_sum=True)
    np.testing.assert_array_equal(a, b)
    assert nseg == nseg_w and secs > 0
    for px in (a[3, 5], a[0, 0], a[11, 23], np.array([0.0, 50.0, 1.0])):
        np.testing.assert_array_equal(l0.write_color(px, spp), l1_64.quantise(0, px, spp))


def test_pin_scatter_against_compiled_reference(l0, l1_64):
    sc = scenes.weekend(400, 225)
    n = np.array([0.0, 0.6, 0.8])
    for kind in (0, 1, 2):
        m = int(np.where(sc.materials["type"] == kind)[0][0])
        mat = sc.materials[m]
        for k, d_in in enumerate(([0.3, -1.0, 0.2], [0.1, 0.5, 0.9], [-2.0, -0.1, -0.3])):
            l0.srand(77 + k)
            ok0, o0, a0 = l0.scatter(kind, mat["albedo"].astype(np.float64), float(mat["param"]), d_in, n)
            l1_64.srand(77 + k)
            ok1, o1, a1 = l1_64.scatter_one(sc, 0, m, d_in, n, 0, libc_rand=True)
            assert ok0 == ok1
            np.testing.assert_array_equal(o0, o1)
            np.testing.assert_array_equal(a0, a1)
