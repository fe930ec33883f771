"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the
oracle and the committed reference golden vectors. Nothing here reads /root/reference."""
import numpy as np
import pytest

from a_dive_into_ray_tracing_b200 import capi, ctypes_defs as D, scenes
from tests import stats_util as SU
from tests.bvh_checks import check_packed_bvh, sah_cost

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def weekend_ctx():
    ctx = capi.Context(profile=0, seed=1984)
    ctx.upload(scenes.weekend(400, 225)).build_accel(1)
    yield ctx
    ctx.close()


def test_native_library_is_loaded():
    lib = capi.load_library()
    assert lib.rt_device_count() >= 1
    assert any("libb200rt.so" in l for l in open("/proc/self/maps").read().splitlines())


def test_closest_hit_vs_reference_golden(weekend_ctx, hits_primary, hits_bounce):
    """Primary rays: primitive ids equal to the reference's (all of them, bit-exact ids),
    t within 1e-5 relative; brute force and BVH give identical bits."""
    for g, tol_abs in ((hits_primary, 0.0), (hits_bounce, 3e-5)):
        ib, tb = weekend_ctx.trace_closest(g["rays"], use_accel=False)
        ia, ta = weekend_ctx.trace_closest(g["rays"], use_accel=True)
        np.testing.assert_array_equal(ia, ib)
        np.testing.assert_array_equal(ta, tb)
        np.testing.assert_array_equal(ia, g["ids"])
        hit = g["ids"] >= 0
        err = np.abs(ta[hit].astype(np.float64) - g["t"][hit])
        assert np.all(err <= 1e-5 * g["t"][hit] + tol_abs), err.max()


def test_closest_hit_full_grid_vs_oracle(weekend_ctx, l1_64):
    """Every pixel-centre primary ray of config 1 (400x225) + jittered lens rays."""
    sc = scenes.weekend(400, 225)
    for lens, jit in (((0.0, 0.0), (0.5, 0.5)), ((0.6, -0.3), (0.13, 0.82))):
        rays = D.primary_rays(sc.camera, 400, 225, 0, s_jitter=jit[0], t_jitter=jit[1], lens=lens)
        ia, ta = weekend_ctx.trace_closest(rays, use_accel=True)
        io, to, _ = l1_64.closest_hit(sc, 0, rays)
        mism = ia != io
        assert mism.sum() <= 2, mism.sum()  # only measure-zero silhouette ties may differ
        same = (~mism) & (io >= 0)
        rel = np.abs(ta[same] - to[same]) / to[same]
        assert rel.max() < 1e-5, rel.max()


def test_bvh_structure_and_quality(weekend_ctx, l1_64, hits_primary):
    sc = scenes.weekend(400, 225)
    nodes, leaf, big = weekend_ctx.accel()
    assert big.tolist() == [0] and len(leaf) == 486 and len(nodes) <= 971
    check_packed_bvh(sc, nodes, leaf, big)
    # the oracle traversing the SAME packed tree finds the same hits and counts the work
    io, to, cnt = l1_64.closest_hit_packed(sc, 0, nodes, leaf, big, hits_primary["rays"])
    np.testing.assert_array_equal(io, hits_primary["ids"])
    assert cnt[0] / len(io) < 40 and cnt[1] / len(io) < 12
    # SAH refinement does not make the tree worse than plain LBVH
    weekend_ctx.build_accel(0)
    n0, l0_, _ = weekend_ctx.accel()
    check_packed_bvh(sc, n0, l0_, big)
    weekend_ctx.build_accel(1)
    n1, _, _ = weekend_ctx.accel()
    assert sah_cost(n1) <= sah_cost(n0) * 1.0001


def test_bvh_edge_cases():
    """empty scene, one primitive, two primitives, only-big scene."""
    base = scenes.weekend(64, 36)
    rays = D.primary_rays(base.camera, 64, 36, 0)
    for keep in ([], [5], [5, 9], [0], [0, 486]):
        sc = scenes.scene_from_rows(scenes.rows_from_scene(base)[keep].reshape(-1, 12))
        sc.camera = base.camera
        with capi.Context(profile=0) as ctx:
            ctx.upload(sc).build_accel(1)
            nodes, leaf, big = ctx.accel()
            check_packed_bvh(sc, nodes, leaf, big)
            ia, ta = ctx.trace_closest(rays, use_accel=True)
            ib, tb = ctx.trace_closest(rays, use_accel=False)
            np.testing.assert_array_equal(ia, ib)
            np.testing.assert_array_equal(ta, tb)
            if not keep:
                assert np.all(ia == -1)
            ctx.render(64, 36, 4)
            a = ctx.accum()
            assert np.all(a[..., 3] == 4) and np.all(np.isfinite(a))


def test_render_vs_reference_golden(render_c1):
    """Converged-image parity (config-1 view, 100x56 @ 64 spp) against the reference's own
    render: per-channel mean |delta| <= 3 sigma, unbiased z-scores, same path length."""
    W, H, spp = int(render_c1["W"]), int(render_c1["H"]), int(render_c1["spp"])
    K = 8
    with capi.Context(profile=0, seed=1984) as ctx:
        ctx.upload(scenes.weekend(W, H)).build_accel(1)
        batches = []
        for k in range(K):  # K independent sample ranges -> variance from batch means
            ctx.clear()
            ctx.render(W, H, spp // K, spp_begin=k * (spp // K))
            a = ctx.accum().astype(np.float64)
            assert np.all(a[..., 3] == spp // K)
            batches.append(a[..., :3] / a[..., 3:4])
        st = ctx.stats()
    mu_a, var_mean_a = SU.batch_variance(batches)
    mu_b, var_b = SU.mean_var(render_c1["sum"], render_c1["sumsq"], spp)
    ok, d, b = SU.three_sigma_check(mu_a, var_mean_a, 1, mu_b, var_b, spp)
    assert ok, (d, b)
    z = SU.zscores(mu_a, np.maximum(var_mean_a, 1e-12), 1, mu_b, var_b, spp)
    assert abs(np.median(z)) < 0.1
    assert st["paths"] == W * H * spp
    assert abs(st["segments"] / st["paths"] - int(render_c1["segments"]) / (W * H * spp)) < 0.02
    psnr = SU.psnr(SU.gamma(mu_a), SU.gamma(mu_b))
    assert psnr > 32.4, psnr  # reference-vs-reference floor at 64 spp is 33.4 dB (weekend_meta.json)


def test_render_matches_oracle_all_profiles(l1_32, l1_64):
    cases = [(scenes.weekend(60, 40), l1_64, 60, 40, 64), (scenes.final_cu(60, 40), l1_32, 60, 40, 64),
             (scenes.next_week(60, 40), l1_32, 60, 40, 64),
             (scenes.obj_room(width=40, height=40, subdivisions=1), l1_32, 40, 40, 128),
             (scenes.cornell_box(40, 40), l1_32, 40, 40, 128)]
    for sc, orc, W, H, spp in cases:
        K = 8
        with capi.Context(profile=sc.profile, seed=7) as ctx:
            ctx.upload(sc).build_accel(1)
            batches = []
            for k in range(K):
                ctx.clear()
                ctx.render(W, H, spp // K, spp_begin=k * (spp // K))
                a = ctx.accum().astype(np.float64)
                batches.append(a[..., :3] / a[..., 3:4])
            st = ctx.stats()
        mu_a, var_mean_a = SU.batch_variance(batches)
        r, r2, nseg = orc.render_parallel(sc, sc.profile, W, H, spp, seed=5)
        mu_b, var_b = SU.mean_var(r, r2, spp)
        ok, d, b = SU.three_sigma_check(mu_a, var_mean_a, 1, mu_b, var_b, spp)
        assert ok, (sc.name, d, b)
        assert abs(st["segments"] / st["paths"] - nseg / (W * H * spp)) < 0.05 * nseg / (W * H * spp), sc.name
        # global means agree within 4 standard errors
        se = np.sqrt((var_mean_a + var_b / spp).reshape(-1, 3).sum(0)) / (W * H)
        dm = np.abs(mu_a.reshape(-1, 3).mean(0) - mu_b.reshape(-1, 3).mean(0))
        assert np.all(dm < 4 * se + 1e-4), (sc.name, dm, se)


def test_closest_hit_general_scenes(l1_64):
    for sc, W, H in ((scenes.next_week(160, 100), 160, 100), (scenes.obj_room(width=96, height=96), 96, 96)):
        with capi.Context(profile=2) as ctx:
            ctx.upload(sc).build_accel(1)
            nodes, leaf, big = ctx.accel()
            check_packed_bvh(sc, nodes, leaf, big)
            rays = D.primary_rays(sc.camera, W, H, sc.profile, time=0.37)
            ia, ta = ctx.trace_closest(rays, t_min=sc.t_min, use_accel=True)
            ib, tb = ctx.trace_closest(rays, t_min=sc.t_min, use_accel=False)
        np.testing.assert_array_equal(ia, ib)
        np.testing.assert_array_equal(ta, tb)
        io, to, _ = l1_64.closest_hit(sc, sc.profile, rays, t_min=sc.t_min)
        mism = ia != io
        assert mism.mean() < 2e-3, (sc.name, mism.sum())
        same = (~mism) & (io >= 0)
        rel = np.abs(ta[same] - to[same]) / to[same]
        assert np.percentile(rel, 99.9) < 1e-5, (sc.name, rel.max())


def test_determinism_and_sample_split():
    """Same seed -> identical bits run to run; a frame rendered as [0,32)+[32,64) equals
    [0,64) up to fp32 summation order (counter-based RNG); different seeds differ."""
    W, H = 96, 64
    sc = scenes.weekend(W, H)
    outs = []
    for seed, split in ((5, False), (5, False), (5, True), (6, False)):
        with capi.Context(profile=0, seed=seed) as ctx:
            ctx.upload(sc).build_accel(1)
            if split:
                ctx.render(W, H, 32, 0)
                ctx.render(W, H, 32, 32)
            else:
                ctx.render(W, H, 64, 0)
            outs.append(ctx.accum())
    np.testing.assert_array_equal(outs[0], outs[1])
    np.testing.assert_allclose(outs[0], outs[2], rtol=2e-6, atol=1e-5)
    assert np.abs(outs[0] - outs[3]).mean() > 1e-3


def test_determinism_full_size_frame_many_items_per_warp():
    """Run-to-run bit identity where it is hard: 1200x800 x 24 spp is ~60 000 work items for 4 736 resident warps,
    handed out by a racing atomic counter and overlapped inside each warp, so which lane traces which sample and the
    order in which a pixel's paths end differ from run to run. The per-warp accumulators are 64-bit fixed-point
    integer sums (order-independent), the partial frames are combined in a fixed order: identical bits."""
    W, H, spp = 1200, 800, 24
    sc = scenes.weekend(W, H)
    outs = []
    for _ in range(3):
        with capi.Context(profile=0, seed=5) as ctx:
            ctx.upload(sc).build_accel(1)
            ctx.render(W, H, spp)
            outs.append(ctx.accum())
    assert np.all(outs[0][..., 3] == spp)
    np.testing.assert_array_equal(outs[0], outs[1])
    np.testing.assert_array_equal(outs[0], outs[2])
    # the general (profile 2) kernel too
    sc2 = scenes.cornell_box(300, 300)
    outs = []
    for _ in range(2):
        with capi.Context(profile=sc2.profile, seed=5) as ctx:
            ctx.upload(sc2).build_accel(1)
            ctx.render(300, 300, 64)
            outs.append(ctx.accum())
    np.testing.assert_array_equal(outs[0], outs[1])


def test_row_band_split_equals_full_frame():
    """rt_render_rows_device: bands [0,24) + [24,40) + [40,54) of a frame, rendered into one zeroed
    device buffer, equal the full-frame render up to fp32 summation order (the sample chunks differ);
    rows outside a band are not touched."""
    import torch
    W, H, spp = 96, 54, 32
    sc = scenes.weekend(W, H)
    with capi.Context(profile=0, seed=5) as ctx:
        ctx.upload(sc).build_accel(1)
        ctx.render(W, H, spp)
        full = ctx.accum()
        buf = torch.zeros(H, W, 4, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        ctx.render_rows_device(W, H, 24, 40, spp, 0, buf.data_ptr(), st)
        torch.cuda.synchronize()
        part = buf.cpu().numpy()
        assert np.all(part[:24] == 0) and np.all(part[40:] == 0) and np.all(part[24:40, :, 3] == spp)
        ctx.render_rows_device(W, H, 0, 24, spp, 0, buf.data_ptr(), st)
        ctx.render_rows_device(W, H, 40, 54, spp, 0, buf.data_ptr(), st)
        ctx.render_rows_device(W, H, 54, 54, spp, 0, buf.data_ptr(), st)  # empty band: no-op
        torch.cuda.synchronize()
        np.testing.assert_allclose(buf.cpu().numpy(), full, rtol=2e-6, atol=1e-5)
        with pytest.raises(capi.RtError):
            ctx.render_rows_device(W, H, 10, 60, spp, 0, buf.data_ptr(), st)


def test_resolve_matches_write_color(l1_64, l1_32):
    W, H = 64, 40
    for sc, orc in ((scenes.weekend(W, H), l1_64), (scenes.final_cu(W, H), l1_32)):
        with capi.Context(profile=sc.profile) as ctx:
            ctx.upload(sc)
            ctx.render(W, H, 16)
            a = ctx.accum()
            lin, rgb = ctx.resolve()
        np.testing.assert_allclose(lin, a[..., :3] / a[..., 3:4], rtol=1e-6)
        assert rgb.shape == (H, W, 3)
        exp = np.zeros((H, W, 3), np.int32)
        for j in range(H):
            for i in range(W):
                exp[H - 1 - j, i] = orc.quantise(sc.profile, a[j, i, :3].astype(np.float64), 16)
        diff = np.abs(exp - rgb.astype(np.int32))
        assert diff.max() <= 1 and (diff > 0).mean() < 0.01  # float vs double sqrt at bin edges


def test_checkpoint_resume():
    W, H = 48, 32
    sc = scenes.weekend(W, H)
    with capi.Context(profile=0, seed=3) as ctx:
        ctx.upload(sc)
        ctx.render(W, H, 8, 0)
        half = ctx.accum()
        ctx.render(W, H, 8, 8)
        full = ctx.accum()
    with capi.Context(profile=0, seed=3) as ctx2:
        ctx2.upload(sc)
        ctx2.accum_upload(half)
        ctx2.render(W, H, 8, 8)
        np.testing.assert_array_equal(ctx2.accum(), full)


def test_error_paths():
    with capi.Context(profile=0) as ctx:
        with pytest.raises(capi.RtError) as e:
            ctx.render(32, 32, 1)
        assert e.value.code == 3
        sc = scenes.next_week(32, 32)  # moving spheres + checker need profile 2
        with pytest.raises(capi.RtError) as e:
            ctx.upload(sc)
        assert e.value.code == 1 and "profile 2" in str(e.value)
        bad = scenes.weekend(32, 32)
        bad.spheres["material"][3] = 9999
        with pytest.raises(capi.RtError):
            ctx.upload(bad)
    with pytest.raises(capi.RtError):
        capi.Context(profile=7)
    with pytest.raises(capi.RtError):
        capi.Context(profile=0, device=99)


def test_full_size_properties():
    """BASELINE config 2 size (1200x800): size-independent properties — every pixel gets
    exactly spp samples, sky rows are pure gradient, energy is bounded, paths*~2.66 = segments."""
    W, H, spp = 1200, 800, 8
    with capi.Context(profile=0, seed=1984) as ctx:
        ctx.upload(scenes.weekend(W, H)).build_accel(1)
        ctx.render(W, H, spp)
        a = ctx.accum()
        st = ctx.stats()
    assert np.all(a[..., 3] == spp)
    mean = a[..., :3] / spp
    assert np.all(np.isfinite(mean)) and mean.min() >= 0 and mean.max() <= 1.0 + 1e-5  # albedo<=1, sky<=1
    assert st["paths"] == W * H * spp
    assert 2.5 < st["segments"] / st["paths"] < 2.8  # 2.663 on the reference
    top = mean[-20:]  # top rows look at the sky: blue >= green >= red
    assert np.all(top[..., 2] >= top[..., 1] - 1e-6) and np.all(top[..., 1] >= top[..., 0] - 1e-6)


def test_gallery_image(golden_dir):
    """The reference's published 1200x800x500spp image (gallery/final.png, 16x16
    box-downsampled): a 32-spp GPU render of the same view, quantised and downsampled the
    same way, must look like it."""
    ref = np.load(golden_dir + "/gallery_final_75x50.npy").astype(np.float64)
    W, H, spp = 1200, 800, 32
    with capi.Context(profile=0, seed=1984) as ctx:
        ctx.upload(scenes.weekend(W, H)).build_accel(1)
        ctx.render(W, H, spp)
        _, rgb = ctx.resolve(want_linear=False)
    ds = rgb.astype(np.float64).reshape(50, 16, 75, 16, 3).mean((1, 3)) / 255.0
    psnr = SU.psnr(ds, ref)
    assert psnr > 38.0, psnr  # the unmodified CPU reference at 6 spp scores 45.5 dB
    assert np.all(np.abs(ds.mean((0, 1)) - ref.mean((0, 1))) < 0.01)


def test_render_into_torch_tensor_on_torch_stream():
    """rt_render_device on a caller stream / caller buffer (the multi-GPU plumbing path)
    produces the same bits as rt_render into the context's own frame."""
    import torch
    from a_dive_into_ray_tracing_b200.dist import render_frame
    W, H, spp = 64, 48, 16
    sc = scenes.weekend(W, H)
    with capi.Context(profile=0, seed=11) as ctx:
        ctx.upload(sc).build_accel(1)
        ctx.render(W, H, spp)
        ref = ctx.accum()
        acc = torch.zeros(H, W, 4, device="cuda", dtype=torch.float32)
        render_frame(ctx, W, H, spp, acc)                      # default (legacy) stream
        torch.cuda.synchronize()
        np.testing.assert_array_equal(acc.cpu().numpy(), ref)
        s = torch.cuda.Stream()
        acc2 = torch.zeros(H, W, 4, device="cuda", dtype=torch.float32)
        torch.cuda.synchronize()
        with torch.cuda.stream(s):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            render_frame(ctx, W, H, spp, acc2)
            e1.record()
        s.synchronize()
        assert e0.elapsed_time(e1) > 0.05  # the kernel really ran on THAT stream
        np.testing.assert_array_equal(acc2.cpu().numpy(), ref)


def test_large_scene_global_memory_path(l1_64):
    """~10^4 spheres: the scene no longer fits shared memory (k_render SMEM=0 variant, nodes
    through L1/L2), the sort runs multi-block. Closest hits: BVH == brute force == oracle;
    the rendered frame agrees with the oracle statistically."""
    W, H = 120, 80
    sc = scenes.sphere_field(100, width=W, height=H)
    assert sc.n_prims > 9000
    with capi.Context(profile=0, seed=3) as ctx:
        ctx.upload(sc).build_accel(1)
        nodes, leaf, big = ctx.accel()
        check_packed_bvh(sc, nodes, leaf, big)
        rays = D.primary_rays(sc.camera, W, H, 0, lens=(0.2, 0.1))
        ia, ta = ctx.trace_closest(rays, use_accel=True)
        ib, tb = ctx.trace_closest(rays, use_accel=False)
        np.testing.assert_array_equal(ia, ib)
        np.testing.assert_array_equal(ta, tb)
        io, to, _ = l1_64.closest_hit(sc, 0, rays)
        assert (ia != io).sum() <= 2
        same = (ia == io) & (io >= 0)
        assert (np.abs(ta[same] - to[same]) / to[same]).max() < 1e-5
        spp, K = 32, 4
        batches = []
        for k in range(K):
            ctx.clear()
            ctx.render(W, H, spp // K, spp_begin=k * (spp // K))
            a = ctx.accum().astype(np.float64)
            batches.append(a[..., :3] / a[..., 3:4])
        st = ctx.stats()
    assert st["smem_plan"] == 0 and st["smem_bytes"] < 60000  # only the accumulators: the scene stayed in global memory
    mu_a, var_mean_a = SU.batch_variance(batches)
    r, r2, nseg = l1_64.render_parallel(sc, 0, W, H, spp, seed=5, use_ref_bvh=True)
    mu_b, var_b = SU.mean_var(r, r2, spp)
    ok, d, b = SU.three_sigma_check(mu_a, var_mean_a, 1, mu_b, var_b, spp)
    assert ok, (d, b)
    assert abs(st["segments"] / st["paths"] - nseg / (W * H * spp)) < 0.05 * nseg / (W * H * spp)


@pytest.mark.parametrize("name", ["weekend", "next_week", "cornell_box", "next_week_final"])
def test_shared_memory_plans_render_the_same_image(name, monkeypatch):
    """k_render's four residency plans (0 global, 3 nodes only, 1 scene + one node copy, 2 scene + four
    quadrant copies) differ in where the data sits and in the node visiting order, never in a path's
    result: frames agree to fp32 summation order."""
    W, H, spp = 64, 48, 8
    sc = getattr(scenes, name)(W, H)
    frames, plans = [], []
    for cap in ("2", "1", "3", "0"):
        monkeypatch.setenv("B200RT_SMEM", cap)
        with capi.Context(profile=sc.profile, seed=11) as ctx:
            ctx.upload(sc).build_accel(1)
            ctx.render(W, H, spp)
            frames.append(ctx.accum())
            plans.append(ctx.stats()["smem_bytes"])
            if cap == "0":
                assert ctx.stats()["smem_plan"] == 0
    assert plans[-1] < 60000  # plan 0: only the accumulators
    if name != "next_week_final":  # (its 6 813 nodes do not fit: every cap ends in plan 0)
        assert plans[0] >= plans[1] >= plans[2] > plans[3]
    for f in frames[1:]:
        np.testing.assert_allclose(f, frames[0], rtol=1e-5, atol=1e-4)


def test_multi_primitive_leaves_give_identical_hits(hits_primary, monkeypatch):
    """Leaf collapsing (K primitives per leaf) changes the tree, never the answer."""
    sc = scenes.weekend(400, 225)
    ref = None
    for K in ("1", "2", "4", "8"):
        monkeypatch.setenv("B200RT_MAX_LEAF", K)
        with capi.Context(profile=0) as ctx:
            ctx.upload(sc).build_accel(1)
            nodes, leaf, big = ctx.accel()
            check_packed_bvh(sc, nodes, leaf, big, max_leaf=int(K))
            ids, ts = ctx.trace_closest(hits_primary["rays"], use_accel=True)
            ctx.render(96, 54, 8)
            img = ctx.accum()
        np.testing.assert_array_equal(ids, hits_primary["ids"])
        if ref is None:
            ref = (ts, img, len(nodes))
        else:
            np.testing.assert_array_equal(ts, ref[0])
            np.testing.assert_allclose(img, ref[1], rtol=1e-5, atol=1e-5)  # same paths; fp32 ties aside
            assert len(nodes) < ref[2]


def test_radix_sort_equals_bitonic_reference(monkeypatch):
    """The LSD radix sort of (Morton code, primitive) pairs yields exactly the order of the
    bitonic sort of the unique 64-bit keys: identical trees (single- and multi-block sizes)."""
    for sc in (scenes.weekend(64, 36), scenes.obj_room(width=32, height=32, subdivisions=3),
               scenes.sphere_field(100, width=64, height=36)):
        built = []
        for mode in ("radix", "bitonic"):
            monkeypatch.setenv("B200RT_SORT", mode)
            with capi.Context(profile=sc.profile) as ctx:
                ctx.upload(sc).build_accel(1)
                built.append(ctx.accel())
        (n0, l0_, b0), (n1, l1_, b1) = built
        np.testing.assert_array_equal(l0_, l1_)
        np.testing.assert_array_equal(b0, b1)
        assert n0.tobytes() == n1.tobytes()


def test_converged_image_vs_reference_gallery_matched_spp(golden_dir):
    """Config 2 at MATCHED spp (1200x800, 500 spp) against the reference's own published
    render of it (gallery/final.png, 8-bit, 4x4 box-downsampled fixture): PSNR >= 40 dB and
    within 1.5 dB of the Monte-Carlo floor (two renders of ours with different seeds),
    per-channel mean |delta| <= 3 sigma of the estimator (sigma from the two-seed spread)."""
    ref = np.load(golden_dir + "/gallery_final_300x200.npy").astype(np.float64) / 4.0 / 255.0
    W, H, spp = 1200, 800, 500
    imgs = []
    for seed in (1984, 7):
        with capi.Context(profile=0, seed=seed) as ctx:
            ctx.upload(scenes.weekend(W, H)).build_accel(1)
            ctx.render(W, H, spp)
            _, rgb = ctx.resolve(want_linear=False)
        imgs.append(rgb.astype(np.float64).reshape(200, 4, 300, 4, 3).mean((1, 3)) / 255.0)
    p_ref = SU.psnr(imgs[0], ref)
    p_floor = SU.psnr(imgs[0], imgs[1])
    assert p_ref >= 40.0, (p_ref, p_floor)
    assert p_ref >= p_floor - 1.5, (p_ref, p_floor)
    sigma = np.abs(imgs[0] - imgs[1]).reshape(-1, 3).mean(0)  # E|a-b| = 1.13 sigma_pixel for two equal estimators
    d = np.abs(imgs[0] - ref).reshape(-1, 3).mean(0)
    assert np.all(d <= 3.0 * sigma + 0.5 / 255.0), (d, sigma)  # + half an 8-bit level of quantisation
    assert np.all(np.abs(imgs[0].mean((0, 1)) - ref.mean((0, 1))) < 2.0 / 255.0)
