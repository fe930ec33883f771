"""rt_reduce / rt_comm_init_all (C ABI): the per-GPU accumulation frames of a sample split,
summed onto rank 0 by the library's own ncclReduce. Needs two CUDA devices (skipped otherwise); the CPU side
(symbols exported, error behaviour without a communicator) is in test_capi_cpu.py / below."""
import ctypes as C
import threading

import numpy as np
import pytest

from a_dive_into_ray_tracing_b200 import capi, scenes
from a_dive_into_ray_tracing_b200.dist import sample_range


def test_comm_entry_points_are_exported():
    lib = capi.load_library()
    for s in ("rt_comm_unique_id", "rt_comm_init", "rt_comm_init_all", "rt_reduce", "rt_host_alloc", "rt_host_free"):
        assert hasattr(lib, s)
    assert lib.rt_reduce(None, 8, 8, None, 0, 0, None) == 1  # RT_ERR_INVALID: null context
    assert lib.rt_comm_init(None, None, 0, 1) == 1


@pytest.mark.gpu
@pytest.mark.parametrize("uniform", [True, False])
def test_reduce_two_gpus_equals_single_gpu(uniform):
    """frame of a 2-GPU sample split after rt_reduce == the single-GPU frame to 1e-6 relative (fp32 summation
    order is the only difference: Philox counters are (pixel, GLOBAL sample, event)); both travelling formats
    (R,G,B + count word / all four lanes)"""
    if capi.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    W, H, spp = 160, 96, 48
    sc = scenes.weekend(W, H)
    with capi.Context(profile=0, device=0, seed=11) as ref:
        ref.upload(sc).build_accel(1)
        ref.render(W, H, spp)
        want = ref.accum().astype(np.float64)
    ctxs = [capi.Context(profile=0, device=g, seed=11) for g in range(2)]
    try:
        for c in ctxs:
            c.upload(sc).build_accel(1)
        handles = (C.c_void_p * 2)(*[c.h for c in ctxs])
        assert ctxs[0].lib.rt_comm_init_all(handles, 2) == 0, ctxs[0].lib.rt_last_error(ctxs[0].h)
        errs = []

        def job(g):
            try:
                b, n = sample_range(spp, g, 2)
                ctxs[g].render(W, H, n, b)
                ctxs[g].reduce(W, H, None, root=0, uniform_count=uniform)
                ctxs[g].sync()
            except Exception as e:  # pragma: no cover
                errs.append(e)

        th = [threading.Thread(target=job, args=(g,)) for g in range(2)]
        [t.start() for t in th]
        [t.join() for t in th]
        assert not errs, errs
        got = ctxs[0].accum().astype(np.float64)
        other = ctxs[1].accum().astype(np.float64)
    finally:
        for c in ctxs:
            c.close()
    assert np.all(got[..., 3] == spp) and np.all(other[..., 3] == spp - spp // 2)  # non-root frames stay as rendered
    rel = np.abs(got[..., :3] - want[..., :3]) / np.maximum(np.abs(want[..., :3]), 1e-3 * spp)
    assert rel.max() < 1e-6 * 8, rel.max()


@pytest.mark.gpu
def test_reduce_single_rank_is_a_no_op_and_pinned_resolve():
    W, H = 64, 40
    sc = scenes.weekend(W, H)
    with capi.Context(profile=0, seed=3) as ctx:
        ctx.upload(sc).build_accel(1)
        ctx.render(W, H, 8)
        a = ctx.accum()
        ctx.reduce(W, H)  # no communicator, one rank: nothing to do
        assert np.array_equal(a, ctx.accum())
        # resolve into page-locked memory (rt_host_alloc) == resolve into an ordinary array
        _, rgb = ctx.resolve(want_linear=False)
        pin = ctx.pinned_array((H, W, 3), np.uint8)
        pin[:] = 0
        ctx.resolve_device(W, H, ctx.lib.rt_accum_device_ptr(ctx.h), out_rgb8=pin)
        assert np.array_equal(pin, rgb)
