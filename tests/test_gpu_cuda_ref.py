"""GPU parity (-m gpu) against what the reference's own CUDA-tree device code computed on a B200
(tests/golden/cuda_ref_*.npz, see tests/test_cuda_ref_pinning.py): closest hits of moving spheres,
triangles (incl. the |N.d| < 0.01 reject and the back-face flip) and rects through the C ABI - brute force,
the plain BVH loop and the render kernel's own traversal.

Tolerances: primitive ids identical to the reference's on every ray where the reference's float result is
robust (= the double restatement of the same formulas, itself pinned bit for bit in its float build, names
the same object); t within 1e-5 relative of that double restatement (+1e-6 of the origin's magnitude in
hit-point position); and within 5e-3 of the reference's own float t, which is all its b*b - a*c sphere
formula supports (it moves by 2e-3 between its fma and non-fma builds)."""
import numpy as np
import pytest

from a_dive_into_ray_tracing_b200 import capi
from tests.cuda_ref_util import RefCase

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["nw_spheres", "nw_media", "tri_room"])
def test_gpu_closest_hit_vs_reference_cuda_vectors(name, l1_64):
    c = RefCase(name)
    sc = c.scene
    ref = c.out("list")
    want_ids = c.prim_ids(ref["obj"])
    h64 = l1_64.pin_hits(sc, sc.profile, c.rays, c.t_min)
    robust = h64["obj"] == ref["obj"]
    assert robust.mean() > 0.995, robust.mean()
    r64 = c.rays.astype(np.float64)
    tol = 1e-5 * np.abs(h64["t"].astype(np.float64)) + 1e-6 * np.linalg.norm(r64[:, :3], axis=1) / np.linalg.norm(r64[:, 4:7], axis=1)
    with capi.Context(profile=sc.profile, seed=7) as ctx:
        ctx.upload(sc).build_accel(1)
        for mode in (0, 1, 2):
            ids, ts = ctx.trace_closest(c.rays, t_min=c.t_min, use_accel=mode)
            mism = (ids != want_ids) & robust
            assert mism.mean() <= 1e-3, (mode, int(mism.sum()))
            k = robust & (ref["hit"] == 1) & (ids == want_ids)
            err = np.abs(ts[k].astype(np.float64) - h64["t"][k])
            assert np.all(err <= tol[k]), (mode, (err / tol[k]).max())
            rel_ref = np.abs(ts[k] - ref["t"][k]) / ref["t"][k]
            assert rel_ref.max() < 5e-3 and np.median(rel_ref) < 1e-5, (mode, rel_ref.max())
