"""CPU model of k_render's BVH step in the shared-memory plans (render_kernels.cuh, RT_HALF_PLANES): node planes as fp16
numbers v = q / 23 rounded outwards from the padded 16-bit coordinates q, the ray's slope per axis as an fp16 pair
rounded apart (entry: towards -inf, exit: towards +inf), t = v * S + C evaluated as FHFMA does (fp16 x fp16 product exact in
fp32, one rounding after the add), min / max dropping NaN. The model walks the threaded tree of the host-emulated builder
and must name exactly the primitive brute force names - on the reference's golden ray sets and on the edge cases of the
encoding (zero direction components: the slope saturates to max-finite / infinity and 0 * inf = NaN; origins on planes;
direction scales 1e-4 .. 1e4). It pins the ARITHMETIC of the encoding without a GPU; the CUDA code itself is checked
against brute force in tests/test_render_kernel_traversal.py."""
import numpy as np
import pytest

from a_dive_into_ray_tracing_b200 import scenes
from tests.emu.pyemu import Emu

f32, f16 = np.float32, np.float16
Q_MAX, KINV = 46335, f32(23.0)


def _rd16(x):
    """float32 -> fp16 towards -inf (cvt.rm.f16.f32): overflow gives max-finite for positive, -inf for negative values"""
    with np.errstate(over="ignore"):
        h = x.astype(f16)
    h = np.where(np.isposinf(h) & np.isfinite(x), f16(65504.0), h).astype(f16)
    bad = h.astype(f32) > x
    h[bad] = np.nextafter(h[bad], f16(-np.inf))
    return h


def _ru16(x):
    return (-_rd16(-x)).astype(f16)


def _fhfma(v, s, c):
    """fma.rn.f32.f16: the product of two fp16 numbers is exact in fp32 (22 significant bits), then one rounding"""
    with np.errstate(invalid="ignore", over="ignore"):
        return (v.astype(np.float64) * s.astype(np.float64) + c.astype(np.float64)).astype(f32)


def _stage(nodes):
    """the staging loop: 16-bit coordinates over the root box with two steps of padding, then fp16 outwards"""
    lo, hi = nodes["bmin"][0].astype(f32), nodes["bmax"][0].astype(f32)
    ext = np.maximum(hi - lo, f32(1e-6) * np.maximum(np.maximum(np.abs(lo), np.abs(hi)), f32(1e-30))).astype(f32)
    fs = (ext * f32(1.0 / (Q_MAX - 8))).astype(f32)
    fi = (f32(1.0) / fs).astype(f32)
    ql = np.clip(np.floor((nodes["bmin"] - lo) * fi) - 2, 0, Q_MAX).astype(f32)
    qh = np.clip(np.ceil((nodes["bmax"] - lo) * fi) + 2, 0, Q_MAX).astype(f32)
    k = f32(1.0) / KINV
    # __fmul_rd / __fmul_ru: the float64 product is exact, round it to float32 in the right direction
    pl, ph = ql.astype(np.float64) * np.float64(k), qh.astype(np.float64) * np.float64(k)
    vl32 = pl.astype(f32); vl32 = np.where(vl32 > pl, np.nextafter(vl32, f32(-np.inf)), vl32).astype(f32)
    vh32 = ph.astype(f32); vh32 = np.where(vh32 < ph, np.nextafter(vh32, f32(np.inf)), vh32).astype(f32)
    return _rd16(vl32), _ru16(vh32), lo, fs


def _walk(sc, nodes, leaf, big, rays, t_min=f32(1e-3)):
    nn = len(nodes)
    vl, vh, fb, fs = _stage(nodes)
    esc, pay = nodes["escape"], nodes["payload"]
    o, d = rays[:, 0:3].astype(f32), rays[:, 4:7].astype(f32)
    safe = np.where(np.abs(d) < 1e-20, np.copysign(f32(1e-20), d), d).astype(f32)
    inv_d = (f32(1) / safe).astype(f32)
    neg = np.signbit(inv_d)
    sl = (KINV * (fs * inv_d).astype(f32)).astype(f32)
    s_e, s_x = _rd16(sl), _ru16(sl)
    c = ((fb - o) * inv_d).astype(f32)
    n = len(rays)
    ht, hid = np.full(n, np.inf, f32), np.full(n, -1, np.int64)
    cen, rad = sc.spheres["center0"].astype(np.float64), sc.spheres["radius"].astype(np.float64)

    def test_sphere(idx, sid):
        oc = o[idx].astype(np.float64) - cen[sid]
        dd = d[idx].astype(np.float64)
        a, hb, cc = (dd * dd).sum(1), (oc * dd).sum(1), (oc * oc).sum(1) - rad[sid] ** 2
        disc = hb * hb - a * cc
        sq = np.sqrt(np.maximum(disc, 0))
        ok = disc > 0
        for root in ((-hb - sq) / a, (-hb + sq) / a):
            take = ok & (root > t_min) & (root < ht[idx])
            ht[idx[take]] = root[take].astype(f32)
            hid[idx[take]] = sid[take]
            ok &= ~take

    for b in big:
        test_sphere(np.arange(n), np.full(n, b))
    node = np.zeros(n, np.int64)  # record index; nn = finished
    steps = 0
    while True:
        act = np.nonzero(node < nn)[0]
        if len(act) == 0:
            break
        k = node[act]
        e = np.where(neg[act], vh[k], vl[k])
        x = np.where(neg[act], vl[k], vh[k])
        with np.errstate(invalid="ignore"):
            te, tx = _fhfma(e, s_e[act], c[act]), _fhfma(x, s_x[act], c[act])
            tn = np.fmax(np.fmax(np.fmax(te[:, 0], te[:, 1]), te[:, 2]), t_min)      # fmax / fmin drop NaN
            tf = np.fmin(np.fmin(np.fmin(tx[:, 0], tx[:, 1]), tx[:, 2]), ht[act])
        hit = tn <= tf
        is_leaf = pay[k] < 0
        lh = hit & is_leaf
        if lh.any():
            test_sphere(act[lh], leaf[(~pay[k[lh]]) >> 3].astype(np.int64))
        node[act] = np.where(hit & ~is_leaf, k + 1, np.where(is_leaf, k + 1, esc[k]))
        steps += len(act)
        assert steps < 400 * n
    return hid, steps / n


def _brute(sc, rays, t_min=1e-3):
    o, d = rays[:, 0:3].astype(np.float64), rays[:, 4:7].astype(np.float64)
    d = np.where(np.abs(d) < 1e-20, np.copysign(1e-20, d), d)
    best, bid = np.full(len(rays), np.inf), np.full(len(rays), -1)
    cen, rad = sc.spheres["center0"].astype(np.float64), sc.spheres["radius"].astype(np.float64)
    for s in range(len(cen)):
        oc = o - cen[s]
        a, hb, cc = (d * d).sum(1), (oc * d).sum(1), (oc * oc).sum(1) - rad[s] ** 2
        disc = hb * hb - a * cc
        sq = np.sqrt(np.maximum(disc, 0))
        for root in ((-hb - sq) / a, (-hb + sq) / a):
            ok = (disc > 0) & (root > t_min) & (root < best.astype(f32))
            best[ok], bid[ok] = root[ok], s
    return bid


@pytest.fixture(scope="module")
def weekend_tree():
    sc = scenes.weekend(400, 225)
    nodes, leaf, big = Emu(sc, max_leaf=1).accel()
    assert np.all(nodes["payload"][nodes["payload"] >= 0] == np.nonzero(nodes["payload"] >= 0)[0] + 1)  # first child = next record
    return sc, nodes, leaf, big


def test_fp16_planes_are_conservative(weekend_tree):
    sc, nodes, leaf, big = weekend_tree
    vl, vh, fb, fs = _stage(nodes)
    x_lo = fb + vl.astype(np.float64) * 23.0 * fs
    x_hi = fb + vh.astype(np.float64) * 23.0 * fs
    pad = 1.5 * fs  # the staging pads by two 16-bit steps; fp16 may only move a plane further out
    assert np.all(x_lo <= nodes["bmin"]) and np.all(x_lo[vl > 0] <= (nodes["bmin"] - pad)[vl > 0])  # (v = 0: the root's own plane)
    assert np.all(x_hi >= nodes["bmax"] + pad)
    assert float(vh.max()) < 2048 and float(vl.min()) >= 0.0


@pytest.mark.parametrize("which", ["weekend_hits_c1", "weekend_hits_bounce"])
def test_model_names_the_brute_force_primitive_on_golden_rays(weekend_tree, golden_dir, which):
    import os
    sc, nodes, leaf, big = weekend_tree
    g = np.load(os.path.join(golden_dir, which + ".npz"))
    rays = g["rays"][:6000]
    hid, steps = _walk(sc, nodes, leaf, big, rays)
    np.testing.assert_array_equal(hid, g["ids"][:6000])
    assert steps < 30


def test_model_edge_cases(weekend_tree):
    sc, nodes, leaf, big = weekend_tree
    rng = np.random.default_rng(5)
    n = 4000
    o = np.array([-11.0, 0.0, -11.0]) + np.array([22.0, 1.5, 22.0]) * rng.random((n, 3))
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:3], rays[:, 4:7] = o, d
    k = np.arange(n)
    rays[k % 8 == 0, 4] = 0.0                 # slope +-1e20 -> max-finite / infinity, 0 * inf = NaN on v = 0 planes
    rays[k % 8 == 1, 5] = -0.0
    rays[k % 8 == 2, 4:6] = 0.0
    rays[k % 8 == 3, 6] = 1e-30
    rays[k % 8 == 4, 4:7] *= 1e-4             # slopes beyond the fp16 range on every axis
    rays[k % 8 == 5, 4:7] *= 1e4              # slopes in the fp16 denormals
    rays[k % 8 == 6, 0:3] = np.round(rays[k % 8 == 6, 0:3] * 2) / 2   # origins on a coarse grid
    hid, _ = _walk(sc, nodes, leaf, big, rays)
    bid = _brute(sc, rays)
    mism = hid != bid
    assert (bid >= 0).mean() > 0.3
    # the model's sphere test is the double-precision one of the brute force: ids must agree exactly
    assert not mism.any(), (int(mism.sum()), hid[mism][:5], bid[mism][:5])
