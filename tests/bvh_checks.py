"""Structural checks of a packed threaded BVH (rt_bvh_node array + big-primitive list),
shared by the emulation tests (CPU) and the GPU tests."""
import numpy as np

from a_dive_into_ray_tracing_b200 import ctypes_defs as D


def prim_boxes(sc, thickness):
    """Reference bounding boxes (sphere.h:79-84, moving_sphere.h:74-82, triangle.h:74-100,
    aarect.h:26-31) as float64 arrays keyed by RT_PRIM_ID."""
    boxes = {}
    for i, s in enumerate(sc.spheres):
        r = abs(float(s["radius"]))
        c0 = s["center0"].astype(np.float64)
        c1 = s["center1"].astype(np.float64) if s["moving"] else c0
        boxes[int(D.prim_id(0, i))] = (np.minimum(c0, c1) - r, np.maximum(c0, c1) + r)
    for i, t in enumerate(sc.triangles):
        v = np.stack([t["v0"], t["v1"], t["v2"]]).astype(np.float64)
        lo, hi = v.min(0), v.max(0)
        flat = np.abs(lo - hi) < thickness
        lo = np.where(flat, lo - thickness, lo)
        hi = np.where(flat, hi + thickness, hi)
        boxes[int(D.prim_id(1, i))] = (lo, hi)
    for i, q in enumerate(sc.quads):
        ax = int(q["axis"])
        ia, ib = (1 if ax == 0 else 0), (1 if ax == 2 else 2)
        lo, hi = np.zeros(3), np.zeros(3)
        lo[ax], hi[ax] = q["k"] - thickness, q["k"] + thickness
        lo[ia], hi[ia] = q["a0"], q["a1"]
        lo[ib], hi[ib] = q["b0"], q["b1"]
        boxes[int(D.prim_id(2, i))] = (lo, hi)
    return boxes


def check_packed_bvh(sc, nodes, big):
    n = len(nodes)
    thickness = 0.01 if (sc.flags & D.RT_FLAG_FLIP_NORMALS) else 0.1
    boxes = prim_boxes(sc, thickness)
    all_ids = set(boxes)
    big_ids = set(int(b) for b in big)
    assert big_ids <= all_ids
    n_small = len(all_ids) - len(big_ids)
    assert n == (2 * n_small - 1 if n_small else 0)
    if n == 0:
        return
    esc, pay = nodes["escape"], nodes["payload"]
    leaves = pay < 0
    leaf_ids = [int(~p) for p in pay[leaves]]
    # every non-big primitive in exactly one leaf
    assert sorted(leaf_ids) == sorted(all_ids - big_ids)
    # threaded layout: depth-first preorder
    assert np.all(esc > np.arange(n)) and np.all(esc <= n) and esc[0] == n
    inner = ~leaves
    assert np.all(pay[inner] == np.arange(n)[inner] + 1)  # first child follows its parent
    assert np.all(esc[leaves] == np.arange(n)[leaves] + 1)
    lo, hi = nodes["bmin"].astype(np.float64), nodes["bmax"].astype(np.float64)
    assert np.all(lo <= hi)
    for i in np.where(inner)[0]:
        l = i + 1
        r = esc[l]
        assert r < n and esc[r] == esc[i]  # the two children tile the parent's range
        assert np.all(lo[i] <= np.minimum(lo[l], lo[r]) + 1e-12) and np.all(hi[i] >= np.maximum(hi[l], hi[r]) - 1e-12)
    for i in np.where(leaves)[0]:
        blo, bhi = boxes[int(~pay[i])]
        assert np.all(lo[i] <= blo) and np.all(hi[i] >= bhi)  # node box contains the primitive box


def sah_cost(nodes):
    """Surface-area-heuristic cost of the packed tree (inner nodes only, unit costs)."""
    lo, hi = nodes["bmin"].astype(np.float64), nodes["bmax"].astype(np.float64)
    e = hi - lo
    area = e[:, 0] * e[:, 1] + e[:, 1] * e[:, 2] + e[:, 2] * e[:, 0]
    return float(area.sum() / area[0])
