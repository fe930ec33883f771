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


def check_packed_bvh(sc, nodes, leaf_prims, big, max_leaf=8):
    n = len(nodes)
    thickness = 0.01 if (sc.flags & D.RT_FLAG_FLIP_NORMALS) else 0.1
    boxes = prim_boxes(sc, thickness)
    all_ids = set(boxes)
    big_ids = set(int(b) for b in big)
    assert big_ids <= all_ids and len(big_ids) == len(big)
    n_small = len(all_ids) - len(big_ids)
    # every non-big primitive appears exactly once in the leaf order
    assert sorted(int(x) for x in leaf_prims) == sorted(all_ids - big_ids)
    if n_small == 0:
        assert n == 0
        return
    esc, pay = nodes["escape"], nodes["payload"]
    leaves = pay < 0
    enc = ~pay[leaves]
    first, count = enc >> 3, (enc & 7) + 1
    assert np.all(count >= 1) and np.all(count <= max_leaf)
    # leaves tile leaf_prims (canonical leaf order; the nodes themselves are laid out in the
    # quadrant's front-to-back visiting order): each primitive in exactly one leaf
    order = np.argsort(first)
    assert np.array_equal(first[order], np.concatenate([[0], np.cumsum(count[order])[:-1]]))
    assert int(count.sum()) == n_small
    n_leaves = int(leaves.sum())
    assert n == 2 * n_leaves - 1  # full binary tree
    lo, hi = nodes["bmin"].astype(np.float64), nodes["bmax"].astype(np.float64)
    assert np.all(lo <= hi)
    inner = ~leaves
    assert np.all((esc >= 0) & (esc <= n)) and np.all((pay[inner] >= 0) & (pay[inner] < n))
    # threaded links, in ANY storage order (depth-first as built, or with the top of the tree moved
    # to the front): the root is node 0 and escapes to n; an inner node's first child is its payload,
    # the second child is the first child's escape, and the second child escapes where its parent does
    assert esc[0] == n
    seen = np.zeros(n, bool)
    stack = [0]
    while stack:
        i = stack.pop()
        assert not seen[i]
        seen[i] = True
        if leaves[i]:
            continue
        l = int(pay[i])
        r = int(esc[l])
        assert r < n and esc[r] == esc[i]
        assert np.all(lo[i] <= np.minimum(lo[l], lo[r]) + 1e-12) and np.all(hi[i] >= np.maximum(hi[l], hi[r]) - 1e-12)
        stack += [l, r]
    assert seen.all()  # every node is reachable exactly once
    for i, f, c in zip(np.where(leaves)[0], first, count):
        for pid in leaf_prims[f:f + c]:
            blo, bhi = boxes[int(pid)]
            assert np.all(lo[i] <= blo) and np.all(hi[i] >= bhi)  # node box contains the primitive boxes


def sah_cost(nodes):
    """Surface-area-heuristic cost of the packed tree (inner nodes only, unit costs)."""
    lo, hi = nodes["bmin"].astype(np.float64), nodes["bmax"].astype(np.float64)
    e = hi - lo
    area = e[:, 0] * e[:, 1] + e[:, 1] * e[:, 2] + e[:, 2] * e[:, 0]
    return float(area.sum() / area[0])
