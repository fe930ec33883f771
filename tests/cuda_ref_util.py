"""Loader for the reference-held golden vectors of the CUDA trees (tests/golden/cuda_ref_*.npz, produced by
tools/make_cuda_golden.py from oracle/ref_cuda_harness.cu = the reference's unmodified device code run on a
B200). Rebuilds the flattened scene exactly as the harness saw it, with the perlin tables and the rotate_y
sines / cosines the REFERENCE computed."""
import os

import numpy as np

from a_dive_into_ray_tracing_b200 import ctypes_defs as D

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["nw_spheres", "nw_cornell_inst", "nw_media", "tri_room", "tri_room_inst"]


class RefCase:
    def __init__(self, name, variant="nofma"):
        g = dict(np.load(os.path.join(GOLDEN, "cuda_ref_%s.npz" % name)))
        self.name, self.g, self.variant = name, g, variant
        self.tree, self.seed, self.t_min = int(g["tree"]), int(g["seed"]), float(g["t_min"])
        self.rays, self.n_color = g["rays"], int(g["n_color"])
        perlin = np.zeros(len(g["nofma_perlin"]), D.PERLIN_DT)
        for k, tab in enumerate(g["nofma_perlin"]):  # identical in both builds (integer / copy operations only)
            perlin[k]["ranvec"] = tab[:768].reshape(256, 3)
            perlin[k]["perm_x"], perlin[k]["perm_y"], perlin[k]["perm_z"] = tab[768:1024], tab[1024:1280], tab[1280:1536]
        mats = g["materials"].copy()
        k = 0
        for m in mats:  # noise textures -> their table, in creation order
            if m["texture"] == D.RT_TEX_NOISE:
                m["albedo2"][1] = k
                k += 1
        media = g["media"].copy()
        for k in range(len(media)):
            media[k]["sin_y"], media[k]["cos_y"] = self.out("med_sc")[k]
        self.scene = D.Scene(spheres=g["spheres"], triangles=g["triangles"], quads=g["quads"], materials=mats,
                             media=media, perlin=perlin, images=[g["image%d" % i] for i in range(int(g["n_images"]))],
                             background=tuple(g["background"]), sky_gradient=0, t_min=self.t_min,
                             max_depth=int(g["max_depth"]), flags=int(g["flags"]), profile=int(g["profile"]), name=name)
        from oracle.pyoracle import L1
        self.inst = None
        if g["inst"]["flag"].any():
            self.inst = np.zeros(len(g["inst"]), L1.INST_DT)
            self.inst["flag"] = g["inst"]["flag"]
            self.inst["offset"] = g["inst"]["off"]
            self.inst["sin_y"], self.inst["cos_y"] = self.out("inst_sc")[:, 0], self.out("inst_sc")[:, 1]

    def out(self, key, variant=None):
        return self.g["%s_%s" % (variant or self.variant, key)]

    def prim_ids(self, obj):
        """flat list index (spheres, triangles, quads[, media]) -> RT_PRIM_ID, -1 stays -1"""
        sc = self.scene
        ns, nt, nq = len(sc.spheres), len(sc.triangles), len(sc.quads)
        obj = np.asarray(obj)
        ids = np.full(obj.shape, -1, np.int32)
        for lo, hi, ty in ((0, ns, D.RT_PRIM_SPHERE), (ns, ns + nt, D.RT_PRIM_TRIANGLE), (ns + nt, ns + nt + nq, D.RT_PRIM_QUAD),
                           (ns + nt + nq, 1 << 30, 3)):
            m = (obj >= lo) & (obj < hi)
            ids[m] = (ty << 28) | (obj[m] - lo)
        return ids
