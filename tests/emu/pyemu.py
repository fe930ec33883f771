"""ctypes binding of tests/emu/libemu.so — TEST-ONLY host emulation of the device code
(see emu.cu). Never used by the product or by any reported number."""
import ctypes as C
import os
import subprocess

import numpy as np

from a_dive_into_ray_tracing_b200.ctypes_defs import BVH_NODE_DT, RtSceneDesc

HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(HERE, "libemu.so")
_vp = C.c_void_p


def build(force=False):
    src = os.path.join(HERE, "emu.cu")
    csrc = os.path.join(HERE, "..", "..", "a_dive_into_ray_tracing_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cuh", ".h"))]
    if force or not os.path.exists(PATH) or any(os.path.getmtime(d) > os.path.getmtime(PATH) for d in deps):
        subprocess.check_call(["nvcc", "-O2", "-std=c++17", "--expt-relaxed-constexpr", "-gencode",
                               "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-shared", "-o", PATH, src],
                              stderr=subprocess.DEVNULL)


class Emu:
    def __init__(self, scene, profile=None, quality=1, big_frac=0.30, big_rounds=3, shuffle=0, max_leaf=4):
        build()
        self.lib = L = C.CDLL(PATH)
        L.emu_create.restype = _vp
        L.emu_create.argtypes = [C.POINTER(RtSceneDesc), C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_char_p,
                                 C.c_int]
        L.emu_destroy.argtypes = [_vp]
        L.emu_counts.argtypes = [_vp, _vp, _vp, _vp]
        L.emu_download.argtypes = [_vp, _vp, _vp, _vp]
        L.emu_trace.argtypes = [_vp, _vp, C.c_int, C.c_float, C.c_float, C.c_int, _vp, _vp, _vp]
        L.emu_trace_inst.argtypes = [_vp, _vp, C.c_int, C.c_float, C.c_float, C.c_int, _vp, _vp, _vp]
        L.emu_render.argtypes = [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_int, _vp, _vp, _vp]
        self.scene = scene
        self.profile = scene.profile if profile is None else profile
        d = scene.desc()
        err = C.create_string_buffer(256)
        self.h = L.emu_create(C.byref(d), self.profile, quality, big_frac, big_rounds, shuffle, max_leaf, err, 256)
        if not self.h:
            raise ValueError(err.value.decode())

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.emu_destroy(self.h)
            self.h = None

    def accel(self):
        nn, nl, nb = C.c_int(), C.c_int(), C.c_int()
        self.lib.emu_counts(self.h, C.addressof(nn), C.addressof(nl), C.addressof(nb))
        nodes = np.zeros(max(nn.value, 1), BVH_NODE_DT)
        leaf = np.zeros(max(nl.value, 1), np.int32)
        big = np.zeros(max(nb.value, 1), np.int32)
        self.lib.emu_download(self.h, nodes.ctypes.data, leaf.ctypes.data, big.ctypes.data)
        return nodes[:nn.value], leaf[:nl.value], big[:nb.value]

    def trace(self, rays, t_min=1e-3, t_max=np.inf, use_accel=1):
        rays = np.ascontiguousarray(rays, np.float32)
        n = len(rays)
        ids = np.empty(n, np.int32)
        ts = np.empty(n, np.float32)
        cnt = np.zeros(2, np.uint64)
        self.lib.emu_trace(self.h, rays.ctypes.data, n, t_min, t_max, use_accel, ids.ctypes.data, ts.ctypes.data,
                           cnt.ctypes.data)
        return ids, ts, cnt

    def trace_inst(self, rays, t_min=1e-3, t_max=np.inf, use_accel=1):
        """two-level scenes: (prim ids, instance ids, t)"""
        rays = np.ascontiguousarray(rays, np.float32)
        n = len(rays)
        ids, inst, ts = np.empty(n, np.int32), np.empty(n, np.int32), np.empty(n, np.float32)
        self.lib.emu_trace_inst(self.h, rays.ctypes.data, n, t_min, t_max, use_accel, ids.ctypes.data, inst.ctypes.data,
                                ts.ctypes.data)
        return ids, inst, ts

    def texture(self, material, prim, p, outward):
        """material colour at hit points p of primitives `prim` with outward normals (device code)"""
        p = np.ascontiguousarray(p, np.float32).reshape(-1, 3)
        outward = np.ascontiguousarray(outward, np.float32).reshape(-1, 3)
        prim = np.ascontiguousarray(np.broadcast_to(np.asarray(prim, np.int32), (len(p),)))
        out = np.zeros((len(p), 3), np.float32)
        self.lib.emu_texture.argtypes = [_vp, C.c_int, C.c_int, _vp, _vp, _vp, _vp]
        self.lib.emu_texture(self.h, material, len(p), prim.ctypes.data, p.ctypes.data, outward.ctypes.data,
                             out.ctypes.data)
        return out

    def aov(self, W, H, spp=1, seed=1984):
        out = np.zeros((H, W, 8), np.float32)
        self.lib.emu_aov.argtypes = [_vp, C.c_int, C.c_int, C.c_int, C.c_uint64, _vp]
        self.lib.emu_aov(self.h, W, H, spp, seed, out.ctypes.data)
        return out

    def render(self, W, H, spp, spp_begin=0, seed=1984, rows=None, threads=None):
        from concurrent.futures import ThreadPoolExecutor
        s = np.zeros((H, W, 3), np.float64)
        s2 = np.zeros((H, W, 3), np.float64)
        j0, j1 = rows if rows is not None else (0, H)
        threads = threads or os.cpu_count() or 1
        edges = np.linspace(j0, j1, min(j1 - j0, threads * 4) + 1).astype(int)
        stats = np.zeros((len(edges) - 1, 4), np.uint64)

        def job(k):
            self.lib.emu_render(self.h, W, H, spp_begin, spp, seed, int(edges[k]), int(edges[k + 1]), s.ctypes.data,
                                s2.ctypes.data, stats[k].ctypes.data)

        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(job, range(len(edges) - 1)))
        return s, s2, stats.sum(0)
