"""oracle/pyoracle.py — ctypes bindings of the two CPU checkers. TEST INFRASTRUCTURE.

  L1  oracle/liboracle.so        plain-C restatement (rt_oracle.c), orc64_* / orc32_*
  L0  oracle/_ref/libref_l0.so   the unmodified reference CPU renderer (ref_harness.cpp)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module. The product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from a_dive_into_ray_tracing_b200.ctypes_defs import BVH_NODE_DT, RtSceneDesc

HERE = os.path.dirname(os.path.abspath(__file__))
L1_PATH = os.path.join(HERE, "liboracle.so")
L0_PATH = os.path.join(HERE, "_ref", "libref_l0.so")
L0_ASIS_PATH = os.path.join(HERE, "_ref", "libref_l0_asis.so")

_vp = C.c_void_p


def build(force=False):
    """make -C oracle (liboracle.so always; _ref only if /root/reference exists)."""
    if force or not os.path.exists(L1_PATH) or (
            os.path.exists("/root/reference") and not os.path.exists(L0_PATH)):
        subprocess.check_call(["make", "-C", HERE, "-s"] + (["-B"] if force else []))


class L1:
    """Plain-C restatement. prec = 64 (CPU renderer arithmetic) or 32 (CUDA trees)."""

    def __init__(self, prec=64):
        if not os.path.exists(L1_PATH):
            build()
        self.lib = C.CDLL(L1_PATH)
        self.prec = prec
        p = "orc%d_" % prec
        self._closest = getattr(self.lib, p + "closest_hit")
        self._closest.restype = C.c_int
        self._closest.argtypes = [C.POINTER(RtSceneDesc), C.c_int, _vp, C.c_int, C.c_double, C.c_double, C.c_int,
                                  C.c_uint64, _vp, _vp, _vp]
        self._packed = getattr(self.lib, p + "closest_hit_packed")
        self._packed.restype = C.c_int
        self._packed.argtypes = [C.POINTER(RtSceneDesc), C.c_int, _vp, C.c_int, _vp, _vp, C.c_int, _vp, C.c_int,
                                 C.c_double, C.c_double, _vp, _vp, _vp]
        self._render = getattr(self.lib, p + "render")
        self._render.restype = C.c_uint64
        self._render.argtypes = [C.POINTER(RtSceneDesc), C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int,
                                 C.c_int, C.c_int, C.c_int, _vp, _vp]
        self._scatter = getattr(self.lib, p + "scatter_one")
        self._scatter.restype = C.c_int
        self._scatter.argtypes = [C.POINTER(RtSceneDesc), C.c_int, C.c_int, _vp, _vp, C.c_uint64, C.c_int, _vp, _vp]
        self._quant = getattr(self.lib, p + "quantise")
        self._quant.restype = None
        self._quant.argtypes = [C.c_int, _vp, C.c_int, _vp]
        self._tex = getattr(self.lib, p + "tex_value_at")
        self._tex.restype = None
        self._tex.argtypes = [C.POINTER(RtSceneDesc), C.c_int, C.c_int, _vp, _vp]
        self._suv = getattr(self.lib, p + "sphere_uv_at")
        self._suv.restype = None
        self._suv.argtypes = [C.c_int, _vp, _vp]
        self._mdemo = getattr(self.lib, p + "medium_reference_order_demo")
        self._mdemo.restype = C.c_double
        self._mdemo.argtypes = [C.POINTER(RtSceneDesc), _vp, C.c_int, C.c_int, C.c_uint64]
        self.lib.orc_srand.argtypes = [C.c_uint]
        # pinning against the reference's CUDA-tree device code (oracle/ref_cuda_harness.cu records)
        for name, args in (("pin_hits", [C.POINTER(RtSceneDesc), C.c_int, _vp, _vp, C.c_int, C.c_double, C.c_double, _vp]),
                           ("pin_scatter", [C.POINTER(RtSceneDesc), C.c_int, _vp, _vp, C.c_int, C.c_double, C.c_double,
                                            C.c_uint64, _vp]),
                           ("pin_media", [C.POINTER(RtSceneDesc), _vp, C.c_int, C.c_uint64, _vp]),
                           ("pin_color", [C.POINTER(RtSceneDesc), C.c_int, _vp, _vp, C.c_int, C.c_uint64, _vp]),
                           ("xorwow_kat", [C.c_uint64, C.c_int, _vp])):
            f = getattr(self.lib, p + name)
            f.restype = None
            f.argtypes = args
            setattr(self, "_" + name, f)

    # records of oracle/ref_cuda_harness.cu (HHit, HScatter, HMediumHit) and orc_inst
    PIN_HIT_DT = np.dtype([("hit", "<i4"), ("obj", "<i4"), ("t", "<f4"), ("p", "<f4", 3), ("n", "<f4", 3), ("u", "<f4"),
                           ("v", "<f4"), ("front_face", "<i4")])
    PIN_SCATTER_DT = np.dtype([("ok", "<i4"), ("att", "<f4", 3), ("o", "<f4", 3), ("d", "<f4", 3), ("tm", "<f4"),
                               ("emitted", "<f4", 3), ("draws", "<i4")])
    PIN_MEDIUM_DT = np.dtype([("hit", "<i4"), ("t", "<f4"), ("p", "<f4", 3), ("draws", "<i4")])
    INST_DT = np.dtype([("sin_y", "<f4"), ("cos_y", "<f4"), ("offset", "<f4", 3), ("flag", "<i4")])

    @staticmethod
    def _inst_ptr(inst):
        if inst is None:
            return None, None
        inst = np.ascontiguousarray(inst, L1.INST_DT)
        return inst, inst.ctypes.data

    def pin_hits(self, scene, profile, rays, t_min, t_max=3.4028234663852886e38, inst=None):
        rays = np.ascontiguousarray(rays, np.float32)
        out = np.zeros(len(rays), self.PIN_HIT_DT)
        d = scene.desc()
        keep, ip = self._inst_ptr(inst)
        self._pin_hits(C.byref(d), profile, ip, rays.ctypes.data, len(rays), float(t_min), float(t_max), out.ctypes.data)
        return out

    def pin_scatter(self, scene, profile, rays, t_min, seed, t_max=3.4028234663852886e38, inst=None):
        rays = np.ascontiguousarray(rays, np.float32)
        out = np.zeros(len(rays), self.PIN_SCATTER_DT)
        d = scene.desc()
        keep, ip = self._inst_ptr(inst)
        self._pin_scatter(C.byref(d), profile, ip, rays.ctypes.data, len(rays), float(t_min), float(t_max), int(seed),
                          out.ctypes.data)
        return out

    def pin_media(self, scene, rays, seed):
        rays = np.ascontiguousarray(rays, np.float32)
        out = np.zeros((len(rays), len(scene.media)), self.PIN_MEDIUM_DT)
        d = scene.desc()
        self._pin_media(C.byref(d), rays.ctypes.data, len(rays), int(seed), out.ctypes.data)
        return out

    def pin_color(self, scene, profile, rays, seed, inst=None):
        rays = np.ascontiguousarray(rays, np.float32)
        out = np.zeros((len(rays), 4), np.float32)
        d = scene.desc()
        keep, ip = self._inst_ptr(inst)
        self._pin_color(C.byref(d), profile, ip, rays.ctypes.data, len(rays), int(seed), out.ctypes.data)
        return out

    def xorwow_kat(self, seed, n=8):
        out = np.zeros(n, np.float32)
        self._xorwow_kat(int(seed), n, out.ctypes.data)
        return out

    def srand(self, seed):
        self.lib.orc_srand(seed)

    def closest_hit(self, scene, profile, rays, t_min=1e-3, t_max=np.inf, mode=0, bvh_seed=1984):
        rays = np.ascontiguousarray(rays, np.float32)
        n = len(rays)
        ids = np.empty(n, np.int32)
        ts = np.empty(n, np.float64)
        cnt = np.zeros(2, np.uint64)
        d = scene.desc()
        tmax = 3.4e38 if not np.isfinite(t_max) else float(t_max)
        rc = self._closest(C.byref(d), profile, rays.ctypes.data, n, float(t_min), tmax, mode, bvh_seed,
                           ids.ctypes.data, ts.ctypes.data, cnt.ctypes.data)
        assert rc == 0
        return ids, ts, cnt

    def closest_hit_packed(self, scene, profile, nodes, leaf_prims, big, rays, t_min=1e-3, t_max=np.inf):
        rays = np.ascontiguousarray(rays, np.float32)
        nodes = np.ascontiguousarray(nodes, BVH_NODE_DT)
        big = np.ascontiguousarray(big, np.int32)
        leaf_prims = np.ascontiguousarray(leaf_prims, np.int32)
        n = len(rays)
        ids = np.empty(n, np.int32)
        ts = np.empty(n, np.float64)
        cnt = np.zeros(2, np.uint64)
        d = scene.desc()
        tmax = 3.4e38 if not np.isfinite(t_max) else float(t_max)
        rc = self._packed(C.byref(d), profile, nodes.ctypes.data, len(nodes), leaf_prims.ctypes.data,
                          big.ctypes.data, len(big),
                          rays.ctypes.data, n, float(t_min), tmax, ids.ctypes.data, ts.ctypes.data, cnt.ctypes.data)
        assert rc == 0
        return ids, ts, cnt

    def render(self, scene, profile, W, H, spp, seed=1, rows=None, use_ref_bvh=False, libc_rand=False,
               want_sumsq=True):
        j0, j1 = rows if rows is not None else (0, H)
        s = np.zeros((H, W, 3), np.float64)
        s2 = np.zeros((H, W, 3), np.float64) if want_sumsq else None
        d = scene.desc()
        nseg = self._render(C.byref(d), profile, W, H, spp, seed, j0, j1, int(use_ref_bvh), int(libc_rand),
                            s.ctypes.data, s2.ctypes.data if want_sumsq else None)
        return s, s2, int(nseg)

    def render_parallel(self, scene, profile, W, H, spp, seed=1, threads=None, use_ref_bvh=False):
        """Row slabs on a thread pool (ctypes releases the GIL; per-pixel RNG streams
        make the image independent of the split)."""
        from concurrent.futures import ThreadPoolExecutor
        threads = threads or os.cpu_count() or 1
        s = np.zeros((H, W, 3), np.float64)
        s2 = np.zeros((H, W, 3), np.float64)
        d = scene.desc()
        edges = np.linspace(0, H, min(H, threads * 4) + 1).astype(int)

        def job(k):
            return self._render(C.byref(d), profile, W, H, spp, seed, int(edges[k]), int(edges[k + 1]),
                                int(use_ref_bvh), 0, s.ctypes.data, s2.ctypes.data)

        with ThreadPoolExecutor(threads) as ex:
            nseg = sum(ex.map(job, range(len(edges) - 1)))
        return s, s2, int(nseg)

    def scatter_one(self, scene, profile, material, d_in, outward_n, seed, libc_rand=False):
        d_in = np.ascontiguousarray(d_in, np.float64)
        n = np.ascontiguousarray(outward_n, np.float64)
        out = np.zeros(3)
        att = np.zeros(3)
        d = scene.desc()
        ok = self._scatter(C.byref(d), profile, material, d_in.ctypes.data, n.ctypes.data, seed, int(libc_rand),
                           out.ctypes.data, att.ctypes.data)
        return ok, out, att

    def medium_reference_order_demo(self, scene, ray8, order, n_trials=20000, seed=1):
        """fraction of trials in which the medium is the closest hit under the REFERENCE's
        constant_medium::hit (no t_max test), media visited before (0) / after (1) the surfaces"""
        ray8 = np.ascontiguousarray(ray8, np.float32)
        d = scene.desc()
        return float(self._mdemo(C.byref(d), ray8.ctypes.data, order, n_trials, seed))

    def tex_value(self, scene, material, uvp):
        """texture value at rows (u, v, px, py, pz)"""
        uvp = np.ascontiguousarray(uvp, np.float64).reshape(-1, 5)
        out = np.zeros((len(uvp), 3))
        d = scene.desc()
        self._tex(C.byref(d), material, len(uvp), uvp.ctypes.data, out.ctypes.data)
        return out

    def sphere_uv(self, normals):
        n = np.ascontiguousarray(normals, np.float64).reshape(-1, 3)
        out = np.zeros((len(n), 2))
        self._suv(len(n), n.ctypes.data, out.ctypes.data)
        return out

    def quantise(self, profile, sum3, spp):
        sum3 = np.ascontiguousarray(sum3, np.float64)
        rgb = np.zeros(3, np.int32)
        self._quant(profile, sum3.ctypes.data, spp, rgb.ctypes.data)
        return rgb


WEEKEND_CAM13 = lambda aspect: np.array([13, 2, 3, 0, 0, 0, 0, 1, 0, 20.0, aspect, 0.1, 10.0], np.float64)


class L0:
    """The reference's own CPU renderer (rt_in_one_weekend), unmodified."""

    def __init__(self, asis=False):
        path = L0_ASIS_PATH if asis else L0_PATH
        if not os.path.exists(path):
            build()
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (reference sources absent and no prebuilt oracle/_ref)")
        self.lib = L = C.CDLL(path)
        L.l0_scene_build.restype = C.c_int
        L.l0_scene_build.argtypes = [C.c_uint]
        L.l0_scene_get.restype = C.c_int
        L.l0_scene_get.argtypes = [_vp, C.c_int]
        L.l0_camera_get.argtypes = [_vp, _vp]
        L.l0_closest_hit.restype = C.c_int
        L.l0_closest_hit.argtypes = [_vp, C.c_int, C.c_double, C.c_double, _vp, _vp]
        L.l0_hit_robustness.restype = C.c_int
        L.l0_hit_robustness.argtypes = [_vp, C.c_int, C.c_double, C.c_double, _vp, _vp]
        L.l0_render.restype = C.c_uint64
        L.l0_render.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _vp, C.c_uint, C.c_int, C.c_int, _vp, _vp]
        L.l0_worker_timed.restype = C.c_double
        L.l0_worker_timed.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _vp, C.c_uint, C.c_int, C.c_int, C.c_int,
                                      _vp, _vp]
        L.l0_write_color.argtypes = [_vp, C.c_int, _vp]
        L.l0_scatter.restype = C.c_int
        L.l0_scatter.argtypes = [C.c_int, _vp, C.c_double, _vp, _vp, _vp, _vp]
        L.l0_srand.argtypes = [C.c_uint]
        L.l0_scene_set.restype = C.c_int
        L.l0_scene_set.argtypes = [_vp, C.c_int]
        L.l0_camera_override.argtypes = [_vp]
        self.n = L.l0_scene_build(1)

    def scene_set(self, rows):
        rows = np.ascontiguousarray(rows, np.float64)
        assert rows.shape[1] == 12
        self.n = self.lib.l0_scene_set(rows.ctypes.data, len(rows))
        assert self.n == len(rows)

    def camera_override(self, cam22):
        if cam22 is None:
            self.lib.l0_camera_override(None)
        else:
            cam22 = np.ascontiguousarray(cam22, np.float64)
            assert cam22.shape == (22,)
            self.lib.l0_camera_override(cam22.ctypes.data)

    def scene_rows(self):
        out = np.zeros((self.n, 12), np.float64)
        rc = self.lib.l0_scene_get(out.ctypes.data, self.n)
        assert rc == self.n, rc
        return out

    def camera22(self, cam13):
        cam13 = np.ascontiguousarray(cam13, np.float64)
        out = np.zeros(22, np.float64)
        self.lib.l0_camera_get(cam13.ctypes.data, out.ctypes.data)
        return out

    def closest_hit(self, rays6, t_min=1e-3, t_max=np.inf):
        rays6 = np.ascontiguousarray(rays6, np.float64)
        n = len(rays6)
        ids = np.empty(n, np.int32)
        ts = np.empty(n, np.float64)
        rc = self.lib.l0_closest_hit(rays6.ctypes.data, n, t_min, t_max, ids.ctypes.data, ts.ctypes.data)
        assert rc == 0
        return ids, ts

    def hit_robustness(self, rays6, t_min=1e-3, t_max=np.inf):
        rays6 = np.ascontiguousarray(rays6, np.float64)
        n = len(rays6)
        gap = np.empty(n, np.float64)
        dr = np.empty(n, np.float64)
        rc = self.lib.l0_hit_robustness(rays6.ctypes.data, n, t_min, t_max, gap.ctypes.data, dr.ctypes.data)
        assert rc == 0
        return gap, dr

    def render(self, W, H, spp, cam13, seed=1, rows=None, depth=50, want_sumsq=True):
        j0, j1 = rows if rows is not None else (0, H)
        cam13 = np.ascontiguousarray(cam13, np.float64)
        s = np.zeros((H, W, 3), np.float64)
        s2 = np.zeros((H, W, 3), np.float64) if want_sumsq else None
        nseg = self.lib.l0_render(W, H, spp, depth, cam13.ctypes.data, seed, j0, j1, s.ctypes.data,
                                  s2.ctypes.data if want_sumsq else None)
        return s, s2, int(nseg)

    def worker_timed(self, W, H, spp, cam13, start, end, threads, seed=1, depth=50, count_segments=True,
                     want_sum=False):
        cam13 = np.ascontiguousarray(cam13, np.float64)
        s = np.zeros((H, W, 3), np.float64) if want_sum else None
        seg = C.c_uint64(0)
        secs = self.lib.l0_worker_timed(W, H, spp, depth, cam13.ctypes.data, seed, start, end, threads,
                                        s.ctypes.data if want_sum else None,
                                        C.addressof(seg) if count_segments else None)
        return secs, int(seg.value), s

    def write_color(self, sum3, spp):
        sum3 = np.ascontiguousarray(sum3, np.float64)
        rgb = np.zeros(3, np.int32)
        self.lib.l0_write_color(sum3.ctypes.data, spp, rgb.ctypes.data)
        return rgb

    def scatter(self, kind, albedo, param, d_in, outward_n):
        albedo = np.ascontiguousarray(albedo, np.float64)
        d_in = np.ascontiguousarray(d_in, np.float64)
        n = np.ascontiguousarray(outward_n, np.float64)
        out = np.zeros(3)
        att = np.zeros(3)
        ok = self.lib.l0_scatter(kind, albedo.ctypes.data, float(param), d_in.ctypes.data, n.ctypes.data,
                                 out.ctypes.data, att.ctypes.data)
        return ok, out, att

    def srand(self, seed):
        self.lib.l0_srand(seed)
