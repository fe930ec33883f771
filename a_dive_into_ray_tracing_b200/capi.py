"""ctypes binding of the C ABI (include/rt_capi.h -> libb200rt.so).

This is the product path: every call below lands in hand-written sm_100a kernels.
There is no CPU fallback — if the shared library is missing or no CUDA device is
present the constructors raise.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from .ctypes_defs import BVH_NODE_DT, RtConfig, RtSceneDesc, RtStats

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libb200rt.so")
CSRC = os.path.join(PKG_DIR, "csrc")

_vp = C.c_void_p
_lib = None

STATUS = {0: "RT_OK", 1: "RT_ERR_INVALID", 2: "RT_ERR_CUDA", 3: "RT_ERR_STATE", 4: "RT_ERR_NOMEM",
          5: "RT_ERR_NODEVICE"}

# every symbol include/rt_capi.h declares
SYMBOLS = ["rt_version", "rt_device_count", "rt_create", "rt_destroy", "rt_last_error", "rt_scene_upload",
           "rt_accel_build", "rt_accel_download", "rt_trace_closest", "rt_render", "rt_render_device", "rt_render_rows_device",
           "rt_accum_clear", "rt_accum_download", "rt_accum_upload", "rt_accum_device_ptr", "rt_resolve",
           "rt_resolve_device", "rt_render_aov", "rt_stats", "rt_stats_reset", "rt_sync", "rt_measure_fp32_peak",
           "rt_host_alloc", "rt_host_free", "rt_comm_unique_id", "rt_comm_init", "rt_comm_init_all", "rt_reduce",
           "rt_trace_closest_inst", "rt_instances_update"]

RT_REDUCE_UNIFORM_COUNT = 1
RT_COMM_ID_BYTES = 128


class RtError(RuntimeError):
    def __init__(self, code, msg=""):
        self.code = code
        super().__init__("%s: %s" % (STATUS.get(code, code), msg))


def build_library(force=False):
    """make -C csrc: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo, in-tree."""
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    deps.append(os.path.join(PKG_DIR, "..", "include", "rt_capi.h"))
    stale = (not os.path.exists(LIB_PATH)) or any(os.path.getmtime(d) > os.path.getmtime(LIB_PATH) for d in deps)
    if force or stale:
        subprocess.check_call(["make", "-C", CSRC, "-s"] + (["-B"] if force else []))
    return LIB_PATH


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("B200RT_LIB", LIB_PATH)  # tuning builds (csrc/Makefile variants)
    if not os.path.exists(path):
        raise OSError("libb200rt.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                      "there is no CPU fallback")
    L = C.CDLL(path)
    L.rt_version.restype = C.c_int
    L.rt_device_count.restype = C.c_int
    L.rt_create.restype = C.c_int
    L.rt_create.argtypes = [C.POINTER(_vp), C.POINTER(RtConfig)]
    L.rt_destroy.restype = None
    L.rt_destroy.argtypes = [_vp]
    L.rt_last_error.restype = C.c_char_p
    L.rt_last_error.argtypes = [_vp]
    L.rt_scene_upload.restype = C.c_int
    L.rt_scene_upload.argtypes = [_vp, C.POINTER(RtSceneDesc)]
    L.rt_accel_build.restype = C.c_int
    L.rt_accel_build.argtypes = [_vp, C.c_int]
    L.rt_accel_download.restype = C.c_int
    L.rt_accel_download.argtypes = [_vp, _vp, C.c_int, _vp, _vp, C.c_int, _vp, _vp, C.c_int, _vp]
    L.rt_trace_closest.restype = C.c_int
    L.rt_trace_closest.argtypes = [_vp, _vp, C.c_int, C.c_float, C.c_float, C.c_int, _vp, _vp]
    L.rt_trace_closest_inst.restype = C.c_int
    L.rt_trace_closest_inst.argtypes = [_vp, _vp, C.c_int, C.c_float, C.c_float, C.c_int, _vp, _vp, _vp]
    L.rt_instances_update.restype = C.c_int
    L.rt_instances_update.argtypes = [_vp, _vp, C.c_int]
    L.rt_render.restype = C.c_int
    L.rt_render.argtypes = [_vp, C.c_int, C.c_int, C.c_int, C.c_int]
    L.rt_render_device.restype = C.c_int
    L.rt_render_device.argtypes = [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp]
    L.rt_accum_clear.restype = C.c_int
    L.rt_accum_clear.argtypes = [_vp]
    L.rt_accum_download.restype = C.c_int
    L.rt_accum_download.argtypes = [_vp, _vp, C.c_size_t]
    L.rt_accum_upload.restype = C.c_int
    L.rt_accum_upload.argtypes = [_vp, C.c_int, C.c_int, _vp, C.c_size_t]
    L.rt_accum_device_ptr.restype = _vp
    L.rt_accum_device_ptr.argtypes = [_vp]
    L.rt_resolve.restype = C.c_int
    L.rt_resolve.argtypes = [_vp, _vp, _vp]
    L.rt_render_rows_device.restype = C.c_int
    L.rt_render_rows_device.argtypes = [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp]
    L.rt_render_aov.restype = C.c_int
    L.rt_render_aov.argtypes = [_vp, C.c_int, C.c_int, C.c_int, _vp]
    L.rt_resolve_device.restype = C.c_int
    L.rt_resolve_device.argtypes = [_vp, C.c_int, C.c_int, _vp, _vp, _vp, _vp]
    L.rt_stats.restype = C.c_int
    L.rt_stats.argtypes = [_vp, C.POINTER(RtStats)]
    L.rt_stats_reset.restype = C.c_int
    L.rt_stats_reset.argtypes = [_vp]
    L.rt_measure_fp32_peak.restype = C.c_int
    L.rt_measure_fp32_peak.argtypes = [_vp, C.POINTER(C.c_float)]
    L.rt_sync.restype = C.c_int
    L.rt_sync.argtypes = [_vp]
    L.rt_host_alloc.restype = _vp
    L.rt_host_alloc.argtypes = [_vp, C.c_size_t]
    L.rt_host_free.restype = None
    L.rt_host_free.argtypes = [_vp, _vp]
    L.rt_comm_unique_id.restype = C.c_int
    L.rt_comm_unique_id.argtypes = [_vp]
    L.rt_comm_init.restype = C.c_int
    L.rt_comm_init.argtypes = [_vp, _vp, C.c_int, C.c_int]
    L.rt_comm_init_all.restype = C.c_int
    L.rt_comm_init_all.argtypes = [C.POINTER(_vp), C.c_int]
    L.rt_reduce.restype = C.c_int
    L.rt_reduce.argtypes = [_vp, C.c_int, C.c_int, _vp, C.c_int, C.c_uint, _vp]
    _lib = L
    return L


def device_count():
    return load_library().rt_device_count()


def comm_unique_id():
    """128 bytes naming a new NCCL communicator (rank 0 creates it and ships it to the other ranks)."""
    buf = C.create_string_buffer(RT_COMM_ID_BYTES)
    rc = load_library().rt_comm_unique_id(buf)
    if rc:
        raise RtError(rc, "rt_comm_unique_id (is libnccl.so.2 loadable?)")
    return buf.raw


class Context:
    """One rt_ctx: a scene + accumulation buffer on ONE CUDA device."""

    def __init__(self, profile=0, device=0, seed=1984, flags=0):
        self.lib = load_library()
        cfg = RtConfig(device=device, profile=profile, flags=flags, reserved=0, seed=seed)
        h = _vp()
        rc = self.lib.rt_create(C.byref(h), C.byref(cfg))
        if rc:
            raise RtError(rc, "rt_create (no CUDA device? there is no CPU fallback)")
        self.h = h
        self.profile = profile
        self.device = device
        self.W = self.H = 0
        self._scene = None
        self._pinned = []

    def close(self):
        if getattr(self, "h", None):
            for ptr in self._pinned:
                self.lib.rt_host_free(self.h, ptr)
            self._pinned = []
            self.lib.rt_destroy(self.h)
            self.h = None

    def pinned_array(self, shape, dtype):
        """numpy array over page-locked host memory (rt_host_alloc); lives as long as the context"""
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        ptr = self.lib.rt_host_alloc(self.h, max(n, 1))
        if not ptr:
            raise RtError(4, "rt_host_alloc(%d)" % n)
        self._pinned.append(ptr)
        buf = (C.c_char * max(n, 1)).from_address(ptr)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def comm_init(self, unique_id, rank, nranks):
        """join the NCCL communicator `unique_id` (comm_unique_id() of rank 0) as `rank` of `nranks`"""
        assert len(unique_id) == RT_COMM_ID_BYTES
        self._ck(self.lib.rt_comm_init(self.h, unique_id, rank, nranks))

    def reduce(self, W, H, d_accum_ptr=None, root=0, uniform_count=True, stream_ptr=None):
        """sum the ranks' float4 frames onto root's (rt_reduce: pack -> ncclReduce -> unpack), asynchronously"""
        self._ck(self.lib.rt_reduce(self.h, W, H, d_accum_ptr, root, RT_REDUCE_UNIFORM_COUNT if uniform_count else 0,
                                    self._stream(stream_ptr)))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        if rc:
            raise RtError(rc, self.lib.rt_last_error(self.h).decode())

    def upload(self, scene):
        d = scene.desc()
        self._ck(self.lib.rt_scene_upload(self.h, C.byref(d)))
        self._scene = scene
        return self

    def build_accel(self, quality=1):
        self._ck(self.lib.rt_accel_build(self.h, quality))
        return self

    def accel(self):
        """(nodes, leaf_prims, big): the packed threaded BVH as the kernels read it."""
        nn, nl, nb = C.c_int(), C.c_int(), C.c_int()
        self._ck(self.lib.rt_accel_download(self.h, None, 0, C.addressof(nn), None, 0, C.addressof(nl), None, 0,
                                            C.addressof(nb)))
        nodes = np.zeros(max(nn.value, 1), BVH_NODE_DT)
        leaf = np.zeros(max(nl.value, 1), np.int32)
        big = np.zeros(max(nb.value, 1), np.int32)
        self._ck(self.lib.rt_accel_download(self.h, nodes.ctypes.data, len(nodes), None, leaf.ctypes.data, len(leaf),
                                            None, big.ctypes.data, len(big), None))
        return nodes[:nn.value], leaf[:nl.value], big[:nb.value]

    def trace_closest(self, rays, t_min=1e-3, t_max=np.inf, use_accel=True, with_instances=False):
        """(prim ids, t) of the closest hits; with_instances: (prim ids, instance ids, t) - the instance whose
        object holds the hit primitive, -1 for world-level primitives and misses."""
        rays = np.ascontiguousarray(rays, np.float32)
        assert rays.ndim == 2 and rays.shape[1] == 8
        n = len(rays)
        ids = np.empty(n, np.int32)
        ts = np.empty(n, np.float32)
        tmax = float("inf") if not np.isfinite(t_max) else float(t_max)
        if with_instances:
            inst = np.empty(n, np.int32)
            self._ck(self.lib.rt_trace_closest_inst(self.h, rays.ctypes.data, n, t_min, tmax, int(use_accel),
                                                    ids.ctypes.data, inst.ctypes.data, ts.ctypes.data))
            return ids, inst, ts
        self._ck(self.lib.rt_trace_closest(self.h, rays.ctypes.data, n, t_min, tmax, int(use_accel),
                                           ids.ctypes.data, ts.ctypes.data))
        return ids, ts

    def update_instances(self, instances):
        """Transform-only update of a two-level scene: rebuilds the top level, keeps the objects' trees."""
        instances = np.ascontiguousarray(instances)
        from .ctypes_defs import INSTANCE_DT
        assert instances.dtype == INSTANCE_DT
        self._ck(self.lib.rt_instances_update(self.h, instances.ctypes.data if len(instances) else None, len(instances)))
        return self

    def render(self, W, H, spp_count, spp_begin=0):
        self._ck(self.lib.rt_render(self.h, W, H, spp_begin, spp_count))
        self.W, self.H = W, H
        return self

    @staticmethod
    def _stream(stream_ptr):
        # None -> the context's own stream (NULL in the C ABI); torch's default stream has
        # handle 0, which must be spelled cudaStreamLegacy (0x1) to be distinguishable.
        if stream_ptr is None:
            return None
        return 1 if int(stream_ptr) == 0 else int(stream_ptr)

    def render_device(self, W, H, spp_count, spp_begin, d_accum_ptr, stream_ptr=None):
        self._ck(self.lib.rt_render_device(self.h, W, H, spp_begin, spp_count, d_accum_ptr,
                                           self._stream(stream_ptr)))

    def render_rows_device(self, W, H, row_begin, row_end, spp_count, spp_begin, d_accum_ptr, stream_ptr=None):
        """rows [row_begin, row_end) only (row 0 = bottom); the rest of the frame is left untouched"""
        self._ck(self.lib.rt_render_rows_device(self.h, W, H, row_begin, row_end, spp_begin, spp_count, d_accum_ptr,
                                                self._stream(stream_ptr)))

    def clear(self):
        self._ck(self.lib.rt_accum_clear(self.h))

    def accum(self):
        """[H][W][4] float32: sum R, G, B and sample count; row 0 = image bottom."""
        a = np.empty((self.H, self.W, 4), np.float32)
        self._ck(self.lib.rt_accum_download(self.h, a.ctypes.data, a.size))
        return a

    def accum_upload(self, a):
        a = np.ascontiguousarray(a, np.float32)
        H, W = a.shape[:2]
        self._ck(self.lib.rt_accum_upload(self.h, W, H, a.ctypes.data, a.size))
        self.W, self.H = W, H

    def resolve(self, want_linear=True, want_rgb8=True):
        lin = np.empty((self.H, self.W, 3), np.float32) if want_linear else None
        rgb = np.empty((self.H, self.W, 3), np.uint8) if want_rgb8 else None
        self._ck(self.lib.rt_resolve(self.h, lin.ctypes.data if want_linear else None,
                                     rgb.ctypes.data if want_rgb8 else None))
        return lin, rgb

    def resolve_device(self, W, H, d_accum_ptr, want_linear=False, want_rgb8=True, stream_ptr=None, out_rgb8=None):
        """out_rgb8: a caller buffer ([H][W][3] uint8, e.g. pinned_array) to resolve into instead of a new array"""
        lin = np.empty((H, W, 3), np.float32) if want_linear else None
        rgb = (out_rgb8 if out_rgb8 is not None else np.empty((H, W, 3), np.uint8)) if want_rgb8 else None
        assert rgb is None or (rgb.shape == (H, W, 3) and rgb.dtype == np.uint8 and rgb.flags.c_contiguous)
        self._ck(self.lib.rt_resolve_device(self.h, W, H, d_accum_ptr, lin.ctypes.data if want_linear else None,
                                            rgb.ctypes.data if want_rgb8 else None, self._stream(stream_ptr)))
        return lin, rgb

    def render_aov(self, W, H, spp=1):
        """first-hit feature buffers [H][W][8]: albedo rgb, normal xyz, t, hit fraction (bottom row first)"""
        out = np.empty((H, W, 8), np.float32)
        self._ck(self.lib.rt_render_aov(self.h, W, H, spp, out.ctypes.data))
        return out

    def stats(self):
        s = RtStats()
        self._ck(self.lib.rt_stats(self.h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in RtStats._fields_}

    def stats_reset(self):
        self._ck(self.lib.rt_stats_reset(self.h))

    def measure_fp32_peak(self):
        v = C.c_float(0)
        self._ck(self.lib.rt_measure_fp32_peak(self.h, C.byref(v)))
        return float(v.value)

    def sync(self):
        self._ck(self.lib.rt_sync(self.h))


def write_ppm(path, rgb8, binary=True):
    """P6 (binary) or P3 (the reference's text format, main.cpp:344-355) PPM; rgb8 is
    [H][W][3] top row first as rt_resolve returns it."""
    H, W = rgb8.shape[:2]
    if binary:
        with open(path, "wb") as fh:
            fh.write(b"P6\n%d %d\n255\n" % (W, H))
            fh.write(np.ascontiguousarray(rgb8, np.uint8).tobytes())
    else:
        with open(path, "w") as fh:
            fh.write("P3\n%d %d\n255\n" % (W, H))
            np.savetxt(fh, rgb8.reshape(-1, 3), fmt="%d")


def write_png(path, rgb8):
    """8-bit RGB PNG (zlib-compressed); rgb8 is [H][W][3], top row first."""
    import struct
    import zlib
    H, W = rgb8.shape[:2]
    raw = b"".join(b"\x00" + np.ascontiguousarray(rgb8[j], np.uint8).tobytes() for j in range(H))

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)

    with open(path, "wb") as fh:
        fh.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", W, H, 8, 2, 0, 0, 0)) +
                 chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))
