"""ctypes / numpy mirrors of the plain-C structs in include/rt_capi.h.

Kept free of any library loading: it only describes the flattened scene
(structs, dtypes, camera construction, deterministic primary rays).
"""
import ctypes as C

import numpy as np

RT_PROFILE_WEEKEND_CPU = 0
RT_PROFILE_FINAL_CU = 1
RT_PROFILE_NEXT_WEEK = 2

RT_FLAG_FLIP_NORMALS = 1
RT_FLAG_DEPTH_BACKGROUND = 2
RT_FLAG_COUNTERS = 4
RT_FLAG_REFERENCE_MEDIUM = 8

RT_PRIM_SPHERE, RT_PRIM_TRIANGLE, RT_PRIM_QUAD, RT_PRIM_MEDIUM, RT_PRIM_INSTANCE, RT_PRIM_BOX = 0, 1, 2, 3, 4, 5
RT_MAT_LAMBERTIAN, RT_MAT_METAL, RT_MAT_DIELECTRIC, RT_MAT_DIFFUSE_LIGHT, RT_MAT_ISOTROPIC = 0, 1, 2, 3, 4
RT_TEX_SOLID, RT_TEX_CHECKER, RT_TEX_NOISE, RT_TEX_IMAGE = 0, 1, 2, 3


def prim_id(ptype, index):
    return np.int32((ptype << 28) | index)


def prim_type_of(pid):
    return (np.asarray(pid).astype(np.uint32) >> 28).astype(np.int32)


def prim_index_of(pid):
    return (np.asarray(pid).astype(np.uint32) & 0x0FFFFFFF).astype(np.int32)


SPHERE_DT = np.dtype(
    [("center0", "<f4", 3), ("radius", "<f4"), ("center1", "<f4", 3), ("material", "<i4"),
     ("time0", "<f4"), ("time1", "<f4"), ("moving", "<i4"), ("reserved", "<i4")]
)
TRIANGLE_DT = np.dtype(
    [("v0", "<f4", 3), ("v1", "<f4", 3), ("v2", "<f4", 3), ("normal", "<f4", 3), ("material", "<i4")]
)
QUAD_DT = np.dtype(
    [("axis", "<i4"), ("a0", "<f4"), ("a1", "<f4"), ("b0", "<f4"), ("b1", "<f4"), ("k", "<f4"), ("material", "<i4")]
)
MATERIAL_DT = np.dtype(
    [("type", "<i4"), ("texture", "<i4"), ("albedo", "<f4", 3), ("param", "<f4"), ("albedo2", "<f4", 3),
     ("reserved", "<f4")]
)
PERLIN_DT = np.dtype([("ranvec", "<f4", (256, 3)), ("perm_x", "<i4", 256), ("perm_y", "<i4", 256),
                      ("perm_z", "<i4", 256)])
MEDIUM_DT = np.dtype(
    [("shape", "<i4"), ("p0", "<f4", 3), ("p1", "<f4", 3), ("sin_y", "<f4"), ("cos_y", "<f4"), ("offset", "<f4", 3),
     ("density", "<f4"), ("material", "<i4")]
)
# two-level scenes (rt_group / rt_instance): an object = ranges of the primitive arrays in OBJECT space; an
# instance places it with a rigid 3x4 matrix (rows: rotation | translation), p_world = M p_object
GROUP_DT = np.dtype([("first_sphere", "<i4"), ("n_spheres", "<i4"), ("first_triangle", "<i4"), ("n_triangles", "<i4"),
                     ("first_quad", "<i4"), ("n_quads", "<i4"), ("reserved", "<i4", 2)])
INSTANCE_DT = np.dtype([("m", "<f4", 12), ("group", "<i4"), ("reserved", "<i4", 3)])
BVH_NODE_DT = np.dtype([("bmin", "<f4", 3), ("escape", "<i4"), ("bmax", "<f4", 3), ("payload", "<i4")])

assert SPHERE_DT.itemsize == 48 and TRIANGLE_DT.itemsize == 52 and QUAD_DT.itemsize == 28
assert MATERIAL_DT.itemsize == 40 and BVH_NODE_DT.itemsize == 32
assert PERLIN_DT.itemsize == 6144 and MEDIUM_DT.itemsize == 56
assert GROUP_DT.itemsize == 32 and INSTANCE_DT.itemsize == 64


def rigid_y(angle_deg, offset=(0.0, 0.0, 0.0), sin_cos=None):
    """The 3x4 matrix of translate(rotate_y(object, angle), offset) - rt_next_week/cuda/hittable.h:49-190:
    rotate_y maps object p to (cos x + sin z, y, -sin x + cos z) (:176-177), translate adds the offset (:73).
    sin_cos: use these (sin, cos) instead of computing them (e.g. the float values the reference computed)."""
    if sin_cos is None:
        rad = np.float32(np.float32(np.pi / 180.0) * np.float32(angle_deg))
        sn, cs = np.float32(np.sin(rad)), np.float32(np.cos(rad))
    else:
        sn, cs = np.float32(sin_cos[0]), np.float32(sin_cos[1])
    return np.array([cs, 0, sn, offset[0], 0, 1, 0, offset[1], -sn, 0, cs, offset[2]], np.float32)


class RtConfig(C.Structure):
    _fields_ = [("device", C.c_int32), ("profile", C.c_int32), ("flags", C.c_uint32), ("reserved", C.c_uint32),
                ("seed", C.c_uint64)]


class RtCamera(C.Structure):
    _fields_ = [("origin", C.c_float * 3), ("lower_left_corner", C.c_float * 3), ("horizontal", C.c_float * 3),
                ("vertical", C.c_float * 3), ("u", C.c_float * 3), ("v", C.c_float * 3), ("w", C.c_float * 3),
                ("lens_radius", C.c_float), ("time0", C.c_float), ("time1", C.c_float)]


class RtSceneDesc(C.Structure):
    _fields_ = [("n_spheres", C.c_int32), ("spheres", C.c_void_p),
                ("n_triangles", C.c_int32), ("triangles", C.c_void_p),
                ("n_quads", C.c_int32), ("quads", C.c_void_p),
                ("n_materials", C.c_int32), ("materials", C.c_void_p),
                ("camera", RtCamera),
                ("background", C.c_float * 3), ("sky_gradient", C.c_int32), ("t_min", C.c_float),
                ("max_depth", C.c_int32), ("flags", C.c_uint32), ("reserved", C.c_uint32),
                ("n_media", C.c_int32), ("media", C.c_void_p),
                ("n_perlin", C.c_int32), ("perlin", C.c_void_p),
                ("n_images", C.c_int32), ("images", C.c_void_p),
                ("n_groups", C.c_int32), ("groups", C.c_void_p),
                ("n_instances", C.c_int32), ("instances", C.c_void_p)]


class RtImage(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("rgb", C.c_void_p)]


class RtStats(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("segments", C.c_uint64), ("box_tests", C.c_uint64),
                ("prim_tests", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("ms_upload", C.c_float), ("ms_build", C.c_float), ("ms_render", C.c_float),
                ("ms_resolve", C.c_float),
                ("n_nodes", C.c_int32), ("n_big_prims", C.c_int32), ("smem_bytes", C.c_int32),
                ("block_threads", C.c_int32), ("grid_blocks", C.c_int32), ("regs_per_thread", C.c_int32),
                ("smem_plan", C.c_int32), ("reserved", C.c_int32),
                ("ms_k_render", C.c_float), ("ms_k_combine", C.c_float)]


def camera_from_lookat(lookfrom, lookat, vup, vfov_deg, aspect, aperture, focus_dist, time0=0.0, time1=0.0,
                       dtype=np.float64):
    """camera constructor — rt_in_one_weekend/camera.h:8-45 (dtype float64) and
    accelerated-rt-cuda/camera.h:20-53, rt_next_week/cuda/camera.h:25-61
    (dtype float32). Returns an RtCamera (fields rounded to float)."""
    f = dtype
    lookfrom = np.asarray(lookfrom, f)
    lookat = np.asarray(lookat, f)
    vup = np.asarray(vup, f)
    if dtype == np.float64:
        theta = f(vfov_deg) * f(3.1415926535897932385) / f(180.0)
        h = np.tan(theta / f(2))
        viewport_height = f(2.0) * h
        viewport_width = f(aspect) * viewport_height
        w = lookfrom - lookat
        w = w * (f(1) / np.sqrt(np.dot(w, w)))
        u = np.cross(vup, w)
        u = u * (f(1) / np.sqrt(np.dot(u, u)))
        v = np.cross(w, u)
        horizontal = f(focus_dist) * viewport_width * u
        vertical = f(focus_dist) * viewport_height * v
        llc = lookfrom - horizontal * f(0.5) - vertical * f(0.5) - f(focus_dist) * w
        lens_radius = f(aperture) / f(2)
    else:
        # float theta = vfov*M_PI/180 (double product rounded to float); tan() in double then rounded
        theta = np.float32(np.float64(np.float32(vfov_deg)) * np.pi / 180.0)
        half_height = np.float32(np.tan(np.float64(theta / np.float32(2))))
        half_width = np.float32(aspect) * half_height
        w = lookfrom - lookat
        w = w * (f(1) / np.sqrt(np.dot(w, w), dtype=f))
        u = np.cross(vup, w).astype(f)
        u = u * (f(1) / np.sqrt(np.dot(u, u), dtype=f))
        v = np.cross(w, u).astype(f)
        horizontal = (f(focus_dist) * f(2.0) * half_width) * u
        vertical = (f(focus_dist) * f(2.0) * half_height) * v
        llc = lookfrom - horizontal * f(0.5) - vertical * f(0.5) - f(focus_dist) * w
        lens_radius = f(aperture) / f(2)
    cam = RtCamera()
    for name, val in (("origin", lookfrom), ("lower_left_corner", llc), ("horizontal", horizontal),
                      ("vertical", vertical), ("u", u), ("v", v), ("w", w)):
        getattr(cam, name)[:] = [float(np.float32(x)) for x in val]
    cam.lens_radius = float(np.float32(lens_radius))
    cam.time0 = float(time0)
    cam.time1 = float(time1)
    return cam


class Scene:
    """Host-side flattened scene: numpy structured arrays + scalar fields.
    `.desc()` returns an RtSceneDesc whose pointers stay valid as long as this
    object is alive."""

    def __init__(self, spheres=None, triangles=None, quads=None, materials=None, camera=None,
                 background=(0.0, 0.0, 0.0), sky_gradient=1, t_min=1e-3, max_depth=50, flags=0, name="",
                 profile=RT_PROFILE_WEEKEND_CPU, media=None, perlin=None, images=None, groups=None, instances=None):
        self.groups = np.ascontiguousarray(groups if groups is not None else np.zeros(0, GROUP_DT))
        self.instances = np.ascontiguousarray(instances if instances is not None else np.zeros(0, INSTANCE_DT))
        assert self.groups.dtype == GROUP_DT and self.instances.dtype == INSTANCE_DT
        self.media = np.ascontiguousarray(media if media is not None else np.zeros(0, MEDIUM_DT))
        self.perlin = np.ascontiguousarray(perlin if perlin is not None else np.zeros(0, PERLIN_DT))
        # image textures: list of uint8 arrays [h][w][3], row 0 = top
        self.images = [np.ascontiguousarray(im, np.uint8) for im in (images or [])]
        assert self.media.dtype == MEDIUM_DT and self.perlin.dtype == PERLIN_DT
        assert all(im.ndim == 3 and im.shape[2] == 3 for im in self.images)
        self.spheres = np.ascontiguousarray(spheres if spheres is not None else np.zeros(0, SPHERE_DT))
        self.triangles = np.ascontiguousarray(triangles if triangles is not None else np.zeros(0, TRIANGLE_DT))
        self.quads = np.ascontiguousarray(quads if quads is not None else np.zeros(0, QUAD_DT))
        self.materials = np.ascontiguousarray(materials if materials is not None else np.zeros(0, MATERIAL_DT))
        assert self.spheres.dtype == SPHERE_DT and self.triangles.dtype == TRIANGLE_DT
        assert self.quads.dtype == QUAD_DT and self.materials.dtype == MATERIAL_DT
        self.camera = camera if camera is not None else RtCamera()
        self.background = tuple(float(x) for x in background)
        self.sky_gradient = int(sky_gradient)
        self.t_min = float(t_min)
        self.max_depth = int(max_depth)
        self.flags = int(flags)
        self.name = name
        self.profile = int(profile)

    @property
    def n_prims(self):
        return len(self.spheres) + len(self.triangles) + len(self.quads)

    def desc(self):
        d = RtSceneDesc()
        d.n_spheres = len(self.spheres)
        d.spheres = self.spheres.ctypes.data if len(self.spheres) else None
        d.n_triangles = len(self.triangles)
        d.triangles = self.triangles.ctypes.data if len(self.triangles) else None
        d.n_quads = len(self.quads)
        d.quads = self.quads.ctypes.data if len(self.quads) else None
        d.n_materials = len(self.materials)
        d.materials = self.materials.ctypes.data if len(self.materials) else None
        d.camera = self.camera
        d.background[:] = self.background
        d.sky_gradient = self.sky_gradient
        d.t_min = self.t_min
        d.max_depth = self.max_depth
        d.flags = self.flags
        d.n_media = len(self.media)
        d.media = self.media.ctypes.data if len(self.media) else None
        d.n_perlin = len(self.perlin)
        d.perlin = self.perlin.ctypes.data if len(self.perlin) else None
        d.n_images = len(self.images)
        if self.images:
            arr = (RtImage * len(self.images))()
            for k, im in enumerate(self.images):
                arr[k].width, arr[k].height, arr[k].rgb = im.shape[1], im.shape[0], im.ctypes.data
            d._images = arr
            d.images = C.addressof(arr)
        d.n_groups = len(self.groups)
        d.groups = self.groups.ctypes.data if len(self.groups) else None
        d.n_instances = len(self.instances)
        d.instances = self.instances.ctypes.data if len(self.instances) else None
        d._keepalive = self
        return d

    def with_camera(self, camera):
        s = Scene(self.spheres, self.triangles, self.quads, self.materials, camera, self.background,
                  self.sky_gradient, self.t_min, self.max_depth, self.flags, self.name, self.profile,
                  self.media, self.perlin, self.images, self.groups, self.instances)
        return s


def primary_rays(camera, W, H, profile, s_jitter=0.5, t_jitter=0.5, lens=(0.0, 0.0), time=0.0, pixels=None):
    """Deterministic primary rays in the layout rt_trace_closest takes ([n][8] floats:
    o.xyz, tm, d.xyz, pad), computed in float32 the way get_ray does
    (camera.h:56-62) with a fixed lens sample `lens` (unit-disk coordinates)
    and fixed sub-pixel offsets. Pixel order: j*W+i, j=0 bottom."""
    f = np.float32
    if pixels is None:
        jj, ii = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
        ii = ii.ravel()
        jj = jj.ravel()
    else:
        pixels = np.asarray(pixels)
        ii = pixels % W
        jj = pixels // W
    if profile == RT_PROFILE_WEEKEND_CPU:
        s = (ii.astype(f) + f(s_jitter)) / f(W - 1)
        t = (jj.astype(f) + f(t_jitter)) / f(H - 1)
    else:
        s = (ii.astype(f) + f(s_jitter)) / f(W)
        t = (jj.astype(f) + f(t_jitter)) / f(H)
    cu = np.array(camera.u[:], f)
    cv = np.array(camera.v[:], f)
    rdx = f(camera.lens_radius) * f(lens[0])
    rdy = f(camera.lens_radius) * f(lens[1])
    offset = cu * rdx + cv * rdy
    origin = np.array(camera.origin[:], f)
    llc = np.array(camera.lower_left_corner[:], f)
    hor = np.array(camera.horizontal[:], f)
    ver = np.array(camera.vertical[:], f)
    d = (llc[None, :] + s[:, None] * hor[None, :]).astype(f)
    d = (d + t[:, None] * ver[None, :]).astype(f)
    d = (d - origin[None, :]).astype(f)
    d = (d - offset[None, :]).astype(f)
    rays = np.zeros((len(ii), 8), f)
    rays[:, 0:3] = (origin + offset)[None, :]
    rays[:, 3] = f(time)
    rays[:, 4:7] = d
    return rays
