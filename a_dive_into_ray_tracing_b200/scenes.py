"""Flattened scenes of the reference's configurations (BASELINE.json `configs`).

The reference hard-codes its scenes inside `main()`s; here they are data:
  weekend()    rt_in_one_weekend/main.cpp:86-131 random_scene() under glibc's default
               seed, extracted once from the compiled reference (tools/make_golden.py)
               and stored float-rounded in scenes/weekend_scene.npy;
  final_cu()   accelerated-rt-cuda/final.cu:100-143 distribution (cuRAND stream not
               reproducible -> Philox-free numpy generator with the same law);
  next_week()  rt_next_week/cuda/main.cu:153-198 (moving spheres, checker ground);
  obj_room()   triangles/cuda/obj_render.cu:384-524 (mesh in a lit, mirrored room).
"""
import os

import numpy as np

from .ctypes_defs import (GROUP_DT, INSTANCE_DT, rigid_y, MATERIAL_DT, MEDIUM_DT, PERLIN_DT, QUAD_DT, RT_FLAG_DEPTH_BACKGROUND, RT_FLAG_FLIP_NORMALS,
                          RT_MAT_DIELECTRIC, RT_MAT_DIFFUSE_LIGHT, RT_MAT_ISOTROPIC, RT_MAT_LAMBERTIAN, RT_MAT_METAL,
                          RT_PROFILE_FINAL_CU, RT_PROFILE_NEXT_WEEK, RT_PROFILE_WEEKEND_CPU, RT_TEX_CHECKER,
                          RT_TEX_IMAGE, RT_TEX_NOISE, RT_TEX_SOLID, SPHERE_DT, TRIANGLE_DT, Scene, camera_from_lookat)

# scene data shipped with the product (not the test tree): the headline scene and the stand-in meshes
DATA_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scenes")


def _mat(mtype, albedo=(0, 0, 0), param=0.0, texture=RT_TEX_SOLID, albedo2=(0, 0, 0)):
    m = np.zeros((), MATERIAL_DT)
    m["type"] = mtype
    m["texture"] = texture
    m["albedo"] = albedo
    m["param"] = param
    m["albedo2"] = albedo2
    return m


def _sphere(c0, r, mat, c1=None, t0=0.0, t1=1.0):
    s = np.zeros((), SPHERE_DT)
    s["center0"] = c0
    s["radius"] = r
    s["center1"] = c0 if c1 is None else c1
    s["material"] = mat
    s["time0"] = t0
    s["time1"] = t1
    s["moving"] = 0 if c1 is None else 1
    return s


def scene_from_rows(rows, name="rows"):
    """rows [n][12] = cx cy cz r kind a0 a1 a2 param (oracle/ref_harness.cpp l0_scene_get).
    One material per sphere, as the reference allocates them."""
    rows = np.asarray(rows, np.float64)
    n = len(rows)
    spheres = np.zeros(n, SPHERE_DT)
    mats = np.zeros(n, MATERIAL_DT)
    spheres["center0"] = rows[:, 0:3]
    spheres["center1"] = rows[:, 0:3]
    spheres["radius"] = rows[:, 3]
    spheres["material"] = np.arange(n)
    spheres["time1"] = 1.0
    mats["type"] = rows[:, 4].astype(np.int32)
    mats["albedo"] = rows[:, 5:8]
    mats["param"] = rows[:, 8]
    return Scene(spheres=spheres, materials=mats, name=name)


def rows_from_scene(scene):
    """Inverse of scene_from_rows (float values widened exactly to double)."""
    n = len(scene.spheres)
    rows = np.zeros((n, 12), np.float64)
    rows[:, 0:3] = scene.spheres["center0"]
    rows[:, 3] = scene.spheres["radius"]
    m = scene.materials[scene.spheres["material"]]
    rows[:, 4] = m["type"]
    rows[:, 5:8] = m["albedo"]
    rows[:, 8] = m["param"]
    return rows


def weekend_camera(aspect, dtype=np.float64):
    """main.cpp:304-311: lookfrom (13,2,3), lookat 0, vfov 20, aperture 0.1, focus 10."""
    return camera_from_lookat((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, aspect, 0.1, 10.0, dtype=dtype)


def weekend(width=1200, height=800):
    """Configs 1, 2, 5: the reference's deterministic 487-sphere scene."""
    rows = np.load(os.path.join(DATA_DIR, "weekend_scene.npy"))
    sc = scene_from_rows(rows, "weekend")
    sc.camera = weekend_camera(width / height)
    sc.sky_gradient = 1
    sc.t_min = 1e-3
    sc.max_depth = 50
    sc.profile = RT_PROFILE_WEEKEND_CPU
    return sc


def _random_spheres(rng, moving):
    """The 22x22 grid law shared by final.cu:105-131 and main.cu:153-198."""
    spheres, mats = [], []

    def add(s, m):
        s["material"] = len(mats)
        spheres.append(s)
        mats.append(m)

    R = lambda: np.float32(rng.random())
    for a in range(-11, 11):
        for b in range(-11, 11):
            choose = R()
            center = (np.float32(a) + R(), np.float32(0.2), np.float32(b) + R())
            if choose < 0.8:
                c1 = None
                if moving:
                    c1 = (center[0], center[1] + R() * np.float32(0.5), center[2])
                add(_sphere(center, 0.2, 0, c1), _mat(RT_MAT_LAMBERTIAN, (R() * R(), R() * R(), R() * R())))
            elif choose < 0.95:
                alb = tuple(np.float32(0.5) * (np.float32(1.0) + R()) for _ in range(3))
                add(_sphere(center, 0.2, 0), _mat(RT_MAT_METAL, alb, np.float32(0.5) * R()))
            else:
                add(_sphere(center, 0.2, 0), _mat(RT_MAT_DIELECTRIC, param=1.5))
    add(_sphere((0, 1, 0), 1.0, 0), _mat(RT_MAT_DIELECTRIC, param=1.5))
    add(_sphere((-4, 1, 0), 1.0, 0), _mat(RT_MAT_LAMBERTIAN, (0.4, 0.2, 0.1)))
    add(_sphere((4, 1, 0), 1.0, 0), _mat(RT_MAT_METAL, (0.7, 0.6, 0.5), 0.0))
    return spheres, mats


def final_cu(width=1200, height=800, seed=1984):
    """accelerated-rt-cuda/final.cu:100-143: 488 spheres, ground at (0,-1000,-1), vfov 30."""
    rng = np.random.Generator(np.random.Philox(seed))
    spheres, mats = _random_spheres(rng, moving=False)
    ground = _sphere((0, -1000.0, -1), 1000, 0)
    ground["material"] = len(mats)
    spheres.insert(0, ground)
    mats.append(_mat(RT_MAT_LAMBERTIAN, (0.5, 0.5, 0.5)))
    sc = Scene(spheres=np.array(spheres, SPHERE_DT), materials=np.array(mats, MATERIAL_DT), name="final_cu")
    sc.camera = camera_from_lookat((13, 2, 3), (0, 0, 0), (0, 1, 0), 30.0, np.float32(width) / np.float32(height),
                                   0.1, 10.0, dtype=np.float32)
    sc.profile = RT_PROFILE_FINAL_CU
    return sc


def next_week(width=1200, height=800, seed=1984):
    """Config 4 — rt_next_week/cuda/main.cu:153-198 + :402-407,462-465: checker ground,
    80 % moving lambertian, camera vfov 20, aperture 0.05, focus |lookfrom-lookat|,
    shutter [0,1], constant background (0.7,0.8,1.0)."""
    rng = np.random.Generator(np.random.Philox(seed))
    spheres, mats = _random_spheres(rng, moving=True)
    ground = _sphere((0, -1000.0, -1), 1000, 0)
    ground["material"] = len(mats)
    spheres.insert(0, ground)
    mats.append(_mat(RT_MAT_LAMBERTIAN, (0.2, 0.3, 0.1), texture=RT_TEX_CHECKER, albedo2=(0.9, 0.9, 0.9)))
    sc = Scene(spheres=np.array(spheres, SPHERE_DT), materials=np.array(mats, MATERIAL_DT), name="next_week")
    lookfrom = np.array([13, 2, 3], np.float32)
    focus = np.sqrt(np.dot(lookfrom, lookfrom), dtype=np.float32)
    sc.camera = camera_from_lookat(lookfrom, (0, 0, 0), (0, 1, 0), 20.0, np.float32(width) / np.float32(height),
                                   0.05, focus, 0.0, 1.0, dtype=np.float32)
    sc.background = (0.70, 0.80, 1.00)
    sc.sky_gradient = 0
    sc.t_min = 1e-3
    sc.profile = RT_PROFILE_NEXT_WEEK
    return sc


# ---------------------------------------------------------------- triangles / OBJ

def triangle_record(v0, v1, v2, vn0, vn1, vn2, material):
    """triangle ctor — triangles/cuda/include/triangle.h:17-53 (float32 arithmetic):
    face normal = +-cross(v1-v0, v2-v0), oriented to agree with -(vn0+vn1+vn2)/3."""
    f = np.float32
    v0, v1, v2 = (np.asarray(x, f) for x in (v0, v1, v2))
    vn0, vn1, vn2 = (np.asarray(x, f) for x in (vn0, vn1, vn2))
    avg = -(((vn0 + vn1).astype(f) + vn2).astype(f)) * (f(1) / f(3.0))
    ab = (v1 - v0).astype(f)
    ac = (v2 - v0).astype(f)
    cand = np.array([ab[1] * ac[2] - ab[2] * ac[1], ab[2] * ac[0] - ab[0] * ac[2], ab[0] * ac[1] - ab[1] * ac[0]], f)
    d = f(f(f(cand[0] * avg[0]) + f(cand[1] * avg[1])) + f(cand[2] * avg[2]))
    n = cand if d > 0 else -cand
    t = np.zeros((), TRIANGLE_DT)
    t["v0"], t["v1"], t["v2"], t["normal"], t["material"] = v0, v1, v2, n, material
    return t


def read_obj_triangles(path):
    """read_triangles — triangles/cuda/include/triangle.h:217-300: `v`, `vn`, `f a/b/c`
    x3 (or a//c); the face is emitted with vertex indices [6],[3],[0] and normal indices
    [8],[5],[2] of the flattened a/b/c list, i.e. with REVERSED winding. Returns
    (v[n][3][3], vn[n][3][3]) float32, un-transformed."""
    vs, vns, tv, tn = [], [], [], []
    with open(path) as fh:
        for line in fh:
            parts = line.split()
            if not parts:
                continue
            if parts[0] == "vn":
                vns.append([float(x) for x in parts[1:4]])
            elif parts[0] == "v":
                vs.append([float(x) for x in parts[1:4]])
            elif parts[0] == "f":
                idx = []
                for sec in parts[1:]:
                    for num in sec.split("/"):
                        idx.append(int(num) - 1 if num else -1)
                    # std::getline drops a trailing empty field: "1//" cannot occur in valid OBJ
                tv.append([vs[idx[6]], vs[idx[3]], vs[idx[0]]])
                tn.append([vns[idx[8]], vns[idx[5]], vns[idx[2]]])
    return np.array(tv, np.float32).reshape(-1, 3, 3), np.array(tn, np.float32).reshape(-1, 3, 3)


def make_blob_mesh(path, subdivisions=2, seed=1984):
    """Procedural stand-in for the missing objs/blender_monkey.obj (SURVEY.md §7 hard part
    6): a displaced, subdivided icosphere (subdivisions=2 -> 320 faces, 3 -> 1280) written
    in the `v` / `vn` / `f a//n b//n c//n` form the reference parser accepts."""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    verts = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
             (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    verts = [np.array(v, np.float64) / np.linalg.norm(v) for v in verts]
    faces = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2),
             (10, 7, 6), (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11),
             (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    for _ in range(subdivisions):
        cache, nf = {}, []

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                m = verts[a] + verts[b]
                verts.append(m / np.linalg.norm(m))
                cache[key] = len(verts) - 1
            return cache[key]

        for a, b, c in faces:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        faces = nf
    V = np.array(verts)
    return _write_displaced_mesh(path, V, faces, seed)


def _write_displaced_mesh(path, V, faces, seed):
    """unit-sphere vertices V displaced radially by a smooth seeded field, area-weighted vertex normals, written
    in the `v` / `vn` / `f a//n b//n c//n` form the reference parser accepts"""
    rng = np.random.Generator(np.random.Philox(seed))
    k = rng.normal(size=(4, 3))
    disp = 1.0 + 0.12 * np.sin(3.0 * V @ k[0]) + 0.08 * np.sin(5.0 * V @ k[1] + 1.0) + 0.05 * np.cos(7.0 * V @ k[2])
    P = V * disp[:, None]
    N = np.zeros_like(P)
    for a, b, c in faces:
        fn = np.cross(P[b] - P[a], P[c] - P[a])
        N[a] += fn
        N[b] += fn
        N[c] += fn
    N /= np.linalg.norm(N, axis=1)[:, None]
    with open(path, "w") as fh:
        fh.write("# procedural blob mesh (stand-in asset), %d faces\n" % len(faces))
        for p in P:
            fh.write("v %.6f %.6f %.6f\n" % tuple(p))
        for n in N:
            fh.write("vn %.6f %.6f %.6f\n" % tuple(n))
        for a, b, c in faces:
            fh.write("f %d//%d %d//%d %d//%d\n" % (a + 1, a + 1, b + 1, b + 1, c + 1, c + 1))
    return len(faces)


def make_blob_mesh_968(path, seed=1984):
    """The config-3 mesh at the triangle count of the reference's asset (objs/blender_monkey.obj, Suzanne: 968
    triangles): a displaced UV sphere, 22 slices x 23 stacks = 2 * 22 * 22 = 968 faces, unit
    bounding radius before displacement."""
    slices, stacks = 22, 23
    verts = [np.array([0.0, 1.0, 0.0])]
    for i in range(1, stacks):
        th = np.pi * i / stacks
        for j in range(slices):
            ph = 2.0 * np.pi * j / slices
            verts.append(np.array([np.sin(th) * np.cos(ph), np.cos(th), np.sin(th) * np.sin(ph)]))
    verts.append(np.array([0.0, -1.0, 0.0]))
    ring = lambda i, j: 1 + (i - 1) * slices + (j % slices)  # ring i = 1 .. stacks-1
    faces = []
    for j in range(slices):
        faces.append((0, ring(1, j + 1), ring(1, j)))
    for i in range(1, stacks - 1):
        for j in range(slices):
            a, b, c, d = ring(i, j), ring(i, j + 1), ring(i + 1, j), ring(i + 1, j + 1)
            faces += [(a, b, d), (a, d, c)]
    last = len(verts) - 1
    for j in range(slices):
        faces.append((last, ring(stacks - 1, j), ring(stacks - 1, j + 1)))
    assert len(faces) == 968
    return _write_displaced_mesh(path, np.array(verts), faces, seed)


def bake_instance(tv, scale=2.5, angle_deg=30.0, offset=(0.0, 1.5, 0.0)):
    """translate(rotate_y(triangle(v*scale), angle), offset) — obj_render.cu:498-511 —
    baked into the vertices. rotate_y::hit maps object p to world as
    x' = cos*x + sin*z, z' = -sin*x + cos*z (rt_next_week/cuda/hittable.h:176-180);
    translate then adds the offset (:73). float32 throughout."""
    f = np.float32
    rad = f(angle_deg) * f(3.1415926535897932385) / f(180.0)
    s, c = f(np.sin(rad, dtype=f)), f(np.cos(rad, dtype=f))
    v = (np.asarray(tv, f) * f(scale)).astype(f)
    x = (c * v[..., 0] + s * v[..., 2]).astype(f)
    z = (-s * v[..., 0] + c * v[..., 2]).astype(f)
    out = np.stack([x, v[..., 1], z], -1).astype(f)
    return (out + np.asarray(offset, f)).astype(f)


def rotate_normals(tn, angle_deg=30.0):
    f = np.float32
    rad = f(angle_deg) * f(3.1415926535897932385) / f(180.0)
    s, c = f(np.sin(rad, dtype=f)), f(np.cos(rad, dtype=f))
    n = np.asarray(tn, f)
    x = (c * n[..., 0] + s * n[..., 2]).astype(f)
    z = (-s * n[..., 0] + c * n[..., 2]).astype(f)
    return np.stack([x, n[..., 1], z], -1).astype(f)


def _group(spheres=(0, 0), triangles=(0, 0), quads=(0, 0)):
    """rt_group: (first, count) ranges of the scene's primitive arrays, in object space"""
    g = np.zeros((), GROUP_DT)
    g["first_sphere"], g["n_spheres"] = spheres
    g["first_triangle"], g["n_triangles"] = triangles
    g["first_quad"], g["n_quads"] = quads
    return g


def _instance(group, angle_deg=0.0, offset=(0.0, 0.0, 0.0), sin_cos=None):
    """rt_instance: translate(rotate_y(object, angle), offset) as a rigid 3x4 matrix"""
    i = np.zeros((), INSTANCE_DT)
    i["m"] = rigid_y(angle_deg, offset, sin_cos)
    i["group"] = group
    return i


def wrap_per_primitive(scene, inst):
    """The reference's way of instancing, as in triangles/cuda/obj_render.cu:498-511: EVERY object is wrapped in
    its own translate(rotate_y(...)). inst: per object in flat list order (spheres, triangles, quads) a record
    with fields flag, sin_y, cos_y, offset. Consecutive primitives of one type that
    share a transform become one group + one instance (the two-level form of the same scene); primitives with
    flag 0 stay at the world level. Returns a copy of `scene` with groups / instances set."""
    ns, nt, nq = len(scene.spheres), len(scene.triangles), len(scene.quads)
    groups, instances = [], []
    for lo, n, field in ((0, ns, "spheres"), (ns, nt, "triangles"), (ns + nt, nq, "quads")):
        k = 0
        while k < n:
            if not inst["flag"][lo + k]:
                k += 1
                continue
            key = (float(inst["sin_y"][lo + k]), float(inst["cos_y"][lo + k]), tuple(float(x) for x in inst["offset"][lo + k]))
            e = k + 1
            while e < n and inst["flag"][lo + e] and (float(inst["sin_y"][lo + e]), float(inst["cos_y"][lo + e]),
                                                       tuple(float(x) for x in inst["offset"][lo + e])) == key:
                e += 1
            groups.append(_group(**{field: (k, e - k)}))
            instances.append(_instance(len(groups) - 1, 0.0, key[2], sin_cos=key[:2]))
            k = e
    out = scene.with_camera(scene.camera)
    out.groups = np.array(groups, GROUP_DT) if groups else np.zeros(0, GROUP_DT)
    out.instances = np.array(instances, INSTANCE_DT) if instances else np.zeros(0, INSTANCE_DT)
    return out


def obj_room(obj_path=None, width=800, height=800, subdivisions=2, mesh=None, instanced=False):
    """Config 3 — triangles/cuda/obj_render.cu:384-524 (obj_model) with camera
    :716-724,736-738: (1,3,7)->(0,2,0), vfov 60, aperture 0, black background,
    t_min 1e-5 (:33), flipping normals (include/hittable.h:29).
    instanced: the mesh stays in OBJECT space (scaled vertices, as triangle(v*scale) stores them) as one rt_group
    placed by one rt_instance = the translate(rotate_y(., 30), (0, 1.5, 0)) the reference wraps around every
    triangle (obj_render.cu:498-511) - the two-level BVH instead of baked vertices."""
    if obj_path is None and mesh == "blob968":  # the bench's config 3: Suzanne's triangle count
        obj_path = os.path.join(DATA_DIR, "blob_968.obj")
        if not os.path.exists(obj_path):
            make_blob_mesh_968(obj_path)
    if obj_path is None:
        obj_path = os.path.join(DATA_DIR, "blob_%d.obj" % subdivisions)
        if not os.path.exists(obj_path):
            make_blob_mesh(obj_path, subdivisions)
    tv, tn = read_obj_triangles(obj_path)
    mats, spheres, quads = [], [], []

    def M(m):
        mats.append(m)
        return len(mats) - 1

    c256 = lambda r, g, b: (np.float32(r) / np.float32(256.0), np.float32(g) / np.float32(256.0),
                            np.float32(b) / np.float32(256.0))
    blue_1 = M(_mat(RT_MAT_LAMBERTIAN, c256(0, 129, 167)))
    red_1 = M(_mat(RT_MAT_LAMBERTIAN, c256(240, 113, 103)))
    yellow_1 = M(_mat(RT_MAT_LAMBERTIAN, c256(253, 252, 220)))
    gold = M(_mat(RT_MAT_METAL, c256(255, 215, 0), 0.5))
    light = M(_mat(RT_MAT_DIFFUSE_LIGHT, (5.0, 5.0, 5.0)))  # color(20,20,20)*0.25
    pink = M(_mat(RT_MAT_DIFFUSE_LIGHT, tuple(np.float32(2) * np.array(c256(255, 59, 148), np.float32))))
    green = M(_mat(RT_MAT_DIFFUSE_LIGHT, tuple(np.float32(2) * np.array(c256(166, 253, 41), np.float32))))
    mirror = lambda: M(_mat(RT_MAT_METAL, (0.8, 0.8, 0.9), 0.0))

    def quad(axis, a0, a1, b0, b1, k, m):
        q = np.zeros((), QUAD_DT)
        q["axis"], q["a0"], q["a1"], q["b0"], q["b1"], q["k"], q["material"] = axis, a0, a1, b0, b1, k, m
        quads.append(q)

    f = np.float32
    spheres.append(_sphere((-1, f(3.69) + f(1), -2.5), 0.3, pink))
    spheres.append(_sphere((1, f(3.69) + f(1), -2.5), 0.3, green))
    XZ, XY, YZ = 1, 2, 0
    quad(XZ, -4, 4, 3, 4, 4 + 1 - 0.01, light)
    quad(XZ, -4, 4, 2, 3, -4 + 0.01, light)
    quad(XY, -4, 4, -4, 4 + 1, -4, yellow_1)
    quad(XY, -3, 3, -4, 4 + 1, -3.999, mirror())
    quad(XZ, -40, 40, -40, 40, -4, red_1)
    quad(XZ, -40, 40, -40, 40, 4 + 1, red_1)
    quad(YZ, -4, 4 + 1, -4, 4, -4, blue_1)
    quad(YZ, -4, 4 + 1, -4, 4, 4, blue_1)
    quad(YZ, -1, 3 + 1, -4, 4, -3.999, mirror())
    quad(YZ, -1, 3 + 1 - 0.001, -4, 4, 3.999, mirror())
    if instanced:
        wv = (np.asarray(tv, np.float32) * np.float32(2.5)).astype(np.float32)
        wn = np.asarray(tn, np.float32)
    else:
        wv = bake_instance(tv)
        wn = rotate_normals(tn)
    tris = np.zeros(len(wv), TRIANGLE_DT)
    for i in range(len(wv)):
        tris[i] = triangle_record(wv[i, 0], wv[i, 1], wv[i, 2], wn[i, 0], wn[i, 1], wn[i, 2], gold)
    sc = Scene(spheres=np.array(spheres, SPHERE_DT), triangles=tris, quads=np.array(quads, QUAD_DT),
               materials=np.array(mats, MATERIAL_DT), name="obj_room")
    if instanced:
        sc.groups = np.array([_group(triangles=(0, len(tris)))], GROUP_DT)
        sc.instances = np.array([_instance(0, 30.0, (0.0, 1.5, 0.0))], INSTANCE_DT)
    lookfrom = np.array([1, 3, 7], np.float32)
    lookat = np.array([0, 2, 0], np.float32)
    dv = lookfrom - lookat
    focus = np.sqrt(np.dot(dv, dv), dtype=np.float32)
    sc.camera = camera_from_lookat(lookfrom, lookat, (0, 1, 0), 60.0, np.float32(width) / np.float32(height), 0.0,
                                   focus, 0.0, 1.0, dtype=np.float32)
    sc.background = (0.0, 0.0, 0.0)
    sc.sky_gradient = 0
    sc.t_min = 1e-5
    sc.flags = RT_FLAG_FLIP_NORMALS | RT_FLAG_DEPTH_BACKGROUND
    sc.profile = RT_PROFILE_NEXT_WEEK
    return sc


def sphere_field(n_side=100, seed=1984, width=600, height=400):
    """Scaling-study generator (SURVEY.md §8d "fallback generator"): the weekend scene's law
    (main.cpp:92-116) on an n_side x n_side grid -> ~n_side^2 small spheres + ground + 3 big
    ones; camera pulled back proportionally. Profile 0."""
    rng = np.random.Generator(np.random.Philox(seed))
    half = n_side // 2
    rows = [[0, -1000, 0, 1000, 0, 0.5, 0.5, 0.5, 0, 0, 0, 0]]
    for a in range(-half, n_side - half):
        for b in range(-half, n_side - half):
            choose = rng.random()
            c = (a + 0.9 * rng.random(), 0.2, b + 0.9 * rng.random())
            if np.hypot(c[0] - 4, c[2]) <= 0.9:
                continue
            if choose < 0.8:
                alb = rng.random(3) * rng.random(3)
                rows.append([*c, 0.2, 0, *alb, 0, 0, 0, 0])
            elif choose < 0.95:
                rows.append([*c, 0.2, 1, *(0.5 + 0.5 * rng.random(3)), 0.5 * rng.random(), 0, 0, 0])
            else:
                rows.append([*c, 0.2, 2, 0, 0, 0, 1.5, 0, 0, 0])
    rows.append([0, 1, 0, 1.0, 2, 0, 0, 0, 1.5, 0, 0, 0])
    rows.append([-4, 1, 0, 1.0, 0, 0.4, 0.2, 0.1, 0, 0, 0, 0])
    rows.append([4, 1, 0, 1.0, 1, 0.7, 0.6, 0.5, 0.0, 0, 0, 0])
    sc = scene_from_rows(np.array(rows, np.float64), "sphere_field_%d" % n_side)
    s = n_side / 22.0
    sc.camera = camera_from_lookat((13 * s, 2 * s, 3 * s), (0, 0, 0), (0, 1, 0), 20.0, width / height, 0.1,
                                   10.0 * s)
    sc.profile = RT_PROFILE_WEEKEND_CPU
    return sc


def _rot_y_f32(pts, angle_deg):
    f = np.float32
    rad = np.deg2rad(np.float64(angle_deg))
    c, s = f(np.cos(rad)), f(np.sin(rad))
    p = np.asarray(pts, f)
    return np.stack([(c * p[..., 0] + s * p[..., 2]).astype(f), p[..., 1], (-s * p[..., 0] + c * p[..., 2]).astype(f)], -1)


def box_as_triangles(p0, p1, angle_deg, offset, material):
    """translate(rotate_y(box(p0,p1), angle), offset) — rt_next_week/cuda/box.h:41-58 with
    main.cu:269-275: the six rects, rotated, are no longer axis aligned and become two
    triangles each; the face normal is the rect's nominal +axis normal (never flipped in the
    rt_next_week tree), rotated. Same construction as include/rtx/rtx.h aa_rect::flatten."""
    f = np.float32
    sides = [(2, p0[0], p1[0], p0[1], p1[1], p1[2]), (2, p0[0], p1[0], p0[1], p1[1], p0[2]),
             (1, p0[0], p1[0], p0[2], p1[2], p1[1]), (1, p0[0], p1[0], p0[2], p1[2], p0[1]),
             (0, p0[1], p1[1], p0[2], p1[2], p1[0]), (0, p0[1], p1[1], p0[2], p1[2], p0[0])]
    out = []
    for axis, a0, a1, b0, b1, k in sides:
        ia, ib = (1 if axis == 0 else 0), (1 if axis == 2 else 2)
        corners = np.zeros((4, 3), f)
        # counter-clockwise seen from the +axis side (the triangle's edge tests follow the vertex order,
        # triangle.h:172-202): (a, b) = (x, z) is left-handed about +y, so the xz rect runs the other way round
        order = ((a0, b0), (a0, b1), (a1, b1), (a1, b0)) if axis == 1 else ((a0, b0), (a1, b0), (a1, b1), (a0, b1))
        for q, (a, b) in enumerate(order):
            corners[q, axis], corners[q, ia], corners[q, ib] = k, a, b
        w = (_rot_y_f32(corners, angle_deg) + np.asarray(offset, f)).astype(f)
        n = np.zeros(3, f)
        n[axis] = 1
        nw = _rot_y_f32(n, angle_deg)
        for t in ((0, 1, 2), (0, 2, 3)):
            out.append(triangle_record(w[t[0]], w[t[1]], w[t[2]], -nw, -nw, -nw, material))
    return out


def cornell_box(width=600, height=600, instanced=False):
    """rt_next_week/cuda/main.cu:252-281,436-443: Cornell box with two rotated boxes, one
    area light, black background; profile 2 (rt_next_week tree: normals never flip).
    instanced: the two boxes stay axis-aligned rects in object space (box.h:41-58), each an rt_group placed by
    an rt_instance = translate(rotate_y(box, angle), offset) (main.cu:269-275)."""
    mats = [_mat(RT_MAT_LAMBERTIAN, (.65, .05, .05)), _mat(RT_MAT_LAMBERTIAN, (.73, .73, .73)),
            _mat(RT_MAT_LAMBERTIAN, (.12, .45, .15)), _mat(RT_MAT_DIFFUSE_LIGHT, (15, 15, 15))]
    red, white, green, light = 0, 1, 2, 3
    quads = []

    def quad(axis, a0, a1, b0, b1, k, m):
        q = np.zeros((), QUAD_DT)
        q["axis"], q["a0"], q["a1"], q["b0"], q["b1"], q["k"], q["material"] = axis, a0, a1, b0, b1, k, m
        quads.append(q)

    quad(0, 0, 555, 0, 555, 555, green)
    quad(0, 0, 555, 0, 555, 0, red)
    quad(1, 213, 343, 227, 332, 554, light)
    quad(1, 0, 555, 0, 555, 0, white)
    quad(1, 0, 555, 0, 555, 555, white)
    quad(2, 0, 555, 0, 555, 555, white)
    groups, instances = [], []
    if instanced:
        tris = []
        for p1, angle, off in (((165, 330, 165), 15, (265, 0, 295)), ((165, 165, 165), -18, (130, 0, 65))):
            groups.append(_group(quads=(len(quads), 6)))
            instances.append(_instance(len(groups) - 1, angle, off))
            quads += box_as_quads((0, 0, 0), p1, white)
    else:
        tris = box_as_triangles((0, 0, 0), (165, 330, 165), 15, (265, 0, 295), white)
        tris += box_as_triangles((0, 0, 0), (165, 165, 165), -18, (130, 0, 65), white)
    # materials in first-use order as the C++ flattening registers them: green, red, light, white
    order = [green, red, light, white]
    remap = {m: i for i, m in enumerate(order)}
    for q in quads:
        q["material"] = remap[int(q["material"])]
    for t in tris:
        t["material"] = remap[int(t["material"])]
    sc = Scene(quads=np.array(quads, QUAD_DT), triangles=np.array(tris, TRIANGLE_DT) if tris else None,
               materials=np.array([mats[m] for m in order], MATERIAL_DT), name="cornell_box",
               groups=np.array(groups, GROUP_DT) if groups else None,
               instances=np.array(instances, INSTANCE_DT) if instances else None)
    sc.camera = camera_from_lookat((278, 278, -800), (278, 278, 0), (0, 1, 0), 40.0,
                                   np.float32(width) / np.float32(height), 0.0, 800.0, 0.0, 1.0, dtype=np.float32)
    sc.background = (0.0, 0.0, 0.0)
    sc.sky_gradient = 0
    sc.t_min = 1e-3
    sc.profile = RT_PROFILE_NEXT_WEEK
    return sc


# ------------------------------------------------------------------ rt_next_week scenes 3-8
def make_perlin(seed=1984):
    """One `perlin` object — rt_next_week/cuda/perlin.h:9-19: 256 gradient vectors
    random_vec3(-1, 1) (not normalised in this tree) and three permutations produced by
    perlin_generate_perm/permute (:76-100: for i = n-1..1 swap p[i], p[random_int(n)])."""
    rng = np.random.Generator(np.random.Philox(seed))
    t = np.zeros((), PERLIN_DT)
    t["ranvec"] = rng.uniform(-1.0, 1.0, (256, 3)).astype(np.float32)
    for name in ("perm_x", "perm_y", "perm_z"):
        perm = np.arange(256, dtype=np.int32)
        for i in range(255, 0, -1):
            target = min(255, int(rng.random() * 256.0))
            perm[i], perm[target] = perm[target], perm[i]
        t[name] = perm
    return t


def procedural_earth(width=512, height=256, seed=7):
    """Stand-in for earthmap.jpeg (a binary asset of the reference that is not redistributed
    here): an equirectangular map with blue oceans, green/brown continents and white caps,
    uint8 [h][w][3], row 0 = top (north)."""
    rng = np.random.Generator(np.random.Philox(seed))
    lon = (np.arange(width) + 0.5) / width * 2 * np.pi
    lat = (0.5 - (np.arange(height) + 0.5) / height) * np.pi
    lo, la = np.meshgrid(lon, lat)
    x, y, z = np.cos(la) * np.cos(lo), np.sin(la), np.cos(la) * np.sin(lo)
    f = np.zeros_like(x)
    for octave in range(5):
        for _ in range(4):
            k = rng.normal(size=3)
            k *= (2.0 ** octave) * 1.5 / np.linalg.norm(k)
            f += np.sin(k[0] * x + k[1] * y + k[2] * z + rng.uniform(0, 2 * np.pi)) / (2.0 ** octave)
    img = np.zeros((height, width, 3), np.float64)
    land = f > 0.35
    img[~land] = (0.05, 0.15, 0.45)
    h = np.clip((f - 0.35) / 1.5, 0, 1)[..., None]
    img = np.where(land[..., None], (1 - h) * np.array((0.15, 0.45, 0.12)) + h * np.array((0.45, 0.35, 0.2)), img)
    img[np.abs(la) > 1.25] = (0.92, 0.94, 0.96)
    return (np.clip(img, 0, 1) * 255.0 + 0.5).astype(np.uint8)


def _quad(axis, a0, a1, b0, b1, k, m):
    q = np.zeros((), QUAD_DT)
    q["axis"], q["a0"], q["a1"], q["b0"], q["b1"], q["k"], q["material"] = axis, a0, a1, b0, b1, k, m
    return q


def _medium_sphere(center, radius, density, material):
    m = np.zeros((), MEDIUM_DT)
    m["shape"], m["p0"], m["p1"], m["cos_y"], m["density"], m["material"] = 0, center, (radius, 0, 0), 1.0, density, material
    return m


def _medium_box(p0, p1, angle_deg, offset, density, material):
    """constant_medium(translate(rotate_y(box(p0, p1), angle), offset), density, colour)"""
    m = np.zeros((), MEDIUM_DT)
    rad = np.deg2rad(np.float64(angle_deg))
    m["shape"], m["p0"], m["p1"], m["offset"], m["density"], m["material"] = 1, p0, p1, offset, density, material
    m["sin_y"], m["cos_y"] = np.float32(np.sin(rad)), np.float32(np.cos(rad))
    return m


def _nw_camera(sc, lookfrom, lookat, vfov, width, height, aperture=0.0):
    lookfrom = np.asarray(lookfrom, np.float32)
    d = lookfrom - np.asarray(lookat, np.float32)
    focus = np.sqrt(np.dot(d, d), dtype=np.float32)  # main.cu:462
    sc.camera = camera_from_lookat(lookfrom, lookat, (0, 1, 0), vfov, np.float32(width) / np.float32(height), aperture,
                                   focus, 0.0, 1.0, dtype=np.float32)
    sc.sky_gradient = 0
    sc.t_min = 1e-3
    sc.profile = RT_PROFILE_NEXT_WEEK
    return sc


def two_perlin_spheres(width=600, height=400, seed=1984):
    """rt_next_week/cuda/main.cu:212-222,417-422: noise_texture(4) on the ground and a ball."""
    mats = [_mat(RT_MAT_LAMBERTIAN, (1, 1, 1), texture=RT_TEX_NOISE, albedo2=(4.0, 0, 0))]
    spheres = [_sphere((0, -1000, 0), 1000, 0), _sphere((0, 2, 0), 2, 0)]
    sc = Scene(spheres=np.array(spheres, SPHERE_DT), materials=np.array(mats, MATERIAL_DT),
               perlin=np.array([make_perlin(seed)], PERLIN_DT), name="two_perlin_spheres")
    sc.background = (0.70, 0.80, 1.00)
    return _nw_camera(sc, (13, 2, 3), (0, 0, 0), 20.0, width, height)


def earth(width=600, height=400, image=None):
    """rt_next_week/cuda/main.cu:224-232,423-426: one sphere with an image_texture."""
    mats = [_mat(RT_MAT_LAMBERTIAN, (1, 1, 1), texture=RT_TEX_IMAGE, albedo2=(0, 0, 0))]
    sc = Scene(spheres=np.array([_sphere((0, 0, 0), 2, 0)], SPHERE_DT), materials=np.array(mats, MATERIAL_DT),
               images=[procedural_earth() if image is None else image], name="earth")
    sc.background = (0.70, 0.80, 1.00)
    return _nw_camera(sc, (13, 2, 3), (0, 0, 0), 40.0, width, height)


def simple_light(width=600, height=400, seed=1984):
    """rt_next_week/cuda/main.cu:234-250,428-434: perlin spheres lit by a rect and a sphere light."""
    mats = [_mat(RT_MAT_LAMBERTIAN, (1, 1, 1), texture=RT_TEX_NOISE, albedo2=(4.0, 0, 0)),
            _mat(RT_MAT_DIFFUSE_LIGHT, (4, 4, 4)), _mat(RT_MAT_DIFFUSE_LIGHT, (6, 4, 4))]
    spheres = [_sphere((0, -1000, 0), 1000, 0), _sphere((0, 2, 0), 2, 0), _sphere((0, 6, 0), 1.5, 2)]
    quads = [_quad(2, 3, 5, 1, 2, -2, 1)]
    sc = Scene(spheres=np.array(spheres, SPHERE_DT), quads=np.array(quads, QUAD_DT), materials=np.array(mats, MATERIAL_DT),
               perlin=np.array([make_perlin(seed)], PERLIN_DT), name="simple_light")
    sc.background = (0.0, 0.0, 0.0)
    return _nw_camera(sc, (26, 3, 6), (0, 2, 0), 20.0, width, height)


def cornell_smoke(width=600, height=600):
    """rt_next_week/cuda/main.cu:283-310,445-451: the Cornell room with the two boxes replaced by
    constant media (density 0.01; black and white smoke)."""
    mats = [_mat(RT_MAT_LAMBERTIAN, (.12, .45, .15)), _mat(RT_MAT_LAMBERTIAN, (.65, .05, .05)),
            _mat(RT_MAT_DIFFUSE_LIGHT, (15, 15, 15)), _mat(RT_MAT_LAMBERTIAN, (.73, .73, .73)),
            _mat(RT_MAT_ISOTROPIC, (0, 0, 0)), _mat(RT_MAT_ISOTROPIC, (1, 1, 1))]
    green, red, light, white = 0, 1, 2, 3
    quads = [_quad(0, 0, 555, 0, 555, 555, green), _quad(0, 0, 555, 0, 555, 0, red),
             _quad(1, 213, 343, 227, 332, 554, light), _quad(1, 0, 555, 0, 555, 0, white),
             _quad(1, 0, 555, 0, 555, 555, white), _quad(2, 0, 555, 0, 555, 555, white)]
    media = [_medium_box((0, 0, 0), (165, 330, 165), 15, (265, 0, 295), 0.01, 4),
             _medium_box((0, 0, 0), (165, 165, 165), -18, (130, 0, 65), 0.01, 5)]
    sc = Scene(quads=np.array(quads, QUAD_DT), materials=np.array(mats, MATERIAL_DT), media=np.array(media, MEDIUM_DT),
               name="cornell_smoke")
    sc.background = (0.0, 0.0, 0.0)
    return _nw_camera(sc, (278, 278, -800), (278, 278, 0), 40.0, width, height)


def box_as_quads(p0, p1, material):
    """box(p0, p1) — rt_next_week/cuda/box.h:41-58: six axis-aligned rects."""
    return [_quad(2, p0[0], p1[0], p0[1], p1[1], p1[2], material), _quad(2, p0[0], p1[0], p0[1], p1[1], p0[2], material),
            _quad(1, p0[0], p1[0], p0[2], p1[2], p1[1], material), _quad(1, p0[0], p1[0], p0[2], p1[2], p0[1], material),
            _quad(0, p0[1], p1[1], p0[2], p1[2], p1[0], material), _quad(0, p0[1], p1[1], p0[2], p1[2], p0[0], material)]


def next_week_final(width=800, height=800, seed=1984, image=None, instanced=False):
    """rt_next_week/cuda/main.cu:312-383,453-459 — the tree's default scene: 400 ground boxes, an
    area light, a moving sphere, glass and metal balls, a blue subsurface ball (dielectric
    boundary + dense medium), thin global fog, an image-textured and a perlin ball, and a
    rotated, translated cluster of 1000 small spheres.
    instanced: the cluster stays in object space ([0,165)^3) as one rt_group under one rt_instance
    (main.cu:373-381: translate(rotate_y(bvh_node(boxes2), 15), vec3(-100, 270, 395)))."""
    rng = np.random.Generator(np.random.Philox(seed))
    f = np.float32
    mats = [_mat(RT_MAT_LAMBERTIAN, (0.48, 0.83, 0.53)),  # 0 ground
            _mat(RT_MAT_DIFFUSE_LIGHT, (7, 7, 7)),          # 1 light
            _mat(RT_MAT_LAMBERTIAN, (0.7, 0.3, 0.1)),       # 2 moving sphere
            _mat(RT_MAT_DIELECTRIC, param=1.5),             # 3 glass
            _mat(RT_MAT_METAL, (0.8, 0.8, 0.9), 1.0),       # 4 metal
            _mat(RT_MAT_ISOTROPIC, (0.2, 0.4, 0.9)),        # 5 blue medium
            _mat(RT_MAT_ISOTROPIC, (1, 1, 1)),              # 6 fog
            _mat(RT_MAT_LAMBERTIAN, (1, 1, 1), texture=RT_TEX_IMAGE, albedo2=(0, 0, 0)),    # 7 earth
            _mat(RT_MAT_LAMBERTIAN, (1, 1, 1), texture=RT_TEX_NOISE, albedo2=(0.1, 0, 0)),  # 8 perlin
            _mat(RT_MAT_LAMBERTIAN, (.73, .73, .73))]       # 9 white
    quads = []
    for i in range(20):
        for j in range(20):
            w = f(100.0)
            x0, z0 = f(-1000.0) + f(i) * w, f(-1000.0) + f(j) * w
            y1 = f(1.0 + 100.0 * rng.random())
            quads += box_as_quads((x0, 0.0, z0), (x0 + w, y1, z0 + w), 0)
    quads.append(_quad(1, 123, 423, 147, 412, 554, 1))
    spheres = [_sphere((400, 400, 200), 50, 2, c1=(430, 400, 200), t0=0.0, t1=1.0),
               _sphere((260, 150, 45), 50, 3), _sphere((0, 150, 145), 50, 4),
               _sphere((360, 150, 145), 70, 3),
               _sphere((400, 200, 400), 100, 7), _sphere((220, 280, 300), 80, 8)]
    media = [_medium_sphere((360, 150, 145), 70, 0.2, 5), _medium_sphere((0, 0, 0), 5000, 0.0001, 6)]
    # cluster: translate(rotate_y(bvh(1000 spheres in [0,165)^3, r = 10), 15), (-100, 270, 395))
    centres = rng.uniform(0.0, 165.0, (1000, 3)).astype(f)
    groups = instances = None
    if instanced:
        groups = np.array([_group(spheres=(len(spheres), 1000))], GROUP_DT)
        instances = np.array([_instance(0, 15.0, (-100.0, 270.0, 395.0))], INSTANCE_DT)
    else:
        centres = (_rot_y_f32(centres, 15) + np.array((-100, 270, 395), f)).astype(f)
    spheres += [_sphere(tuple(c), 10, 9) for c in centres]
    sc = Scene(spheres=np.array(spheres, SPHERE_DT), quads=np.array(quads, QUAD_DT), materials=np.array(mats, MATERIAL_DT),
               media=np.array(media, MEDIUM_DT), perlin=np.array([make_perlin(seed)], PERLIN_DT),
               images=[procedural_earth() if image is None else image], name="next_week_final",
               groups=groups, instances=instances)
    sc.background = (0.0, 0.0, 0.0)
    return _nw_camera(sc, (478, 278, -600), (278, 278, 0), 40.0, width, height)


# ------------------------------------------------------------------ text scene files
def load_scene_file(path, width=None, height=None):
    """Parse a text scene file (format: include/rtx/scene_file.h) with the C++ host layer
    (build/librtx_host.so — g++ only, no CUDA) and return the flattened Scene plus the file's
    default (width, height, spp)."""
    import ctypes as C
    import subprocess

    from .ctypes_defs import RtCamera
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call(["make", "-C", os.path.join(root, "apps"), "-s", "../build/librtx_host.so"])
    L = C.CDLL(os.path.join(root, "build", "librtx_host.so"))
    vp = C.c_void_p
    L.rtx_host_load_scene_file.argtypes = [C.c_char_p, C.c_int, C.c_int, vp, vp, C.c_char_p, C.c_int]
    L.rtx_host_counts.argtypes = [vp]
    L.rtx_host_get.argtypes = [vp, vp, vp, vp, vp]
    L.rtx_host_counts2.argtypes = [vp, vp]
    L.rtx_host_get2.argtypes = [vp, vp, vp]
    opts = np.zeros(8, np.int32)
    fopts = np.zeros(4, np.float32)
    err = C.create_string_buffer(512)
    if L.rtx_host_load_scene_file(os.fsencode(path), width or 0, height or 0, opts.ctypes.data, fopts.ctypes.data, err, 512):
        raise ValueError(err.value.decode())
    n = np.zeros(5, np.int32)
    L.rtx_host_counts(n.ctypes.data)
    sph, tri = np.zeros(n[0], SPHERE_DT), np.zeros(n[1], TRIANGLE_DT)
    quad, mats = np.zeros(n[2], QUAD_DT), np.zeros(n[3], MATERIAL_DT)
    cam = RtCamera()
    L.rtx_host_get(sph.ctypes.data, tri.ctypes.data, quad.ctypes.data, mats.ctypes.data, C.addressof(cam))
    n2 = np.zeros(3, np.int32)
    dims = np.zeros(64, np.int32)
    L.rtx_host_counts2(n2.ctypes.data, dims.ctypes.data)
    media, perlin = np.zeros(n2[0], MEDIUM_DT), np.zeros(n2[1], PERLIN_DT)
    images = [np.zeros((dims[2 * i + 1], dims[2 * i], 3), np.uint8) for i in range(n2[2])]
    ptrs = (C.c_void_p * max(len(images), 1))(*[im.ctypes.data for im in images])
    L.rtx_host_get2(media.ctypes.data, perlin.ctypes.data, C.addressof(ptrs))
    n3 = np.zeros(2, np.int32)  # `flags instancing`: objects under an XFORM arrive as groups + instances
    L.rtx_host_counts3.argtypes = [vp]
    L.rtx_host_get3.argtypes = [vp, vp]
    L.rtx_host_counts3(n3.ctypes.data)
    groups, instances = np.zeros(n3[0], GROUP_DT), np.zeros(n3[1], INSTANCE_DT)
    L.rtx_host_get3(groups.ctypes.data, instances.ctypes.data)
    sc = Scene(spheres=sph, triangles=tri, quads=quad, materials=mats, camera=cam, background=tuple(fopts[:3]),
               sky_gradient=int(opts[1]), t_min=float(fopts[3]), max_depth=int(opts[2]), flags=int(opts[3]),
               name=os.path.basename(path), profile=int(opts[0]), media=media, perlin=perlin, images=images,
               groups=groups, instances=instances)
    return sc, (int(opts[4]), int(opts[5]), int(opts[6]))
