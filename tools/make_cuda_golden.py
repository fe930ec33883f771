#!/usr/bin/env python
"""Golden vectors from the reference's UNMODIFIED CUDA-tree device code (oracle/ref_cuda_harness.cu).

Three steps (the GPU box has no /root/reference, this container has no GPU):
  1. here:        make -C oracle refcuda                      -> oracle/_ref/ref_cuda_{nw,tri}[_nofma]
                  python tools/make_cuda_golden.py inputs     -> build/cuda_ref/<case>.in
  2. GPU box:     python tools/make_cuda_golden.py run        -> gpurun_out/cuda_ref/<case>.<variant>.out
  3. here:        python tools/make_cuda_golden.py collect    -> tests/golden/cuda_ref_<case>.npz

Each .npz holds the flattened scene, the rays and what the reference computed for them: closest hit
over its hittable_list and through its bvh_node (object, t, p, normal, u, v, front_face), material::scatter
+ emitted at that hit (with the ray's own cuRAND XORWOW stream), constant_medium::hit per medium,
get_color per ray, texture values, the perlin tables its constructor drew, the sin/cos its rotate_y
computed, and known answers of the generator. Two builds of the same sources: "fma" = nvcc's defaults
(what the reference Makefiles use) and "nofma" = -fmad=false (rounds like the C restatement).
"""
import os
import struct
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from a_dive_into_ray_tracing_b200 import ctypes_defs as D, scenes  # noqa: E402

IN_DIR = os.path.join(ROOT, "build", "cuda_ref")
OUT_DIR = os.path.join(ROOT, "gpurun_out", "cuda_ref")
GOLDEN = os.path.join(ROOT, "tests", "golden")
FLT_MAX = 3.4028234663852886e38

TRI_IN_DT = np.dtype([("v0", "<f4", 3), ("v1", "<f4", 3), ("v2", "<f4", 3), ("vn0", "<f4", 3), ("vn1", "<f4", 3),
                      ("vn2", "<f4", 3), ("material", "<i4"), ("pad", "<i4")])
INST_DT = np.dtype([("angle_deg", "<f4"), ("off", "<f4", 3), ("flag", "<i4")])
MEDIUM_IN_DT = np.dtype([("shape", "<i4"), ("p0", "<f4", 3), ("p1", "<f4", 3), ("angle_deg", "<f4"), ("off", "<f4", 3),
                         ("density", "<f4"), ("albedo", "<f4", 3)])
TEXQ_DT = np.dtype([("material", "<i4"), ("u", "<f4"), ("v", "<f4"), ("p", "<f4", 3)])
HIT_DT = np.dtype([("hit", "<i4"), ("obj", "<i4"), ("t", "<f4"), ("p", "<f4", 3), ("n", "<f4", 3), ("u", "<f4"),
                   ("v", "<f4"), ("front_face", "<i4")])
SCATTER_DT = np.dtype([("ok", "<i4"), ("att", "<f4", 3), ("o", "<f4", 3), ("d", "<f4", 3), ("tm", "<f4"),
                       ("emitted", "<f4", 3), ("draws", "<i4")])
MEDHIT_DT = np.dtype([("hit", "<i4"), ("t", "<f4"), ("p", "<f4", 3), ("draws", "<i4")])
HEADER_FMT = "<12i8fQ"
assert struct.calcsize(HEADER_FMT) == 88 and HIT_DT.itemsize == 48 and SCATTER_DT.itemsize == 60


class Case:
    def __init__(self, name, tree, scene, rays, tri_in=None, inst=None, media_in=None, texq=None, n_color=0, seed=1984,
                 t_min=None):
        self.name, self.tree, self.scene, self.seed = name, tree, scene, seed
        self.rays = np.ascontiguousarray(rays, np.float32)
        n_prims = scene.n_prims
        self.tri_in = tri_in if tri_in is not None else np.zeros(0, TRI_IN_DT)
        assert len(self.tri_in) == len(scene.triangles)
        self.inst = inst if inst is not None else np.zeros(n_prims, INST_DT)
        assert len(self.inst) == n_prims
        self.media_in = media_in if media_in is not None else np.zeros(0, MEDIUM_IN_DT)
        assert len(self.media_in) == len(scene.media)
        self.texq = texq if texq is not None else np.zeros(0, TEXQ_DT)
        self.n_color = n_color
        self.t_min = scene.t_min if t_min is None else t_min

    def write(self, path):
        sc = self.scene
        with open(path, "wb") as fh:
            fh.write(struct.pack(HEADER_FMT, 0x31484352, self.tree, len(sc.spheres), len(sc.triangles), len(sc.quads),
                                 len(sc.materials), len(sc.images), len(sc.media), len(self.rays), len(self.texq),
                                 self.n_color, 0, self.t_min, FLT_MAX, *sc.background, sc.camera.time0, sc.camera.time1, 0.0,
                                 self.seed))
            for a in (sc.spheres, self.tri_in, sc.quads, sc.materials, self.inst, self.media_in):
                fh.write(a.tobytes())
            wh = np.array([[im.shape[1], im.shape[0]] for im in sc.images], np.int32).reshape(-1)
            fh.write(wh.tobytes())
            for im in sc.images:
                fh.write(im.tobytes())
            fh.write(self.texq.tobytes())
            fh.write(self.rays.tobytes())


def read_output(path, case):
    with open(path, "rb") as fh:
        fh.read(88)

        def arr(dt):
            (nbytes,) = struct.unpack("<Q", fh.read(8))
            return np.frombuffer(fh.read(nbytes), dt).copy()

        o = {"list": arr(HIT_DT), "bvh": arr(HIT_DT), "scatter": arr(SCATTER_DT), "media": arr(MEDHIT_DT),
             "color": arr(np.float32).reshape(-1, 2, 4), "tex": arr(np.float32).reshape(-1, 3),
             "perlin": arr(np.float32).reshape(-1, 1536), "xorwow": arr(np.float32).reshape(16, 8),
             "inst_sc": arr(np.float32).reshape(-1, 2), "med_sc": arr(np.float32).reshape(-1, 2)}
    o["media"] = o["media"].reshape(len(case.rays), -1) if len(case.scene.media) else o["media"]
    return o


# ------------------------------------------------------------------ rays
def surface_rays(sc, n, rng, inside_frac=0.15):
    """rays leaving the surfaces of random spheres (as scattered rays do: origin ON the surface, random
    direction in the outer hemisphere; a share starts inside and points anywhere) + random times"""
    sp = sc.spheres[np.abs(sc.spheres["radius"]) < 100]
    k = rng.integers(0, len(sp), n)
    nrm = rng.normal(size=(n, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    tm = rng.random(n).astype(np.float32)
    c = sp["center0"][k].astype(np.float64)
    mv = sp["moving"][k] != 0
    c = c + mv[:, None] * tm[:, None] * (sp["center1"][k].astype(np.float64) - c)
    r = np.abs(sp["radius"][k]).astype(np.float64)
    o = c + nrm * r[:, None] * (1 + 2e-4)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    flip = (np.sum(d * nrm, axis=1) < 0)
    d[flip] *= -1
    ins = rng.random(n) < inside_frac
    o[ins] = c[ins] + nrm[ins] * r[ins, None] * rng.random((ins.sum(), 1)) * 0.9
    d[ins] = rng.normal(size=(ins.sum(), 3))
    d *= rng.uniform(0.3, 2.0, (n, 1))
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:3], rays[:, 3], rays[:, 4:7] = o, tm, d
    return rays


def box_rays(lo, hi, n, rng):
    o = np.asarray(lo) + (np.asarray(hi) - np.asarray(lo)) * rng.random((n, 3))
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d *= rng.uniform(0.3, 2.0, (n, 1))
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:3], rays[:, 3], rays[:, 4:7] = o, rng.random(n), d
    return rays


def camera_rays(sc, W, H, n, rng):
    px = rng.choice(W * H, size=min(n, W * H), replace=False)
    rays = D.primary_rays(sc.camera, W, H, sc.profile, s_jitter=0.37, t_jitter=0.61, pixels=np.sort(px))
    rays[:, 3] = rng.random(len(rays))
    return rays


def media_in_from(sc):
    """rt_medium (sin/cos) -> the harness's constructor arguments (angle in degrees, albedo colour)"""
    out = np.zeros(len(sc.media), MEDIUM_IN_DT)
    for k, m in enumerate(sc.media):
        out[k]["shape"], out[k]["p0"], out[k]["p1"], out[k]["off"], out[k]["density"] = (m["shape"], m["p0"], m["p1"],
                                                                                            m["offset"], m["density"])
        out[k]["angle_deg"] = np.float32(np.rad2deg(np.arctan2(np.float64(m["sin_y"]), np.float64(m["cos_y"]))))
        out[k]["albedo"] = sc.materials[m["material"]]["albedo"]
    return out


# ------------------------------------------------------------------ cases
def case_nw_spheres():
    """config 4: moving / static spheres, checker ground, metal, glass (rt_next_week/cuda/main.cu:153-198)"""
    rng = np.random.default_rng(41)
    sc = scenes.next_week(200, 133)
    rays = np.concatenate([camera_rays(sc, 200, 133, 1200, rng), surface_rays(sc, 2300, rng)])
    texq = np.zeros(400, TEXQ_DT)
    texq["material"] = int(np.nonzero(sc.materials["texture"] == D.RT_TEX_CHECKER)[0][0])
    texq["p"] = rng.uniform(-12, 12, (400, 3))
    texq["p"][:, 1] = rng.uniform(-0.05, 0.05, 400)
    return Case("nw_spheres", 1, sc, rays, texq=texq, n_color=1500)


def case_nw_cornell_inst():
    """Cornell room (aarect.h, diffuse_light) with the two boxes as translate(rotate_y(box)) INSTANCES
    (main.cu:252-281), a noise-textured and an image-textured sphere (texture.h:55-124) and a metal one"""
    rng = np.random.default_rng(42)
    M = scenes._mat
    mats = [M(D.RT_MAT_LAMBERTIAN, (.12, .45, .15)), M(D.RT_MAT_LAMBERTIAN, (.65, .05, .05)),
            M(D.RT_MAT_DIFFUSE_LIGHT, (15, 15, 15)), M(D.RT_MAT_LAMBERTIAN, (.73, .73, .73)),
            M(D.RT_MAT_LAMBERTIAN, (1, 1, 1), texture=D.RT_TEX_NOISE, albedo2=(0.05, 0, 0)),
            M(D.RT_MAT_LAMBERTIAN, (1, 1, 1), texture=D.RT_TEX_IMAGE, albedo2=(0, 0, 0)),
            M(D.RT_MAT_METAL, (0.8, 0.85, 0.88), 0.3), M(D.RT_MAT_DIELECTRIC, param=1.5)]
    Q = scenes._quad
    quads = [Q(0, 0, 555, 0, 555, 555, 0), Q(0, 0, 555, 0, 555, 0, 1), Q(1, 213, 343, 227, 332, 554, 2),
             Q(1, 0, 555, 0, 555, 0, 3), Q(1, 0, 555, 0, 555, 555, 3), Q(2, 0, 555, 0, 555, 555, 3)]
    n_walls = len(quads)
    quads += scenes.box_as_quads((0, 0, 0), (165, 330, 165), 3) + scenes.box_as_quads((0, 0, 0), (165, 165, 165), 3)
    spheres = [scenes._sphere((400, 90, 120), 60, 4), scenes._sphere((150, 240, 150), 70, 5),
               scenes._sphere((300, 60, 60), 55, 6), scenes._sphere((460, 330, 300), 50, 7)]
    sc = D.Scene(spheres=np.array(spheres, D.SPHERE_DT), quads=np.array(quads, D.QUAD_DT),
                 materials=np.array(mats, D.MATERIAL_DT), perlin=np.array([scenes.make_perlin(1)], D.PERLIN_DT),
                 images=[scenes.procedural_earth(128, 64)], name="nw_cornell_inst")
    sc.background = (0.0, 0.0, 0.0)
    scenes._nw_camera(sc, (278, 278, -800), (278, 278, 0), 40.0, 96, 96)
    inst = np.zeros(sc.n_prims, INST_DT)
    first = len(spheres) + n_walls
    inst[first:first + 6] = np.array([(15.0, (265, 0, 295), 1)], INST_DT)
    inst[first + 6:first + 12] = np.array([(-18.0, (130, 0, 65), 1)], INST_DT)
    rays = np.concatenate([camera_rays(sc, 96, 96, 1200, rng), box_rays((5, 5, 5), (550, 550, 550), 1600, rng),
                           surface_rays(sc, 700, rng, inside_frac=0.1)])
    texq = np.zeros(600, TEXQ_DT)
    texq["material"][:300] = 4
    texq["p"][:300] = rng.uniform(0, 555, (300, 3))
    texq["material"][300:] = 5
    texq["u"][300:] = rng.uniform(-0.05, 1.05, 300)
    texq["v"][300:] = rng.uniform(-0.05, 1.05, 300)
    return Case("nw_cornell_inst", 1, sc, rays, inst=inst, texq=texq, n_color=1500)


def case_nw_media():
    """cornell_smoke (main.cu:283-310): two rotated box media + a sphere medium inside a glass ball"""
    rng = np.random.default_rng(43)
    sc = scenes.cornell_smoke(96, 96)
    mats = list(sc.materials) + [scenes._mat(D.RT_MAT_ISOTROPIC, (0.2, 0.4, 0.9)), scenes._mat(D.RT_MAT_DIELECTRIC, param=1.5)]
    media = list(sc.media) + [scenes._medium_sphere((400, 380, 200), 80, 0.02, 6)]
    sc2 = D.Scene(spheres=np.array([scenes._sphere((400, 380, 200), 80, 7)], D.SPHERE_DT), quads=sc.quads,
                  materials=np.array(mats, D.MATERIAL_DT), media=np.array(media, D.MEDIUM_DT), camera=sc.camera,
                  background=sc.background, sky_gradient=0, t_min=1e-3, profile=D.RT_PROFILE_NEXT_WEEK, name="nw_media")
    rays = np.concatenate([camera_rays(sc2, 96, 96, 1500, rng), box_rays((5, 5, 5), (550, 550, 550), 1500, rng)])
    return Case("nw_media", 1, sc2, rays, media_in=media_in_from(sc2), n_color=1500)


def _room_triangles(baked):
    path = os.path.join(scenes.DATA_DIR, "blob_2.obj")
    tv, tn = scenes.read_obj_triangles(path)
    f = np.float32
    if baked:
        wv, wn = scenes.bake_instance(tv), scenes.rotate_normals(tn)
    else:
        wv, wn = (np.asarray(tv, f) * f(2.5)).astype(f), np.asarray(tn, f)  # obj_render.cu:498-511: vertices * scale
    return wv, wn


def case_tri_room(baked=True):
    """config 3: the mesh in the lit, mirrored room (triangles/cuda/obj_render.cu:384-524); baked = instance
    transform applied to the vertices (what the flattened scene holds), not baked = every triangle wrapped in
    translate(rotate_y(triangle, 30), (0, 1.5, 0)) exactly as obj_render.cu:498-511 does"""
    rng = np.random.default_rng(44)  # the same rays for the baked and the instanced form
    sc = scenes.obj_room(width=96, height=96)
    wv, wn = _room_triangles(baked)
    tri_in = np.zeros(len(wv), TRI_IN_DT)
    tris = sc.triangles.copy()
    for i in range(len(wv)):
        tri_in[i] = (wv[i, 0], wv[i, 1], wv[i, 2], wn[i, 0], wn[i, 1], wn[i, 2], tris[i]["material"], 0)
        if not baked:
            tris[i] = scenes.triangle_record(wv[i, 0], wv[i, 1], wv[i, 2], wn[i, 0], wn[i, 1], wn[i, 2], tris[i]["material"])
    sc2 = D.Scene(spheres=sc.spheres, triangles=tris, quads=sc.quads, materials=sc.materials, camera=sc.camera,
                  background=sc.background, sky_gradient=0, t_min=sc.t_min, flags=sc.flags, profile=sc.profile,
                  name="tri_room" if baked else "tri_room_inst")
    inst = np.zeros(sc2.n_prims, INST_DT)
    if not baked:
        inst[len(sc.spheres):len(sc.spheres) + len(tris)] = np.array([(30.0, (0, 1.5, 0), 1)], INST_DT)
    rays = np.concatenate([camera_rays(sc2, 96, 96, 1500, rng), box_rays((-3.9, -3.9, -3.9), (3.9, 4.9, 3.9), 1500, rng)])
    rays[:, 3] = 0.0
    return Case(sc2.name, 2, sc2, rays, tri_in=tri_in, inst=inst, n_color=1500)


def all_cases():
    return [case_nw_spheres(), case_nw_cornell_inst(), case_nw_media(), case_tri_room(True), case_tri_room(False)]


BIN = {1: "ref_cuda_nw", 2: "ref_cuda_tri"}


def main():
    cmd = sys.argv[1] if len(sys.argv) > 1 else "inputs"
    cases = all_cases()
    if cmd == "inputs":
        os.makedirs(IN_DIR, exist_ok=True)
        for c in cases:
            c.write(os.path.join(IN_DIR, c.name + ".in"))
            print("wrote", c.name, len(c.rays), "rays")
    elif cmd == "run":
        os.makedirs(OUT_DIR, exist_ok=True)
        for c in cases:
            for variant, suffix in (("fma", ""), ("nofma", "_nofma")):
                exe = os.path.join(ROOT, "oracle", "_ref", BIN[c.tree] + suffix)
                subprocess.check_call([exe, os.path.join(IN_DIR, c.name + ".in"),
                                       os.path.join(OUT_DIR, "%s.%s.out" % (c.name, variant))])
    elif cmd == "collect":
        for c in cases:
            sc = c.scene
            d = {"tree": c.tree, "seed": c.seed, "t_min": np.float32(c.t_min), "rays": c.rays, "n_color": c.n_color,
                 "spheres": sc.spheres, "triangles": sc.triangles, "quads": sc.quads, "materials": sc.materials,
                 "media": sc.media, "tri_in": c.tri_in, "inst": c.inst, "texq": c.texq, "flags": sc.flags,
                 "background": np.array(sc.background, np.float32), "profile": sc.profile, "max_depth": sc.max_depth,
                 "n_images": len(sc.images)}
            for k, im in enumerate(sc.images):
                d["image%d" % k] = im
            for variant in ("fma", "nofma"):
                o = read_output(os.path.join(OUT_DIR, "%s.%s.out" % (c.name, variant)), c)
                for key, val in o.items():
                    d["%s_%s" % (variant, key)] = val
            path = os.path.join(GOLDEN, "cuda_ref_%s.npz" % c.name)
            np.savez_compressed(path, **d)
            print(path, os.path.getsize(path) >> 10, "KiB")
    else:
        raise SystemExit("usage: make_cuda_golden.py inputs|run|collect")


if __name__ == "__main__":
    main()
