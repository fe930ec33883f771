"""The C++ host layer (include/rtx: the reference's scene-description surface; apps/):
CPU tests of scene construction + flattening against the reference golden scene and the
Python flattening; GPU test of the render_cli program (the reference's main() replacement)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from a_dive_into_ray_tracing_b200 import ctypes_defs as D
from a_dive_into_ray_tracing_b200 import scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_vp = C.c_void_p


@pytest.fixture(scope="module")
def host():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "apps"), "-s", "../build/librtx_host.so"])
    L = C.CDLL(os.path.join(ROOT, "build", "librtx_host.so"))
    L.rtx_host_build.argtypes = [C.c_int, C.c_uint, C.c_double, C.c_char_p]
    L.rtx_host_counts.argtypes = [_vp]
    L.rtx_host_get.argtypes = [_vp, _vp, _vp, _vp, _vp]

    def build(which, seed=1, aspect=1.5, obj=b""):
        assert L.rtx_host_build(which, seed, aspect, obj) == 0
        n = np.zeros(5, np.int32)
        L.rtx_host_counts(n.ctypes.data)
        s = np.zeros(n[0], D.SPHERE_DT)
        t = np.zeros(n[1], D.TRIANGLE_DT)
        q = np.zeros(n[2], D.QUAD_DT)
        m = np.zeros(n[3], D.MATERIAL_DT)
        cam = D.RtCamera()
        L.rtx_host_get(s.ctypes.data, t.ctypes.data, q.ctypes.data, m.ctypes.data, C.addressof(cam))
        return s, t, q, m, cam, bool(n[4])

    return build


def test_weekend_scene_built_with_cpp_classes_equals_reference(host, golden_dir):
    """random_scene() written against include/rtx reproduces the reference's scene exactly
    (487 spheres, same values after float rounding) and the camera of camera.h:8-45."""
    s, t, q, m, cam, accel = host(0, seed=1, aspect=1200 / 800)
    rows = np.load(os.path.join(scenes.DATA_DIR, "weekend_scene.npy"))
    assert len(s) == 487 and len(m) == 487 and len(t) == 0 and len(q) == 0
    np.testing.assert_array_equal(s["center0"], rows[:, 0:3].astype(np.float32))
    np.testing.assert_array_equal(s["radius"], rows[:, 3].astype(np.float32))
    mm = m[s["material"]]
    np.testing.assert_array_equal(mm["type"], rows[:, 4].astype(np.int32))
    np.testing.assert_array_equal(mm["albedo"], rows[:, 5:8].astype(np.float32))
    np.testing.assert_array_equal(mm["param"], rows[:, 8].astype(np.float32))
    ref = scenes.weekend(1200, 800).camera
    for f in ("origin", "lower_left_corner", "horizontal", "vertical", "u", "v", "w"):
        np.testing.assert_array_equal(np.array(getattr(cam, f)[:]), np.array(getattr(ref, f)[:]))
    assert cam.lens_radius == ref.lens_radius


def test_next_week_scene_raw_pointer_style(host):
    s, t, q, m, cam, accel = host(1, seed=7, aspect=1.5)
    assert accel and len(s) == 488 and s[0]["radius"] == 1000
    assert m[s[0]["material"]]["texture"] == D.RT_TEX_CHECKER
    assert 0.7 < s["moving"].mean() < 0.9  # ~80 % moving lambertian
    mv = s[s["moving"] == 1]
    assert np.all(mv["center1"][:, 1] >= mv["center0"][:, 1]) and np.all(mv["center1"][:, 0] == mv["center0"][:, 0])
    ref = scenes.next_week(1200, 800).camera
    np.testing.assert_allclose(np.array(cam.horizontal[:]), np.array(ref.horizontal[:]), rtol=2e-6)
    assert cam.time1 == 1.0 and abs(cam.lens_radius - 0.025) < 1e-7


def test_obj_room_instances_are_baked_like_python(host, tmp_path):
    """translate(rotate_y(triangle*2.5, 30), (0,1.5,0)) per triangle (obj_render.cu:498-511)
    is baked into vertices; C++ and Python flattening agree; shared materials are shared."""
    obj = str(tmp_path / "blob.obj")
    scenes.make_blob_mesh(obj, subdivisions=1)
    s, t, q, m, cam, accel = host(2, seed=1, aspect=1.0, obj=obj.encode())
    ref = scenes.obj_room(obj, 64, 64)
    assert accel and len(t) == len(ref.triangles) == 80 and len(q) == 10 and len(s) == 2
    np.testing.assert_allclose(t["v0"], ref.triangles["v0"], rtol=0, atol=3e-6)
    np.testing.assert_allclose(t["v2"], ref.triangles["v2"], rtol=0, atol=3e-6)
    np.testing.assert_allclose(t["normal"], ref.triangles["normal"], rtol=0, atol=2e-5)
    np.testing.assert_array_equal(q["axis"], ref.quads["axis"])
    np.testing.assert_allclose(q["k"], ref.quads["k"], rtol=1e-7)
    assert len(set(t["material"])) == 1 and m[t["material"][0]]["type"] == D.RT_MAT_METAL
    assert q["material"][4] == q["material"][5]  # red_1 shared by floor and ceiling
    light = m[m["type"] == D.RT_MAT_DIFFUSE_LIGHT]
    assert len(light) == 3 and np.allclose(sorted(light["albedo"][:, 0]), sorted([5.0, 2 * 255 / 256, 2 * 166 / 256]))


@pytest.mark.gpu
def test_render_cli_writes_the_reference_style_ppm(tmp_path):
    """render_cli = the reference's main(): P3 PPM on stdout, timing on stderr; the image is
    bit-identical to the one rendered through the Python binding with the same seed."""
    from a_dive_into_ray_tracing_b200 import capi
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "apps"), "-s"])
    W, H, spp = 96, 64, 8
    p = subprocess.run([os.path.join(ROOT, "build", "render_cli"), "--scene", "weekend", "--width", str(W),
                        "--height", str(H), "--spp", str(spp), "--seed", "1984"], capture_output=True, text=True,
                       timeout=120)
    assert p.returncode == 0, p.stderr
    assert "took" in p.stderr and "seconds" in p.stderr
    tok = p.stdout.split()
    assert tok[0] == "P3" and int(tok[1]) == W and int(tok[2]) == H and tok[3] == "255"
    img = np.array(tok[4:], np.int32).reshape(H, W, 3)
    with capi.Context(profile=0, seed=1984) as ctx:
        ctx.upload(scenes.weekend(W, H)).build_accel(1)
        ctx.render(W, H, spp)
        _, rgb = ctx.resolve(want_linear=False)
    np.testing.assert_array_equal(img, rgb.astype(np.int32))


def test_obj_reader_variants(host, tmp_path):
    """Beyond the reference's parser: v/vt/vn corners, negative indices, polygons (fan) and
    faces without normals load to the same triangles as the plain a//n triangle form."""
    v = ["v 0 0 0", "v 1 0 0", "v 1 1 0", "v 0 1 0", "v 0.5 0.5 1"]
    vn = ["vn 0 0 -1", "vn 0 0 -1", "vn 0 0 -1", "vn 0 0 -1", "vn 0 0 1"]
    vt = ["vt 0 0", "vt 1 0"]
    fancy = v + vn + vt + ["f 1/1/1 2/2/2 3/1/3 4/2/4",   # quad with texture coordinates
                           "f -5//-5 -4//-4 -1//-1",        # negative indices
                           "f 2 3 5"]                        # no normals
    plain = v + vn + ["f 1//1 2//2 3//3", "f 1//1 3//3 4//4", "f 1//1 2//2 5//5"]
    a = tmp_path / "fancy.obj"
    b = tmp_path / "plain.obj"
    a.write_text("\n".join(fancy) + "\n")
    b.write_text("\n".join(plain) + "\n")
    _, ta, _, _, _, _ = host(2, obj=str(a).encode())
    _, tb, _, _, _, _ = host(2, obj=str(b).encode())
    assert len(ta) == 4 and len(tb) == 3
    for f in ("v0", "v1", "v2", "normal"):
        np.testing.assert_array_equal(ta[f][:3], tb[f])
    # the normal-less face keeps the (reversed) triangle's own orientation
    n = ta["normal"][3].astype(np.float64)
    e1, e2 = ta["v1"][3] - ta["v0"][3], ta["v2"][3] - ta["v0"][3]
    assert np.dot(np.cross(e1, e2), n) > 0


def test_png_writers(tmp_path):
    from PIL import Image
    from a_dive_into_ray_tracing_b200 import capi
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "apps"), "-s", "../build/librtx_host.so"])
    L = C.CDLL(os.path.join(ROOT, "build", "librtx_host.so"))
    rng = np.random.Generator(np.random.Philox(3))
    img = rng.integers(0, 256, (37, 301, 3), dtype=np.uint8)  # > 65535 raw bytes: several stored blocks
    p1, p2 = str(tmp_path / "a.png"), str(tmp_path / "b.png")
    assert L.rtx_host_write_png(p1.encode(), img.ctypes.data_as(C.c_void_p), 301, 37) == 0
    capi.write_png(p2, img)
    for p in (p1, p2):
        np.testing.assert_array_equal(np.asarray(Image.open(p).convert("RGB")), img)
    p3 = str(tmp_path / "c.ppm")
    capi.write_ppm(p3, img, binary=True)
    np.testing.assert_array_equal(np.asarray(Image.open(p3).convert("RGB")), img)


def test_cornell_box_boxes_and_rotations(host):
    """box (box.h:41-58) + rotate_y + translate (main.cu:269-275): C++ and Python flatten the
    Cornell box to the same rects / triangles; rotated box sides keep their nominal normals."""
    s, t, q, m, cam, accel = host(3, aspect=1.0)
    ref = scenes.cornell_box(600, 600)
    assert accel and len(s) == 0 and len(q) == 6 and len(t) == 24 and len(m) == 4
    for f in ("axis", "a0", "a1", "b0", "b1", "k", "material"):
        np.testing.assert_array_equal(q[f], ref.quads[f])
    np.testing.assert_array_equal(m["type"], ref.materials["type"])
    np.testing.assert_allclose(m["albedo"], ref.materials["albedo"], rtol=1e-7)
    for f in ("v0", "v1", "v2"):
        np.testing.assert_allclose(t[f], ref.triangles[f], rtol=0, atol=1e-4)
    np.testing.assert_allclose(t["normal"], ref.triangles["normal"], rtol=2e-6, atol=0.5)
    np.testing.assert_array_equal(t["material"], ref.triangles["material"])
    # first side of box1 = xy_rect at z = 165 rotated by 15 degrees: normal (sin15, 0, cos15)
    n = t["normal"][0] / np.linalg.norm(t["normal"][0])
    np.testing.assert_allclose(n, [np.sin(np.deg2rad(15)), 0, np.cos(np.deg2rad(15))], atol=1e-5)
    np.testing.assert_allclose(np.array(cam.origin[:]), [278, 278, -800])


def _host_extras():
    L = C.CDLL(os.path.join(ROOT, "build", "librtx_host.so"))
    n = np.zeros(3, np.int32)
    dims = np.zeros(16, np.int32)
    L.rtx_host_counts2.argtypes = [_vp, _vp]
    L.rtx_host_get2.argtypes = [_vp, _vp, _vp]
    L.rtx_host_counts2(n.ctypes.data, dims.ctypes.data)
    media = np.zeros(n[0], D.MEDIUM_DT)
    perlin = np.zeros(n[1], D.PERLIN_DT)
    images = [np.zeros((dims[2 * i + 1], dims[2 * i], 3), np.uint8) for i in range(n[2])]
    ptrs = (C.c_void_p * max(len(images), 1))(*[im.ctypes.data for im in images])
    L.rtx_host_get2(media.ctypes.data, perlin.ctypes.data, C.addressof(ptrs))
    return media, perlin, images


def test_next_week_scenes_3_to_8_cpp_equals_python(host):
    """noise_texture / image_texture / isotropic / constant_medium of include/rtx flatten to the
    same description as the Python scenes (main.cu:212-383)."""
    # 4: two perlin spheres — two lambertians sharing one noise texture -> one table
    s, t, q, m, cam, accel = host(4, seed=3)
    media, perlin, images = _host_extras()
    ref = scenes.two_perlin_spheres(600, 400)
    assert len(s) == 2 and len(m) == 2 and len(perlin) == 1 and len(media) == 0 and not images
    assert np.all(m["texture"] == D.RT_TEX_NOISE) and np.all(m["albedo2"][:, 0] == 4.0) and np.all(m["albedo2"][:, 1] == 0.0)
    np.testing.assert_array_equal(s["center0"], ref.spheres["center0"])
    for name in ("perm_x", "perm_y", "perm_z"):
        assert sorted(perlin[name][0].tolist()) == list(range(256))
    assert np.abs(perlin["ranvec"]).max() <= 1.0 and perlin["ranvec"].std() > 0.4
    # 5: earth — the image bytes travel intact
    s, t, q, m, cam, accel = host(5)
    media, perlin, images = _host_extras()
    assert len(images) == 1 and images[0].shape == (32, 64, 3) and m["texture"][0] == D.RT_TEX_IMAGE
    assert images[0].std() > 10  # not blank
    # 6: simple_light
    s, t, q, m, cam, accel = host(6)
    ref = scenes.simple_light(600, 400)
    assert len(s) == 3 and len(q) == 1
    np.testing.assert_array_equal(m[s["material"]]["type"], ref.materials[ref.spheres["material"]]["type"])
    np.testing.assert_array_equal(m[q["material"]]["albedo"], [[4, 4, 4]])
    np.testing.assert_allclose(np.array(cam.origin[:]), [26, 3, 6])
    # 7: cornell_smoke — rotated + translated boxes become oriented media
    s, t, q, m, cam, accel = host(7, aspect=1.0)
    media, perlin, images = _host_extras()
    ref = scenes.cornell_smoke(600, 600)
    assert len(q) == 6 and len(t) == 0 and len(media) == 2
    for f in ("shape", "p0", "p1", "offset", "density"):
        np.testing.assert_array_equal(media[f], ref.media[f])
    np.testing.assert_allclose(media["sin_y"], ref.media["sin_y"], atol=1e-7)
    np.testing.assert_allclose(media["cos_y"], ref.media["cos_y"], atol=1e-7)
    np.testing.assert_array_equal(m[media["material"]]["type"], [D.RT_MAT_ISOTROPIC] * 2)
    np.testing.assert_array_equal(m[media["material"]]["albedo"], [[0, 0, 0], [1, 1, 1]])
    # 8: the final scene — counts and the fixed objects
    s, t, q, m, cam, accel = host(8, seed=5, aspect=1.0)
    media, perlin, images = _host_extras()
    ref = scenes.next_week_final(800, 800)
    assert len(q) == 400 * 6 + 1 == len(ref.quads) and len(s) == 6 + 1000 == len(ref.spheres)
    assert len(media) == 2 and len(perlin) == 1 and len(images) == 1
    np.testing.assert_array_equal(s["center0"][:6], ref.spheres["center0"][:6])
    np.testing.assert_array_equal(s["radius"][:6], ref.spheres["radius"][:6])
    assert s["moving"][0] == 1 and tuple(s["center1"][0]) == (430, 400, 200)
    np.testing.assert_array_equal(media["p0"], ref.media["p0"])
    np.testing.assert_array_equal(media["p1"], ref.media["p1"])
    np.testing.assert_allclose(media["density"], [0.2, 0.0001])
    # the cluster: rotated by 15 degrees and moved to (-100, 270, 395): inside the transformed cube
    c = s["center0"][6:].astype(np.float64) - [-100, 270, 395]
    ang = np.deg2rad(15)
    local = np.stack([np.cos(ang) * c[:, 0] - np.sin(ang) * c[:, 2], c[:, 1], np.sin(ang) * c[:, 0] + np.cos(ang) * c[:, 2]], 1)
    assert local.min() > -1e-3 and local.max() < 165 + 1e-3
    # boxes: heights in [1, 101), footprint 100 x 100
    tops = q["k"][(q["axis"] == 1)][:-1].reshape(400, 2)[:, 0]
    assert tops.min() >= 1 and tops.max() < 101


def test_constant_medium_rejects_unsupported_boundaries(tmp_path):
    src = tmp_path / "bad.cpp"
    src.write_text('#include "constant_medium.h"\n#include "aarect.h"\nint main() {\n'
                   '  auto m = new lambertian(color(1, 1, 1));\n'
                   '  hittable *r = new xy_rect(0, 1, 0, 1, 0, m);\n'
                   '  hittable *cm = new constant_medium(r, 0.1, color(1, 1, 1));\n'
                   '  flat_scene fs;\n'
                   '  try { cm->flatten(fs, transform()); } catch (const std::invalid_argument &) { return 0; }\n'
                   '  return 1;\n}\n')
    exe = tmp_path / "bad"
    subprocess.check_call(["g++", "-std=c++17", "-I", os.path.join(ROOT, "include", "rtx"), "-I",
                           os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    assert subprocess.run([str(exe)]).returncode == 0


@pytest.mark.gpu
def test_render_cli_final_scene(tmp_path):
    """the rt_next_week tree's default scene through the C++ host program"""
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "apps"), "-s"])
    out = tmp_path / "final.ppm"
    with open(out, "wb") as f:
        p = subprocess.run([os.path.join(ROOT, "build", "render_cli"), "--scene", "final", "--width", "120", "--height",
                            "120", "--spp", "64", "--binary"], stdout=f, stderr=subprocess.PIPE, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    assert "2 media" in p.stderr and "1006 spheres" in p.stderr
    raw = out.read_bytes()
    assert raw.startswith(b"P6\n120 120\n255\n")
    img = np.frombuffer(raw[len(b"P6\n120 120\n255\n"):], np.uint8).reshape(120, 120, 3)
    assert 20 < img.mean() < 200 and img.std() > 10


@pytest.mark.gpu
def test_render_cli_multi_gpu_sample_split(tmp_path):
    """render_cli --gpus 2: one context and host thread per device, frames combined by rt_reduce (ncclReduce) or,
    with --host-combine, summed on the host; both equal the single-device image up to fp32 summation order
    (counter-based RNG)."""
    from a_dive_into_ray_tracing_b200 import capi
    if capi.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "apps"), "-s"])
    imgs = []
    for extra in ([], ["--gpus", "2"], ["--gpus", "2", "--host-combine"]):
        p = subprocess.run([os.path.join(ROOT, "build", "render_cli"), "--scene", "weekend", "--width", "120", "--height",
                            "80", "--spp", "32", "--binary"] + extra, capture_output=True, timeout=300)
        assert p.returncode == 0, p.stderr.decode()
        head = b"P6\n120 80\n255\n"
        imgs.append(np.frombuffer(p.stdout[len(head):], np.uint8).reshape(80, 120, 3).astype(np.int32))
    for other in imgs[1:]:
        d = np.abs(imgs[0] - other)
        assert d.max() <= 1 and (d > 0).mean() < 0.01, (d.max(), (d > 0).mean())
