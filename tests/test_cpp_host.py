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
    rows = np.load(golden_dir + "/weekend_scene.npy")
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
