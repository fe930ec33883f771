"""Text scene files (include/rtx/scene_file.h; SURVEY.md §8f rank 3 "scene file instead of
compiled-in scenes"): the C++ parser builds the reference's own classes, which flatten to the same
description as the hand-written scenes; errors carry file:line; the device code renders a scene
using every statement the same as the oracle."""
import os
import subprocess

import numpy as np
import pytest

from a_dive_into_ray_tracing_b200 import ctypes_defs as D
from a_dive_into_ray_tracing_b200 import scenes
from tests import stats_util as SU
from tests.emu.pyemu import Emu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCENES = os.path.join(ROOT, "scenes")


def test_cornell_smoke_file_equals_compiled_scene():
    sc, (w, h, spp) = scenes.load_scene_file(os.path.join(SCENES, "cornell_smoke.scene"))
    ref = scenes.cornell_smoke(600, 600)
    assert (w, h, spp) == (600, 600, 200) and sc.profile == 2 and sc.sky_gradient == 0
    for f in ("axis", "a0", "a1", "b0", "b1", "k"):
        np.testing.assert_array_equal(sc.quads[f], ref.quads[f])
    np.testing.assert_array_equal(sc.materials[sc.quads["material"]]["albedo"], ref.materials[ref.quads["material"]]["albedo"])
    for f in ("shape", "p0", "p1", "offset", "density"):
        np.testing.assert_array_equal(sc.media[f], ref.media[f])
    np.testing.assert_allclose(sc.media["sin_y"], ref.media["sin_y"], atol=1e-7)
    np.testing.assert_array_equal(sc.materials[sc.media["material"]]["albedo"], [[0, 0, 0], [1, 1, 1]])
    for f in ("origin", "lower_left_corner", "horizontal", "vertical"):
        np.testing.assert_allclose(np.array(getattr(sc.camera, f)[:]), np.array(getattr(ref.camera, f)[:]), rtol=1e-6)


def test_weekend_profile_file_and_camera_override():
    sc, dims = scenes.load_scene_file(os.path.join(SCENES, "three_spheres.scene"))
    assert sc.profile == 0 and sc.sky_gradient == 1 and dims == (400, 225, 100)
    np.testing.assert_array_equal(sc.spheres["radius"], np.array([100, 0.5, 0.5, -0.45, 0.5], np.float32))
    assert sc.spheres["material"][2] == sc.spheres["material"][3]  # the two glass spheres share one material
    assert len(sc.materials) == 4
    ref = D.camera_from_lookat((-2, 2, 1), (0, 0, -1), (0, 1, 0), 20.0, 400 / 225, 0.1, 3.4, dtype=np.float64)
    for f in ("origin", "lower_left_corner", "horizontal", "vertical", "u", "v"):
        np.testing.assert_array_equal(np.array(getattr(sc.camera, f)[:]), np.array(getattr(ref, f)[:]))
    # a caller-chosen frame changes the aspect ratio of the camera
    sc2, dims2 = scenes.load_scene_file(os.path.join(SCENES, "three_spheres.scene"), width=300, height=300)
    assert dims2[:2] == (300, 300)
    assert abs(np.linalg.norm(sc2.camera.horizontal[:]) / np.linalg.norm(sc2.camera.vertical[:]) - 1.0) < 1e-6


def test_obj_and_image_statements(tmp_path):
    obj = tmp_path / "mesh.obj"
    scenes.make_blob_mesh(str(obj), subdivisions=1)
    img = scenes.procedural_earth(16, 8)
    with open(tmp_path / "map.ppm", "wb") as f:
        f.write(b"P6\n# a comment\n16 8\n255\n" + img.tobytes())
    (tmp_path / "s.scene").write_text(
        "profile next_week\nimage 64 64 4\ncamera lookfrom 0 0 9 lookat 0 0 0 vfov 40\nbackground 0.5 0.5 0.5\n"
        "flags flip_normals depth_background\ntmin 0.00001\n"
        "material gold metal 1 0.84 0 0.5\nmaterial map lambertian image map.ppm\n"
        "obj mesh.obj gold scale 2.5 rotate_y 30 translate 0 1.5 0\nsphere 3 0 0 1 map\n")
    sc, _ = scenes.load_scene_file(str(tmp_path / "s.scene"))
    ntri = len(scenes.read_obj_triangles(str(obj))[0])
    assert len(sc.triangles) == ntri and len(sc.images) == 1
    np.testing.assert_array_equal(sc.images[0], img)
    assert sc.flags == D.RT_FLAG_FLIP_NORMALS | D.RT_FLAG_DEPTH_BACKGROUND and abs(sc.t_min - 1e-5) < 1e-12
    # the same placement as the compiled obj_room scene: scale 2.5, rotate 30 degrees, lift 1.5
    ref = scenes.obj_room(str(obj), 64, 64)
    np.testing.assert_allclose(sc.triangles["v0"], ref.triangles["v0"][:ntri], atol=2e-6)
    np.testing.assert_allclose(sc.triangles["normal"], ref.triangles["normal"][:ntri], atol=2e-5)


@pytest.mark.parametrize("text,needle", [
    ("camera lookfrom 0 0 1 lookat 0 0 0 vfov 40\nsphere 0 0 0 1 nope\n", ":2: unknown material 'nope'"),
    ("camera lookfrom 0 0 1 lookat 0 0 0 vfov 40\nsphere 0 0 zero 1 m\n", ":2: expected a number"),
    ("camera lookfrom 0 0 1 lookat 0 0 0 vfov 40\nteapot 1 2 3\n", ":2: unknown statement 'teapot'"),
    ("material m lambertian 1 1 1\nsphere 0 0 0 1 m\n", "no camera statement"),
    ("camera lookfrom 0 0 1 vfov 40\n", ":1: camera needs lookfrom, lookat and vfov"),
    ("camera lookfrom 0 0 1 lookat 0 0 0 vfov 40\nmaterial m metal 1 1 1\n", ":2: missing fuzz"),
    ("camera lookfrom 0 0 1 lookat 0 0 0 vfov 40\nmaterial m lambertian 1 1 1\nsphere 0 0 0 1 m extra\n", ":3: unexpected 'extra'"),
    ("camera lookfrom 0 0 1 lookat 0 0 0 vfov 40\nmedium sphere 0 0 0 1 0 1 1 1\n", ":2: medium density must be positive"),
    ("camera lookfrom 0 0 1 lookat 0 0 0 vfov 40\nmaterial i lambertian image missing.ppm\n", ":2: cannot read image"),
])
def test_errors_name_file_and_line(tmp_path, text, needle):
    p = tmp_path / "bad.scene"
    p.write_text(text)
    with pytest.raises(ValueError) as e:
        scenes.load_scene_file(str(p))
    assert needle in str(e.value) and "bad.scene" in str(e.value)


def test_showcase_scene_device_code_vs_oracle(l1_32):
    """every statement kind in one scene: emulated device code vs the restatement"""
    W, H, spp = 48, 30, 128
    sc, _ = scenes.load_scene_file(os.path.join(SCENES, "showcase.scene"), W, H)
    assert len(sc.media) == 2 and len(sc.perlin) == 1 and sc.spheres["moving"].sum() == 1
    s, s2, st = Emu(sc, quality=1).render(W, H, spp, seed=7)
    r, r2, nseg = l1_32.render_parallel(sc, 2, W, H, spp, seed=5)
    mu_a, var_a = SU.mean_var(s, s2, spp)
    mu_b, var_b = SU.mean_var(r, r2, spp)
    ok, d, b = SU.three_sigma_check(mu_a, var_a, spp, mu_b, var_b, spp)
    assert ok, (d, b)
    z = SU.zscores(mu_a, var_a, spp, mu_b, var_b, spp)
    assert abs(z.mean()) < 0.15 and 0.8 < z.std() < 1.25, (z.mean(), z.std())


@pytest.mark.gpu
def test_render_cli_scene_file_equals_python_binding(tmp_path):
    """render_cli --scene-file == loading the same file through the Python binding (same seed)"""
    from a_dive_into_ray_tracing_b200 import capi
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "apps"), "-s"])
    path = os.path.join(SCENES, "showcase.scene")
    W, H, spp = 96, 60, 16
    out = tmp_path / "o.ppm"
    with open(out, "wb") as f:
        p = subprocess.run([os.path.join(ROOT, "build", "render_cli"), "--scene-file", path, "--width", str(W), "--height",
                            str(H), "--spp", str(spp), "--binary"], stdout=f, stderr=subprocess.PIPE, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    raw = out.read_bytes()
    head = b"P6\n%d %d\n255\n" % (W, H)
    img = np.frombuffer(raw[len(head):], np.uint8).reshape(H, W, 3)
    sc, _ = scenes.load_scene_file(path, W, H)
    with capi.Context(profile=sc.profile, seed=7) as ctx:  # the file says `seed 7`
        ctx.upload(sc).build_accel(1)
        ctx.render(W, H, spp)
        _, rgb = ctx.resolve(want_linear=False)
    np.testing.assert_array_equal(img, rgb)
    assert img.std() > 10


def test_parser_never_crashes_on_garbage(tmp_path):
    """fuzz: random statements built from the format's vocabulary either load or raise a ValueError
    naming the file - the C++ parser must not crash or accept trailing junk silently"""
    from hypothesis import HealthCheck, given, settings
    from hypothesis import strategies as st
    words = ["profile", "next_week", "weekend", "image", "camera", "lookfrom", "lookat", "vfov", "vup", "aperture",
             "focus", "shutter", "sky", "background", "tmin", "depth", "seed", "flags", "flip_normals", "material",
             "m", "lambertian", "checker", "noise", "image", "metal", "dielectric", "light", "sphere",
             "moving_sphere", "xy_rect", "xz_rect", "yz_rect", "box", "triangle", "obj", "medium", "rotate_y",
             "translate", "scale", "0", "1", "-2.5", "1e3", "nan", "inf", "-", "#", "x.ppm", "y.obj", "0.5"]
    line = st.lists(st.sampled_from(words), min_size=0, max_size=14).map(" ".join)
    p = tmp_path / "fuzz.scene"

    @settings(max_examples=150, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
    @given(st.lists(line, min_size=0, max_size=8))
    def run(lines):
        p.write_text("camera lookfrom 0 0 5 lookat 0 0 0 vfov 40\nmaterial m lambertian 0.5 0.5 0.5\n" + "\n".join(lines) + "\n")
        try:
            sc, dims = scenes.load_scene_file(str(p))
        except ValueError as e:
            assert "fuzz.scene" in str(e)
            return
        assert dims[0] >= 2 and dims[1] >= 2 and dims[2] >= 1
        assert np.all(sc.spheres["material"] < len(sc.materials))

    run()


def test_instancing_and_reference_medium_flags(tmp_path):
    """`flags instancing`: objects under an XFORM arrive as rt_group + rt_instance (object-space geometry), and the
    emulated two-level walk sees the same surfaces as the baked form of the same file; `flags reference_medium`
    sets RT_FLAG_REFERENCE_MEDIUM"""
    body = """profile next_week
image 64 64 8
camera lookfrom 278 278 -800 lookat 278 278 0 vfov 40
background 0 0 0
material white lambertian 0.73 0.73 0.73
material lamp light 15 15 15
xz_rect 213 343 227 332 554 lamp
xz_rect 0 555 0 555 0 white
box 0 0 0 165 330 165 white rotate_y 15 translate 265 0 295
box 0 0 0 165 165 165 white rotate_y -18 translate 130 0 65
sphere 400 100 100 60 white translate 0 20 0
"""
    a = tmp_path / "baked.scene"
    b = tmp_path / "inst.scene"
    a.write_text(body)
    b.write_text(body + "flags instancing reference_medium\n")
    sa, _ = scenes.load_scene_file(str(a))
    sb, _ = scenes.load_scene_file(str(b))
    assert len(sa.groups) == 0 and len(sa.instances) == 0 and len(sa.triangles) == 24
    assert len(sb.groups) == 3 and len(sb.instances) == 3 and len(sb.triangles) == 0
    assert sb.flags & D.RT_FLAG_REFERENCE_MEDIUM and not (sa.flags & D.RT_FLAG_REFERENCE_MEDIUM)
    np.testing.assert_array_equal(sb.groups["n_quads"], [6, 6, 0])
    np.testing.assert_array_equal(sb.groups["n_spheres"], [0, 0, 1])
    np.testing.assert_allclose(sb.spheres["center0"][0], [400, 100, 100])         # object space
    np.testing.assert_allclose(sb.instances["m"][2][[3, 7, 11]], [0, 20, 0])        # the translation
    rays = D.primary_rays(sa.camera, 64, 64, 2)
    ia, ta, _ = Emu(sa).trace(rays, t_min=sa.t_min)
    ib, nb, tb = Emu(sb, max_leaf=1).trace_inst(rays, t_min=sb.t_min)
    assert ((ia >= 0) != (ib >= 0)).mean() < 2e-3
    k = (ia >= 0) & (ib >= 0)
    assert np.quantile(np.abs(ta[k] - tb[k]) / ta[k], 0.999) < 1e-5
    assert set(np.unique(nb[k]).tolist()) == {-1, 0, 1, 2}
