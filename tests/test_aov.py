"""Denoiser feature buffers (rt_render_aov). The reference has no AOVs, so
parity is structural: the buffers must agree with the closest-hit query and the scene description
(CPU, device code emulated), and the kernel must agree with its own device code run on the CPU."""
import numpy as np
import pytest

from a_dive_into_ray_tracing_b200 import ctypes_defs as D
from a_dive_into_ray_tracing_b200 import scenes
from tests.emu.pyemu import Emu


def pinhole(sc, W, H):
    """the scene with a zero-aperture camera, so that a pixel's samples stay inside the pixel"""
    cam = D.camera_from_lookat((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, W / H, 0.0, 10.0,
                               dtype=np.float64 if sc.profile == 0 else np.float32)
    return sc.with_camera(cam)


@pytest.mark.parametrize("name", ["weekend", "next_week"])
def test_aov_agrees_with_closest_hit_and_materials(name):
    W, H = 96, 64
    sc = pinhole(getattr(scenes, name)(W, H), W, H)
    e = Emu(sc, quality=1)
    a = e.aov(W, H, spp=4)
    assert a.shape == (H, W, 8) and np.all(np.isfinite(a))
    hitf = a[..., 7]
    assert np.all((hitf >= 0) & (hitf <= 1))
    # pixel-centre rays through the closest-hit query
    rays = D.primary_rays(sc.camera, W, H, sc.profile, time=0.0)  # the pinhole camera's shutter is closed: tm = 0
    ids, ts, _ = e.trace(rays, t_min=sc.t_min)
    ids, ts = ids.reshape(H, W), ts.reshape(H, W)
    # interior pixels: the 3x3 neighbourhood sees one primitive (or only sky)
    same = np.ones((H, W), bool)
    for dj in (-1, 0, 1):
        for di in (-1, 0, 1):
            same &= np.roll(np.roll(ids, dj, 0), di, 1) == ids
    same[[0, -1], :] = False
    same[:, [0, -1]] = False
    static = np.ones((H, W), bool)
    if name == "next_week":  # moving spheres: their depth / normal depend on the sample's time
        moving = np.zeros(len(sc.spheres) + 1, bool)
        moving[:-1] = sc.spheres["moving"] != 0
        static = ~moving[np.where(ids >= 0, D.prim_index_of(ids), -1)]
    inner_hit = same & (ids >= 0) & static
    inner_sky = same & (ids < 0)
    assert inner_hit.sum() > 500 and inner_sky.sum() > 50
    assert np.all(hitf[inner_hit] == 1.0) and np.all(hitf[inner_sky] == 0.0)
    # depth: the mean over the pixel's samples is close to the pixel-centre value
    rel = np.abs(a[..., 6][inner_hit] - ts[inner_hit]) / ts[inner_hit]
    assert np.percentile(rel, 99) < 0.02, np.percentile(rel, 99)
    # normals: unit length on interior pixels (the mean of nearly equal unit vectors), facing the camera ray
    # when the profile flips them (profile 0)
    nrm = a[..., 3:6]
    ln = np.linalg.norm(nrm[inner_hit], axis=1)
    assert np.all(ln < 1.0 + 1e-5) and np.all(ln > 0.9) and np.median(ln) > 0.995, (ln.min(), ln.max())
    assert np.all(a[..., 3:6][inner_sky] == 0.0) and np.all(a[..., 6][inner_sky] == 0.0)
    if sc.profile == 0:
        d = rays[:, 4:7].reshape(H, W, 3)
        assert np.all(np.einsum("...k,...k", nrm, d)[inner_hit] < 0)
    # albedo: solid materials give exactly their colour, glass gives 1
    mat = sc.materials[sc.spheres["material"][D.prim_index_of(ids[inner_hit])]]
    want = np.where((mat["type"] == D.RT_MAT_DIELECTRIC)[:, None], 1.0, mat["albedo"])
    solid = mat["texture"] == D.RT_TEX_SOLID
    np.testing.assert_allclose(a[..., 0:3][inner_hit][solid], want[solid], atol=1e-6)
    # sky pixels carry the miss colour
    if name == "next_week":
        np.testing.assert_allclose(a[..., 0:3][inner_sky], np.tile(sc.background, (inner_sky.sum(), 1)), atol=1e-6)
    else:
        assert np.all(a[..., 2][inner_sky] > 0.99) and np.all(a[..., 0][inner_sky] >= 0.5)


def test_aov_textures_lights_and_flipped_normals():
    W, H = 64, 64
    sc = scenes.cornell_box(W, H)
    a = Emu(sc).aov(W, H, spp=2)
    # the light (emission 15) is clamped to 1; walls carry their lambertian colours
    assert abs(a[..., 0:3].max() - 1.0) < 1e-6
    cols = {tuple(round(float(x), 2) for x in c) for c in a[..., 0:3].reshape(-1, 3)[a[..., 7].reshape(-1) == 1.0]}
    assert {(0.65, 0.05, 0.05), (0.12, 0.45, 0.15), (0.73, 0.73, 0.73), (1.0, 1.0, 1.0)} <= cols
    # rt_next_week rects keep their nominal +axis normal; with RT_FLAG_FLIP_NORMALS they face the ray
    sc.flags |= D.RT_FLAG_FLIP_NORMALS
    b = Emu(sc).aov(W, H, spp=2)
    rays = D.primary_rays(sc.camera, W, H, sc.profile).reshape(H, W, 8)
    full = b[..., 7] == 1.0
    assert np.all(np.einsum("...k,...k", b[..., 3:6], rays[..., 4:7])[full] < 1e-6)
    assert np.any(np.einsum("...k,...k", a[..., 3:6], rays[..., 4:7])[full] > 0.1)
    # noise texture: albedo in [0, 1], varying
    scn = scenes.two_perlin_spheres(W, H)
    c = Emu(scn).aov(W, H, spp=1)
    hit = c[..., 7] == 1.0
    assert 0.0 <= c[..., 0][hit].min() < 0.3 and 0.7 < c[..., 0][hit].max() <= 1.0


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["weekend", "final_cu", "next_week", "cornell_box", "simple_light"])
def test_gpu_aov_equals_emulated_device_code(name):
    from a_dive_into_ray_tracing_b200 import capi
    W, H, spp = 80, 60, 3
    sc = getattr(scenes, name)(W, H)
    ref = Emu(sc, quality=1).aov(W, H, spp, seed=99)
    with capi.Context(profile=sc.profile, seed=99) as ctx:
        ctx.upload(sc).build_accel(1)
        got = ctx.render_aov(W, H, spp)
        with pytest.raises(capi.RtError):
            ctx.render_aov(W, H, 0)
    assert got.shape == ref.shape and np.all(np.isfinite(got))
    # same camera streams on both sides; MUFU arithmetic may flip a sample on a silhouette
    close = np.all(np.abs(got - ref) <= 1e-3 + 1e-4 * np.abs(ref), axis=-1)
    assert close.mean() > 0.99, (name, close.mean())
    assert abs(got[..., 7].mean() - ref[..., 7].mean()) < 2e-3
