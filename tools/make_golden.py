"""tools/make_golden.py — generate the committed golden fixtures in tests/golden/ by
running the UNMODIFIED reference (oracle/_ref/libref_l0.so, built from
/root/reference/rt_in_one_weekend) in this container. Re-run only where the
reference sources are present.

  weekend_scene.npy        [487][12] float64 rows, values rounded to float32: the
                           reference's random_scene() under glibc's default seed
  weekend_scene_f64.npy    the same rows un-rounded (as the reference holds them)
  weekend_hits_c1.npz      config-1 camera (400x225): float32 primary rays on a 100x56
                           sub-grid + jittered rays; reference closest-hit (id, t)
                           on the float-rounded scene; robustness measures
  weekend_render_c1.npz    reference render 100x56 @ 64 spp of the config-1 view
                           (per-pixel sum, sum of squares, segment count)
  weekend_refvsref.json    reference-vs-reference PSNR floor (two seeds)
  gallery_final_75x50.npy  the reference's gallery/final.png (1200x800, its 500-spp
                           config-2 image) box-downsampled 16x16, uint8->float mean
  gallery_final_300x200.npy  the same image box-downsampled 4x4 (uint16 = 4 x mean level)
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from a_dive_into_ray_tracing_b200 import ctypes_defs as D  # noqa: E402
from a_dive_into_ray_tracing_b200 import scenes  # noqa: E402
from oracle import pyoracle  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")


def fnv1a64(b):
    h = 0xCBF29CE484222325
    for x in b:
        h = ((h ^ x) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


def psnr(a, b, peak=1.0):
    mse = np.mean((a - b) ** 2)
    return 10 * np.log10(peak * peak / mse)


def main():
    pyoracle.build()
    l0 = pyoracle.L0()
    rows64 = l0.scene_rows()
    assert rows64.shape == (487, 12)
    np.save(os.path.join(G, "weekend_scene_f64.npy"), rows64)
    rows = rows64.astype(np.float32).astype(np.float64)
    np.save(os.path.join(ROOT, "scenes", "weekend_scene.npy"), rows)
    meta = {"n": 487, "kinds": np.bincount(rows[:, 4].astype(int)).tolist(),
            "fnv1a64_f64_rows": "%016x" % fnv1a64(rows64[:, :9].tobytes()),
            "fnv1a64_f32_rows": "%016x" % fnv1a64(rows[:, :9].astype(np.float32).tobytes())}

    # ---- closest-hit golden vectors on the float-rounded scene
    l0.scene_set(rows)
    W, H = 400, 225
    sc = scenes.weekend(W, H)
    cam = sc.camera
    sub = [(j * W + i) for j in range(2, H, 4) for i in range(2, W, 4)]
    r_center = D.primary_rays(cam, W, H, 0, pixels=sub)
    rng = np.random.Generator(np.random.Philox(7))
    px = rng.integers(0, W * H, 6000)
    parts = [r_center]
    for k in range(3):
        sel = px[2000 * k:2000 * (k + 1)]
        a = rng.random() * 2 * np.pi
        rr = np.sqrt(rng.random())
        parts.append(D.primary_rays(cam, W, H, 0, s_jitter=rng.random(), t_jitter=rng.random(),
                                    lens=(rr * np.cos(a), rr * np.sin(a)), pixels=sel))
    rays = np.concatenate(parts).astype(np.float32)
    rays6 = np.concatenate([rays[:, 0:3], rays[:, 4:7]], 1).astype(np.float64)
    ids, ts = l0.closest_hit(rays6, 1e-3, np.inf)
    gap, drel = l0.hit_robustness(rays6, 1e-3, np.inf)
    np.savez_compressed(os.path.join(G, "weekend_hits_c1.npz"), rays=rays, ids=ids, t=ts, gap=gap, disc_rel=drel)
    meta["hits"] = {"n_rays": int(len(rays)), "n_hit": int((ids >= 0).sum()), "n_ground": int((ids == 0).sum())}

    # ---- secondary (bounce) rays: origins on hit points, random directions
    hit = ids >= 0
    p = rays6[hit, 0:3] + ts[hit, None] * rays6[hit, 3:6]
    d = rng.normal(size=p.shape)
    d /= np.linalg.norm(d, axis=1)[:, None]
    d *= rng.uniform(0.2, 2.0, size=(len(d), 1))
    rays_b = np.zeros((len(p), 8), np.float32)
    rays_b[:, 0:3] = p
    rays_b[:, 4:7] = d
    rb6 = np.concatenate([rays_b[:, 0:3], rays_b[:, 4:7]], 1).astype(np.float64)
    idb, tb = l0.closest_hit(rb6, 1e-3, np.inf)
    gapb, drelb = l0.hit_robustness(rb6, 1e-3, np.inf)
    np.savez_compressed(os.path.join(G, "weekend_hits_bounce.npz"), rays=rays_b, ids=idb, t=tb, gap=gapb,
                        disc_rel=drelb)
    meta["hits_bounce"] = {"n_rays": int(len(rays_b)), "n_hit": int((idb >= 0).sum())}

    # ---- reference render, config-1 view, 100x56 @ 64 spp, float-rounded scene and camera
    Wr, Hr, spp = 100, 56, 64
    cam13 = pyoracle.WEEKEND_CAM13(Wr / Hr)
    s, s2, nseg = l0.render(Wr, Hr, spp, cam13, seed=1)
    np.savez_compressed(os.path.join(G, "weekend_render_c1.npz"), sum=s, sumsq=s2, spp=spp, segments=nseg,
                        W=Wr, H=Hr)
    meta["render_c1"] = {"W": Wr, "H": Hr, "spp": spp, "segments": nseg, "seg_per_path": nseg / (Wr * Hr * spp),
                         "mean_rgb": (s.mean((0, 1)) / spp).tolist()}
    sb, _, _ = l0.render(Wr, Hr, spp, cam13, seed=2, want_sumsq=False)
    g = lambda x: np.sqrt(np.clip(x / spp, 0, 1))
    meta["refvsref"] = {"spp": spp, "psnr_gamma_db": float(psnr(g(s), g(sb))),
                        "psnr_linear_db": float(psnr(np.clip(s / spp, 0, 1), np.clip(sb / spp, 0, 1))),
                        "mean_abs_delta": np.abs(s / spp - sb / spp).mean((0, 1)).tolist()}

    # ---- the reference's gallery image (its converged config-2 output)
    from PIL import Image
    im = np.asarray(Image.open("/root/reference/gallery/final.png").convert("RGB"), np.float64)
    assert im.shape == (800, 1200, 3)
    ds = im.reshape(50, 16, 75, 16, 3).mean((1, 3)) / 255.0
    np.save(os.path.join(G, "gallery_final_75x50.npy"), ds.astype(np.float32))
    # 4x4 box-downsampled copy (sum of 16 8-bit values / 4, exact in uint16) for the matched-spp PSNR test
    ds4 = im.reshape(200, 4, 300, 4, 3).mean((1, 3))
    np.save(os.path.join(G, "gallery_final_300x200.npy"), np.round(ds4 * 4).astype(np.uint16))
    meta["gallery"] = {"shape": [50, 75, 3], "mean_rgb": ds.mean((0, 1)).tolist(),
                       "note": "top row first, gamma-space [0,1]"}
    with open(os.path.join(G, "weekend_meta.json"), "w") as fh:
        json.dump(meta, fh, indent=1)
    print(json.dumps(meta, indent=1))


if __name__ == "__main__":
    main()
