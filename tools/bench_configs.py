"""All five BASELINE.json configs through the C ABI on one GPU (full sizes): frame time,
path-bounces/s, BVH build time. Config 5 is the 4K x 5000 spp frame (sample-split over
GPUs in bench.py's multi-GPU mode; here its single-GPU time). Usage:
  python tools/bench_configs.py [--quick]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from a_dive_into_ray_tracing_b200 import capi, scenes  # noqa: E402

quick = "--quick" in sys.argv
CONFIGS = [
    ("C1 weekend 400x225 10spp", lambda: scenes.weekend(400, 225), 400, 225, 10),
    ("C2 weekend 1200x800 500spp", lambda: scenes.weekend(1200, 800), 1200, 800, 500),
    ("C3 obj_room(blob mesh 320 tris) 800x800 1500spp", lambda: scenes.obj_room(width=800, height=800), 800, 800, 1500),
    ("C3b obj_room(blob mesh 1280 tris) 800x800 1500spp",
     lambda: scenes.obj_room(width=800, height=800, subdivisions=3), 800, 800, 1500),
    ("C4 next_week 1200x800 1000spp", lambda: scenes.next_week(1200, 800), 1200, 800, 1000),
    ("C5 weekend 3840x2160 5000spp", lambda: scenes.weekend(3840, 2160), 3840, 2160, 5000),
    # rt_next_week/cuda/main.cu:453-459,507-510: the tree's default scene at its default size
    ("N8 next_week_final 800x800 5000spp", lambda: scenes.next_week_final(800, 800), 800, 800, 5000),
    ("N7 cornell_smoke 600x600 1000spp", lambda: scenes.cornell_smoke(600, 600), 600, 600, 1000),
    ("N6 cornell_box 600x600 1000spp", lambda: scenes.cornell_box(600, 600), 600, 600, 1000),
    ("N3 two_perlin_spheres 1200x800 500spp", lambda: scenes.two_perlin_spheres(1200, 800), 1200, 800, 500),
]
only = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--only=")]
if only:
    CONFIGS = [c for c in CONFIGS if any(c[0].startswith(o) for o in only[0].split(","))]
out = []
for name, mk, W, H, spp in CONFIGS:
    if quick:
        spp = max(1, spp // 10)
    sc = mk()
    with capi.Context(profile=sc.profile, seed=1984) as ctx:
        ctx.upload(sc).build_accel(1)
        ctx.render(W, H, min(spp, 8))  # warm-up
        ctx.clear()
        ctx.stats_reset()
        t0 = time.perf_counter()
        ctx.render(W, H, spp)
        wall = time.perf_counter() - t0
        st = ctx.stats()
    rec = {"config": name, "spp": spp, "ms": st["ms_render"], "wall_ms": 1e3 * wall,
           "Mbounces_per_s": st["segments"] / st["ms_render"] / 1e3, "seg_per_path": st["segments"] / st["paths"],
           "paths": st["paths"], "bvh_build_ms": st["ms_build"], "n_nodes": st["n_nodes"], "n_big": st["n_big_prims"],
           "smem_bytes": st["smem_bytes"], "regs": st["regs_per_thread"]}
    out.append(rec)
    print(json.dumps(rec), flush=True)
