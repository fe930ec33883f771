"""Traversal statistics of one scene on the GPU (RT_FLAG_COUNTERS kernel variant): box and
primitive tests per segment, segments per path, frame time with and without counters.
Usage: python tools/scene_stats.py <scene function in scenes.py> [W H spp]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from a_dive_into_ray_tracing_b200 import capi, scenes  # noqa: E402
from a_dive_into_ray_tracing_b200.ctypes_defs import RT_FLAG_COUNTERS  # noqa: E402

name = sys.argv[1]
W, H, spp = (int(x) for x in sys.argv[2:5]) if len(sys.argv) >= 5 else (800, 800, 64)
sc = getattr(scenes, name)(W, H) if name != "obj_room" else scenes.obj_room(width=W, height=H)
for flags in (0, RT_FLAG_COUNTERS):
    with capi.Context(profile=sc.profile, seed=1984, flags=flags) as ctx:
        ctx.upload(sc).build_accel(1)
        ctx.render(W, H, 8)
        ctx.clear()
        ctx.stats_reset()
        ctx.render(W, H, spp)
        st = ctx.stats()
    seg = st["segments"]
    print("%s %dx%dx%d %s: %.2f ms, %.0f Mseg/s, %.2f seg/path, %.1f box tests/seg, %.2f prim tests/seg, nodes %d, big %d, smem %d" % (
        name, W, H, spp, "counters" if flags else "plain", st["ms_render"], seg / st["ms_render"] / 1e3, seg / st["paths"],
        st["box_tests"] / seg, st["prim_tests"] / seg, st["n_nodes"], st["n_big_prims"], st["smem_bytes"]))
