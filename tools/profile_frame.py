"""One config-2 frame (Weekend scene, 1200x800) through the C ABI — the short command
profiled under ncu (profiles/). Usage: python tools/profile_frame.py [spp] [profile-scene]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from a_dive_into_ray_tracing_b200 import capi, scenes  # noqa: E402

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 500
which = sys.argv[2] if len(sys.argv) > 2 else "weekend"
W, H = (1200, 800) if which in ("weekend", "final_cu", "next_week", "two_perlin_spheres") else (800, 800)
if which == "obj_room":
    sc = scenes.obj_room(width=W, height=H)
else:
    sc = getattr(scenes, which)(W, H)  # any scene function of scenes.py taking (width, height)
with capi.Context(profile=sc.profile, seed=1984) as ctx:
    ctx.upload(sc).build_accel(1)
    for _ in range(2):
        ctx.clear()
        ctx.render(W, H, spp)
    st = ctx.stats()
    print("%s %dx%d %d spp: render %.3f ms, %.1f Mbounces/s, %.3f seg/path, regs %d, smem %d" % (
        which, W, H, spp, st["ms_render"], st["segments"] / 2 / st["ms_render"] / 1e3,
        st["segments"] / st["paths"], st["regs_per_thread"], st["smem_bytes"]))
