"""Frame times of the kernels' main regimes at reduced sample counts (a development loop, not a bench):
config 2 (sphere kernel, eight orderings), config 3 on the 968-triangle mesh (general kernel, one ordering),
config 4, the rt_next_week final scene (extended kernel, nodes-only plan). Usage: python tools/quick_perf.py [names]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from a_dive_into_ray_tracing_b200 import capi, scenes  # noqa: E402

CASES = {
    "c2": (lambda: scenes.weekend(1200, 800), 1200, 800, 500),
    "c3": (lambda: scenes.obj_room(width=800, height=800, mesh="blob968"), 800, 800, 150),
    "c4": (lambda: scenes.next_week(1200, 800), 1200, 800, 200),
    "nw": (lambda: scenes.next_week_final(800, 800), 800, 800, 200),
    "nwi": (lambda: scenes.next_week_final(800, 800, instanced=True), 800, 800, 200),
    "c3i": (lambda: scenes.obj_room(width=800, height=800, mesh="blob968", instanced=True), 800, 800, 150),
    "smoke": (lambda: scenes.cornell_smoke(600, 600), 600, 600, 200),
}
names = sys.argv[1:] or ["c2", "c3", "c4", "nw"]
for n in names:
    mk, W, H, spp = CASES[n]
    sc = mk()
    with capi.Context(profile=sc.profile, seed=1984) as ctx:
        ctx.upload(sc).build_accel(1)
        ctx.render(W, H, 8)
        best = 1e9
        for _ in range(2):
            ctx.clear()
            ctx.stats_reset()
            ctx.render(W, H, spp)
            st = ctx.stats()
            best = min(best, st["ms_render"])
    print("%-5s %dx%dx%d: %8.3f ms  %7.1f Mbounces/s  plan %d regs %d smem %d" % (
        n, W, H, spp, best, st["segments"] / best / 1e3, st["smem_plan"], st["regs_per_thread"], st["smem_bytes"]), flush=True)
