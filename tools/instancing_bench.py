"""Baked vs instanced (two-level BVH) frame times of the scenes that use translate / rotate_y, plus the cost of a
transform-only update against a full rebuild. Usage: python tools/instancing_bench.py [spp-divisor]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from a_dive_into_ray_tracing_b200 import capi, scenes  # noqa: E402
from a_dive_into_ray_tracing_b200.ctypes_defs import RT_FLAG_COUNTERS, rigid_y  # noqa: E402

div = int(sys.argv[1]) if len(sys.argv) > 1 else 10
CASES = [("cornell_box", lambda i: scenes.cornell_box(600, 600, instanced=i), 600, 600, 1000),
         ("obj_room blob968", lambda i: scenes.obj_room(width=800, height=800, mesh="blob968", instanced=i), 800, 800, 1500),
         ("next_week_final", lambda i: scenes.next_week_final(800, 800, instanced=i), 800, 800, 5000)]
for name, mk, W, H, spp in CASES:
    spp = max(8, spp // div)
    for inst in (False, True):
        sc = mk(inst)
        rec = {"scene": name, "instanced": inst, "W": W, "H": H, "spp": spp}
        for flags in (0, RT_FLAG_COUNTERS):
            with capi.Context(profile=2, seed=1984, flags=flags) as ctx:
                ctx.upload(sc).build_accel(1)
                ctx.render(W, H, 8)
                ctx.clear()
                ctx.stats_reset()
                ctx.render(W, H, spp)
                st = ctx.stats()
                if flags:
                    rec.update(box_tests_per_seg=st["box_tests"] / st["segments"], prim_tests_per_seg=st["prim_tests"] / st["segments"])
                else:
                    rec.update(ms=st["ms_render"], Mbounces_per_s=st["segments"] / st["ms_render"] / 1e3, seg_per_path=st["segments"] / st["paths"],
                               build_ms=st["ms_build"], n_nodes=st["n_nodes"], smem_plan=st["smem_plan"], regs=st["regs_per_thread"])
                    if inst:
                        moved = sc.instances.copy()
                        ms = []
                        for k in range(5):
                            moved["m"][0] = rigid_y(10.0 * k, tuple(moved["m"][0][[3, 7, 11]]))
                            ctx.update_instances(moved)
                            ms.append(ctx.stats()["ms_build"])
                        rec["update_ms_median"] = float(np.median(ms))
        print(json.dumps(rec), flush=True)
