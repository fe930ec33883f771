"""Large-mesh scaling of the config-3 room (closed procedural mesh subdivided N times): BVH build and
render time through the C ABI. Usage: python tools/big_mesh.py <subdivisions>   (4..7 = 5k..328k triangles)"""
import os
import sys
import time

sys.path.insert(0, os.getcwd())
from a_dive_into_ray_tracing_b200 import capi, scenes
sub = int(sys.argv[1])
sc = scenes.obj_room(width=800, height=800, subdivisions=sub)
with capi.Context(profile=sc.profile, seed=1984) as ctx:
    ctx.upload(sc).build_accel(1)
    ctx.render(800, 800, 4); ctx.clear(); ctx.stats_reset()
    ctx.render(800, 800, 64)
    st = ctx.stats()
    lin,_ = ctx.resolve()
print("obj_room sub=%d tris=%d: build %.2f ms, render %.1f ms, %.0f Mseg/s, %.1f seg/path, nodes %d, smem %d, mean %.4f" % (
    sub, len(sc.triangles), st["ms_build"], st["ms_render"], st["segments"]/st["ms_render"]/1e3, st["segments"]/st["paths"], st["n_nodes"], st["smem_bytes"], lin.mean()))
