"""Registers / spills per kernel of libb200rt (ptxas -v of csrc/b200rt.cu). Usage: python tools/ptxas_report.py [filter]"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = os.path.join(ROOT, "a_dive_into_ray_tracing_b200", "csrc", "b200rt.cu")
out = subprocess.run(["nvcc", "-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
                      "--expt-relaxed-constexpr", "-Xptxas", "-v", "-cubin", "-o", "/tmp/b200rt.cubin", src],
                     capture_output=True, text=True).stderr
flt = sys.argv[1] if len(sys.argv) > 1 else "k_render"
cur = None
rows = []
for line in out.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        cur = {"name": m.group(1)}
        rows.append(cur)
        continue
    if cur is None:
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m:  # (out-of-line callees report their own frames too: keep the largest)
        v = tuple(int(x) for x in m.groups())
        if v[0] >= cur.get("stack", 0):
            cur["stack"], cur["st"], cur["ld"] = v
    m = re.search(r"Used (\d+) registers", line)
    if m:
        cur["regs"] = int(m.group(1))
for r in sorted(rows, key=lambda r: r["name"]):
    if flt in r["name"]:
        name = subprocess.run(["c++filt", r["name"]], capture_output=True, text=True).stdout.strip().split("(RenderParams")[0]
        print("%-70s regs %3d  stack %4d  spill st %4d ld %4d" % (name[-70:], r.get("regs", -1), r.get("stack", 0), r.get("st", 0), r.get("ld", 0)))
