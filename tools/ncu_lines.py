"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by source line:
   python tools/ncu_lines.py report.ncu-rep [top]  -> instructions executed / stall samples per (file, line)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = ""
hdr = None
agg = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) > 4 and r[0] == "Line No":
        hdr = r
        i_inst = hdr.index("Instructions Executed")
        i_thr = hdr.index("Thread Instructions Executed")
        i_smp = hdr.index("# Samples")
        continue
    if hdr is None or len(r) <= i_thr or not r[0]:
        continue
    try:
        agg.append((cur_file, int(r[0]), r[1].strip()[:90], int(r[i_inst]), int(r[i_thr]), int(r[i_smp])))
    except ValueError:
        pass
tot_i = sum(a[3] for a in agg) or 1
tot_s = sum(a[5] for a in agg) or 1
print("total warp-instr %.4g, samples %d" % (tot_i, tot_s))
byfile = {}
for a in agg:
    f = byfile.setdefault(a[0], [0, 0])
    f[0] += a[3]; f[1] += a[5]
for f, v in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
    print("  %-24s %5.1f %% instr  %5.1f %% samples" % (f, 100.0 * v[0] / tot_i, 100.0 * v[1] / tot_s))
print("%-22s %5s %6s %6s %5s  %s" % ("file", "line", "instr%", "smpl%", "lanes", "source"))
for a in sorted(agg, key=lambda a: -a[3])[:top]:
    print("%-22s %5d %6.2f %6.2f %5.1f  %s" % (a[0], a[1], 100.0 * a[3] / tot_i, 100.0 * a[5] / tot_s, a[4] / max(a[3], 1), a[2]))
