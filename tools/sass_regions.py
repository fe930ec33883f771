"""Static SASS instruction counts of one kernel by source file / line region (nvdisasm -g of the cubin
tools/ptxas_report.py leaves in /tmp/b200rt.cubin). Usage: python tools/sass_regions.py <mangled-name-substring>"""
import collections
import re
import subprocess
import sys

sub = sys.argv[1] if len(sys.argv) > 1 else "_Z8k_renderILi2ELb1ELi1ELb0ELb1ELb0ELb0E"
txt = subprocess.run(["nvdisasm", "-g", "-c", "/tmp/b200rt.cubin"], capture_output=True, text=True).stdout
cur = None
inside = False
cnt = collections.Counter()
for line in txt.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", line)
    if m:
        inside = sub in m.group(1)
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", line) and cur:
        cnt[cur] += 1
tot = sum(cnt.values())
print("instructions %d = %.1f KB" % (tot, tot * 16 / 1024))
byfile = collections.Counter()
for (f, l), n in cnt.items():
    byfile[f] += n
print(byfile.most_common())
for f in list(byfile)[:6]:
    top = sorted(((l, n) for (ff, l), n in cnt.items() if ff == f), key=lambda x: -x[1])[:16]
    print(f, top)
