#!/usr/bin/env python
"""Static SASS instruction count per source line / region of one kernel in a .o (nvdisasm -g).

    python tools/sass_lines.py obj.o <kernel-substring> [main.cu "name:lo-hi,..."]
"""
import collections, os, re, subprocess, sys, tempfile
obj, pat = sys.argv[1], sys.argv[2]
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, check=True, capture_output=True)
cub = [os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
cnt = collections.Counter(); cur = None; on = False
for line in txt.splitlines():
    if line.startswith("//---") and ".text." in line:
        on = pat in line
    elif "//## File" in line:
        m = re.search(r'"([^"]+)", line (\d+)', line)
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
    elif on and re.match(r"\s+/\*[0-9a-f]{4,6}\*/", line):
        cnt[cur] += 1
tot = sum(cnt.values())
print("total", tot)
if len(sys.argv) > 4:
    main = sys.argv[3]; regs = []
    for part in sys.argv[4].split(","):
        n, r = part.split(":"); lo, hi = r.split("-"); regs.append((n, int(lo), int(hi)))
    agg = collections.Counter()
    for (f, ln), c in cnt.items():
        key = f
        if f == main:
            key = "other"
            for n, lo, hi in regs:
                if lo <= ln <= hi: key = n; break
        agg[key] += c
    for k, v in agg.most_common(): print(f"{v:6d} {v/tot*100:5.1f}%  {k}")
else:
    for (f, ln), c in cnt.most_common(40): print(f"{c:6d}  {f}:{ln}")
