#!/usr/bin/env python
"""Summarise an ncu report per CUDA source line: warp instructions executed and stall samples.

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [kernel-id ':::1'] [top N]
"""
import csv, subprocess, sys, io, collections

rep = sys.argv[1]
kid = sys.argv[2] if len(sys.argv) > 2 else ":::1"
top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-id", kid],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file, hdr = None, None
agg = collections.OrderedDict()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < len(hdr):
        continue
    if r[2] != "-":       # sass rows carry an address; source rows have '-'
        continue
    d = dict(zip(hdr[4:], r[4:]))
    try:
        inst = int(d["Instructions Executed"]); samp = int(d["# Samples"])
    except ValueError:
        continue
    key = (cur_file, int(r[0]))
    a = agg.setdefault(key, dict(src=r[1].strip(), inst=0, samp=0, st=collections.Counter()))
    a["inst"] += inst; a["samp"] += samp
    for k, v in d.items():
        if k.startswith("stall_") and "Not Issued" not in k:
            try: a["st"][k[6:]] += int(v)
            except ValueError: pass
tot_i = sum(a["inst"] for a in agg.values()); tot_s = sum(a["samp"] for a in agg.values())
print(f"total warp instructions {tot_i}, samples {tot_s}")
print("== by instructions")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1]["inst"])[:top]:
    print(f"{a['inst']/tot_i*100:5.1f}% inst {a['samp']/max(tot_s,1)*100:5.1f}% smp  {f}:{ln:<4d} {a['src'][:90]}")
print("== by samples")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1]["samp"])[:top]:
    st = ", ".join(f"{k}:{v}" for k, v in a["st"].most_common(3))
    print(f"{a['samp']/max(tot_s,1)*100:5.1f}% smp {a['inst']/tot_i*100:5.1f}% inst  {f}:{ln:<4d} {a['src'][:70]}   [{st}]")
