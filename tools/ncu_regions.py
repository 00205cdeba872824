#!/usr/bin/env python
"""Aggregate an ncu source page (CUDA view) by line ranges of the main file: instructions, samples, stall reasons.

    python tools/ncu_regions.py rep.ncu-rep main.cu "name:lo-hi,name:lo-hi,..."
Lines of other files (inlined helpers) are listed per file.
"""
import csv, subprocess, sys, io, collections
rep, main, spec = sys.argv[1], sys.argv[2], sys.argv[3]
regions = []
for part in spec.split(","):
    n, r = part.split(":"); lo, hi = r.split("-"); regions.append((n, int(lo), int(hi)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-id", ":::1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur, hdr = None, None
agg = collections.OrderedDict()
tot = collections.Counter()
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) or r[0] == "Function Name": continue
    if r[2] != "-": continue          # SASS rows carry an address
    d = dict(zip(hdr, r))
    try: inst = int(d["Instructions Executed"]); samp = int(d["# Samples"]); ln = int(d["Line No"])
    except (ValueError, KeyError): continue
    key = cur
    if cur == main:
        key = "other-lines"
        for n, lo, hi in regions:
            if lo <= ln <= hi: key = n; break
    a = agg.setdefault(key, dict(inst=0, samp=0, st=collections.Counter()))
    a["inst"] += inst; a["samp"] += samp
    for k, v in d.items():
        if k.startswith("stall_") and "Not Issued" not in k:
            try: a["st"][k[6:]] += int(v); tot[k[6:]] += int(v)
            except ValueError: pass
ti = sum(a["inst"] for a in agg.values()); ts = sum(a["samp"] for a in agg.values())
print(f"total warp instructions {ti}, samples {ts}")
print("stall totals: " + ", ".join(f"{k}:{v/ts*100:.1f}%" for k, v in tot.most_common(12)))
for k, a in agg.items():
    st = ", ".join(f"{x}:{v}" for x, v in a["st"].most_common(4))
    print(f"{a['inst']/ti*100:5.1f}% inst {a['samp']/ts*100:5.1f}% smp  {k:28s} [{st}]")
