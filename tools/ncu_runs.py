#!/usr/bin/env python
"""Sequential view of a kernel's SASS: runs of instructions with the same executed count."""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; kid = sys.argv[2] if len(sys.argv) > 2 else ":::1"; W = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-id", kid], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; seq = []
for r in rows:
    if r and r[0] == "Address": hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    d = dict(zip(hdr, r))
    try: c = int(d["Instructions Executed"])
    except ValueError: continue
    s = d["Source"].strip()
    if s.startswith("@"): s = s.split(None, 1)[1]
    seq.append((c, s.split()[0], int(d["# Samples"] or 0)))
runs = []
for c, op, smp in seq:
    key = round(c / W, 2)
    if runs and abs(runs[-1][0] - key) < 0.011: runs[-1][1].append(op); runs[-1][2] += smp
    else: runs.append([key, [op], smp])
tot = sum(k * len(o) for k, o, s in runs)
acc = 0
for k, o, s in runs:
    acc += k * len(o)
    cnt = collections.Counter(x.split(".")[0] for x in o)
    top = " ".join(f"{a}:{b}" for a, b in cnt.most_common(7))
    marks = [x for x in o if x.startswith(("BAR", "STG", "LDG", "LDGSTS", "LDGDEPBAR", "DEPBAR", "EXIT", "CALL"))]
    mk = collections.Counter(x.split(".")[0] for x in marks)
    print(f"x{k:5.2f} n={len(o):4d} cum={acc/tot*100:5.1f}% smp={s:5d} | {top} | {dict(mk) if mk else ''}")
