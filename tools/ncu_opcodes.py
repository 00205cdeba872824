#!/usr/bin/env python
"""Executed warp instructions by SASS opcode for one kernel of an ncu report."""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; kid = sys.argv[2] if len(sys.argv) > 2 else ":::1"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-id", kid], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; ops = collections.Counter(); n = 0
for r in rows:
    if r and r[0] == "Address": hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    d = dict(zip(hdr, r))
    try: c = int(d["Instructions Executed"])
    except ValueError: continue
    s = d["Source"].strip()
    if s.startswith("@"): s = s.split(None, 1)[1]
    op = s.split()[0].split(".")[0]
    ops[op] += c; n += 1
tot = sum(ops.values())
print(f"{n} SASS instructions, {tot} executed warp instructions")
for k, v in ops.most_common(40): print(f"{v/tot*100:5.1f}%  {v:>12d}  {k}")
