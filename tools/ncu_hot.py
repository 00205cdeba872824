#!/usr/bin/env python
"""Hot instruction footprint of a kernel from an ncu report with source (SASS view): how many distinct SASS instructions were
executed at least `frac` x (tiles) times, in KB (16 B each), split by the role (warp id range is not in the report, so the
split is by source region of the main file).

    python tools/ncu_hot.py rep.ncu-rep <tiles> [main.cu "name:lo-hi,..."]
"""
import csv, io, subprocess, sys, collections
rep, tiles = sys.argv[1], int(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-id", ":::1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
ex = []
for r in rows:
    if r and r[0] in ("Address", "#"): hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    d = dict(zip(hdr, r))
    try: ex.append((d.get("Address", ""), d.get("Source", ""), int(d["Instructions Executed"]), int(d.get("# Samples", "0") or 0)))
    except (ValueError, KeyError): pass
print("SASS instructions", len(ex), "=", len(ex) * 16 / 1024, "KB")
for name, lo in (("never", 0), ("> 0", 1), (">= tiles/100", tiles // 100), (">= tiles/10", tiles // 10), (">= tiles/2", tiles // 2), (">= tiles", tiles), (">= 3 tiles", 3 * tiles), (">= 12 tiles", 12 * tiles)):
    n = sum(1 for a in ex if (a[2] == 0 if name == "never" else a[2] >= lo))
    print(f"{name:14s} {n:6d} instr {n * 16 / 1024:7.1f} KB")
# contiguous hot runs: 128-byte lines (8 instructions) touched by instructions executed >= tiles/2
lines = set()
for i, a in enumerate(ex):
    if a[2] >= tiles // 2: lines.add(i // 8)
print("128-byte lines holding an instruction executed >= tiles/2 times:", len(lines), "=", len(lines) * 128 / 1024, "KB")
