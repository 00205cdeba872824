#!/usr/bin/env python
"""Turn the ncu outputs a gpurun call left in gpurun_out/ into the tracked summaries under profiles/.

    python tools/make_profiles.py r01          # round tag used in the file names
"""
import csv, io, json, os, shutil, subprocess, sys, collections

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out, prof = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(prof, exist_ok=True)
run = lambda *a: subprocess.run(list(a), capture_output=True, text=True, cwd=ROOT).stdout

# 1. launch list of `bench.py --steps 3 --warmup 3 --no-e2e --no-cpu` (ncu --metrics gpu__time_duration.sum)
rows = list(csv.reader(open(os.path.join(out, "launches.csv"))))
h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[h]
ki, mi = hdr.index("Kernel Name"), hdr.index("Metric Value")
launches = [(r[ki], float(r[mi].replace(",", ""))) for r in rows[h + 1:] if len(r) > mi]
shutil.copy(os.path.join(out, "launches.csv"), os.path.join(prof, f"{tag}_bench_launches.csv"))
ours = [(n, t) for n, t in launches if n.startswith(("gbc::", "void gbc::"))]
# the last 3 launches of each of our step kernels = the timed steps
per = collections.OrderedDict()
for n, t in ours:
    per.setdefault(n.split("(")[0], []).append(t)
with open(os.path.join(prof, f"{tag}_bench_launches_summary.txt"), "w") as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none, python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu\n")
    f.write("(per-launch times are serialised and cold-cache; what must agree with bench.py is each kernel's SHARE of the step)\n\n")
    step_kernels = {k: v for k, v in per.items() if len(v) >= 6}
    tot = sum(sum(v[-3:]) / 3 for v in step_kernels.values())
    for k, v in step_kernels.items():
        m = sum(v[-3:]) / 3
        f.write(f"{m / 1e3:9.2f} us  {m / tot * 100:5.1f} %  x{len(v)}  {k}\n")
    f.write(f"{tot / 1e3:9.2f} us  100.0 %  one resident step (sum of its kernels)\n\n")
    f.write("other launches of this library (input synthesis):\n")
    for k, v in per.items():
        if k not in step_kernels:
            f.write(f"{sum(v) / len(v) / 1e3:9.2f} us  x{len(v)}  {k}\n")

# 2. full capture of the tile kernel
rep = os.path.join(out, "prof_loss.ncu-rep")
summary = run("python", "tools/ncu_summary.py", rep)
open(os.path.join(prof, f"{tag}_loss_tile_kernel_ncu_full.txt"), "w").write(
    "ncu --set full --clock-control none --import-source on -k regex:loss_tile_kernel -s 3 -c 2, same bench command\n\n" + summary)
open(os.path.join(prof, f"{tag}_loss_tile_kernel_lines.txt"), "w").write(run("python", "tools/ncu_lines.py", rep, ":::1", "40"))
open(os.path.join(prof, f"{tag}_loss_tile_kernel_opcodes.txt"), "w").write(run("python", "tools/ncu_opcodes.py", rep, ":::1"))

# 3. DRAM traffic per launch -> bench.py's roofline.traffic
raw = run("ncu", "-i", rep, "--page", "raw", "--csv")
r = list(csv.reader(io.StringIO(raw)))
hd, units = r[0], r[1]
def col(name, row):
    i = hd.index(name); v = float(row[i].replace(",", "")); u = units[i].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
vals = [col("dram__bytes_read.sum", x) + col("dram__bytes_write.sum", x) for x in r[2:]]
dur = [float(x[hd.index("gpu__time_duration.sum")].replace(",", "")) for x in r[2:]]
json.dump({"loss_kernel_dram_bytes_per_launch": sum(vals) / len(vals), "launches_captured": len(vals),
           "kernel": r[2][hd.index("Kernel Name")], "gpu_time_us_under_ncu": sum(dur) / len(dur),
           "source": f"profiles/{tag}_loss_tile_kernel_ncu_full.txt (dram__bytes_read.sum + dram__bytes_write.sum)"},
          open(os.path.join(prof, "traffic.json"), "w"), indent=1)
print(open(os.path.join(prof, f"{tag}_bench_launches_summary.txt")).read())
print(open(os.path.join(prof, "traffic.json")).read())
