#!/usr/bin/env python
"""Turn what tools/gpu_final.sh left in gpurun_out/ into the tracked round-2 summaries under profiles/.

    python tools/make_profiles_r02.py
"""
import csv, io, json, os, shutil, subprocess, collections

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out, prof = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
run = lambda *a: subprocess.run(list(a), capture_output=True, text=True, cwd=ROOT).stdout
for src, dst in (("bench_n1.json", "r02_bench_n1.json"), ("bench_ref.json", "r02_bench_reference_n1.json"),
                 ("pytest_gpu.log", "r02_pytest_gpu.log"), ("smoke.log", "r02_smoke.log"), ("launches.csv", "r02_bench_launches.csv")):
    shutil.copy(os.path.join(out, src), os.path.join(prof, dst))

# launch list of `bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-extras`
rows = list(csv.reader(open(os.path.join(out, "launches.csv"))))
h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[h]
ki, mi = hdr.index("Kernel Name"), hdr.index("Metric Value")
launches = [(r[ki], float(r[mi].replace(",", ""))) for r in rows[h + 1:] if len(r) > mi]
per = collections.OrderedDict()
for n, t in launches:
    per.setdefault(n.split("(")[0], []).append(t)
ours = {k: v for k, v in per.items() if k.startswith(("gbc::", "void gbc::"))}
step = {k: v for k, v in ours.items() if len(v) >= 6}
bench = json.load(open(os.path.join(out, "bench_n1.json")))
with open(os.path.join(prof, "r02_bench_launches_summary.txt"), "w") as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none, python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-extras\n")
    f.write("(per-launch times are serialised and cold-cache; what must agree with bench.py is each kernel's SHARE of the step)\n\n")
    tot = sum(sum(v[-3:]) / 3 for v in step.values())
    for k, v in step.items():
        m = sum(v[-3:]) / 3
        f.write(f"{m / 1e3:9.2f} us  {m / tot * 100:5.1f} %  x{len(v)}  {k[:90]}\n")
    f.write(f"{tot / 1e3:9.2f} us  100.0 %  one resident step (sum of its kernels)\n")
    r = bench["roofline"]
    f.write(f"bench.py on the same box (no profiler): step kernel {r['kernel_ms']:.4f} of {bench['ms_per_step']:.4f} ms = {r['kernel_ms'] / bench['ms_per_step'] * 100:.1f} %\n\n")
    f.write("other launches in the capture:\n")
    for k, v in per.items():
        if k not in step:
            f.write(f"{sum(v) / len(v) / 1e3:9.2f} us  x{len(v)}  {k[:110]}\n")

# the step kernel's full capture
rep = os.path.join(out, "prof_step.ncu-rep")
head = "ncu --set full --clock-control none --import-source on -k regex:step_pipe_kernel -s 3 -c 1 tools/bench_loss 1024 17 64 48 5 3\n\n"
open(os.path.join(prof, "r02_step_pipe_kernel_ncu_full.txt"), "w").write(head + run("python", "tools/ncu_summary.py", rep))
lines = run("python", "tools/ncu_lines.py", rep, ":::1", "45")
raw = run("ncu", "-i", rep, "--page", "raw", "--csv")
r = list(csv.reader(io.StringIO(raw)))
hd, units, row = r[0], r[1], r[2]
def val(name):
    i = hd.index(name)
    return row[i] + " " + units[i]
extra = "\n".join(f"{m} {val(m)}" for m in ("sm__icc_request_hit_rate.pct", "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed",
                                             "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                             "smsp__inst_executed.sum", "gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active") if m in hd)
hot = run("python", "tools/ncu_hot.py", rep, "17408")
open(os.path.join(prof, "r02_step_pipe_kernel_lines.txt"), "w").write(head + lines + "\n" + extra + "\n\nhot instruction footprint (tools/ncu_hot.py, tiles = 17408):\n" + hot)
def col(name):
    i = hd.index(name); v = float(row[i].replace(",", "")); u = units[i].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
json.dump({"loss_kernel_dram_bytes_per_launch": col("dram__bytes_read.sum") + col("dram__bytes_write.sum"), "launches_captured": 1,
           "kernel": "step_pipe_kernel<12,16,4,3,1>", "gpu_time_us_under_ncu": float(row[hd.index("gpu__time_duration.sum")].replace(",", "")),
           "source": "profiles/r02_step_pipe_kernel_ncu_full.txt (dram__bytes_read.sum + dram__bytes_write.sum)"},
          open(os.path.join(prof, "traffic.json"), "w"), indent=1)
# the three designs on one box
with open(os.path.join(prof, "r02_step_designs.jsonl"), "a") as f:
    for l in open(os.path.join(out, "designs.log")):
        d = json.loads(l); d["session"] = "final build of round 2 (variance tile first, target row carried, merged pass B)"
        f.write(json.dumps(d) + "\n")
print(open(os.path.join(prof, "r02_bench_launches_summary.txt")).read())
print(open(os.path.join(prof, "traffic.json")).read())
print(extra); print(hot)
