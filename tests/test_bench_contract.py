"""CPU: the reference arm of bench.py prints exactly ONE JSON line on stdout with the keys the driver reads;
the B200 arm refuses to run without a device instead of falling back to anything."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "1"], cwd=ROOT,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "heatmaps/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1 and d["scaling"] == "weak" and d["vs_baseline"] is None
    # the unmodified reference when its tree is present (build container: /root/reference; GPU box: baseline/_ref), else the port
    from tests import refload
    assert d["cpu_baseline"]["kind"] == ("reference" if refload.find() else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"] and d["gpu_launches"] == 0


def test_reference_arm_of_the_decode_workloads():
    """--config decode_flip / decode: the same contract, their own metric (BASELINE configs[2] and [4])."""
    for config, what in (("decode_flip", "96x72, decode with flip test"), ("decode", "64x48, decode + offset")):
        r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--config", config, "--steps", "1", "--warmup", "1"], cwd=ROOT,
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        lines = [l for l in r.stdout.splitlines() if l.strip()]
        assert len(lines) == 1, r.stdout
        d = json.loads(lines[0])
        assert d["impl"] == "reference" and what in d["metric"] and d["unit"] == "heatmaps/s" and d["value"] > 0
        from tests import refload
        assert d["cpu_baseline"]["kind"] == ("reference" if refload.find() else "port")
        assert d["cpu_baseline"]["value"] == d["value"] and d["gpu_launches"] == 0
        assert "BASELINE configs" in d["config"]["workload"] and "model" not in d["config"]


def test_reference_arm_falls_back_to_the_port_without_the_tree(tmp_path):
    env = dict(os.environ, GBCODEC_NO_REFERENCE="1")
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "1"], cwd=ROOT,
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.strip()][0])
    assert d["cpu_baseline"]["kind"] == "port" and d["value"] > 0


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"], cwd=ROOT,
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_b200_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    r = subprocess.run([sys.executable, "bench.py", "--steps", "1", "--warmup", "1"], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and r.stdout.strip() == ""
    assert "no CUDA device" in r.stderr
