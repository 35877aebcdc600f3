"""bench.py contract (CPU part): the reference arm prints one JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "T",
                        "--steps", "1", "--warmup", "1", "--cpu-sample", "2"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "genes_per_sec_gmm_fit" and line["unit"] == "genes/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["steps"] == 1 and line["warmup"] == 1
    assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "genes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]


def test_non_zero_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--workload", "T"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


import pytest


@pytest.mark.gpu
def test_gpu_arm_json_line(cuda_device):
    """The GPU arm on a reduced workload: one JSON line carrying every key of the measurement contract."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "2", "--warmup", "3", "--workload", "V",
                        "--genes", "400", "--pairs-genes", "120", "--emd-genes", "3", "--s10-genes", "6",
                        "--cpu-sample", "4"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    b = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert k in b, k
    assert b["n_gpus"] == 1 and b["steps"] == 2 and b["warmup"] == 3 and b["value"] > 0 and b["gpu_launches"] >= 2
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(b["roofline"])
    assert 0 < b["roofline"]["frac"] < 1 and b["roofline"]["unit"] == "TFLOP/s"
    assert b["e2e"]["h2d_bytes_per_step"] > 0 and b["e2e"]["d2h_bytes_per_step"] > 0 and 0 < b["e2e"]["value"] <= b["value"] * 1.05
    assert b["cpu_baseline"]["kind"] == "port" and b["cpu_baseline"]["cores"] >= 1 and b["cpu_baseline"]["value"] > 0
    assert b["pairs"]["pairs"] == 120 * 119 // 2 and b["emd"]["optimal"] == 3 and b["s10"]["genes_ok"] == 6
    assert b["mds"]["genes"] == 120 and b["mds"]["n_iter_best"] >= 1
    assert "workload" in b["config"] and "model" not in b["config"]
