"""bench.py's contract, as far as a machine without a GPU can check it: the reference arm (the CPU oracle port, the one
leg that may execute oracle/) prints one JSON line with the keys the driver reads; the product arm refuses to run
without a B200 instead of computing anything on the host; both arms name the workload with the same string."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench(*args):
    return subprocess.run([sys.executable, "bench.py", *args], cwd=ROOT, capture_output=True, text=True, timeout=900)


def test_reference_arm_prints_the_contract_line():
    r = _bench("--impl", "reference", "--steps", "1", "--warmup", "1", "--batch", "128")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"].startswith("policy+value evals/sec") and line["unit"] == "evals/s"
    assert line["value"] > 0 and line["n_gpus"] == 1 and line["steps"] == 1 and line["warmup"] >= 3      # W >= 3 is enforced
    assert line["higher_is_better"] is True and line["scaling"] == "weak" and line["vs_baseline"] is None
    assert line["dtype"] == "f32" and line["data"] == "synthetic" and "batch 128" in line["config"]["workload"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["unit"] == "evals/s" and cb["sample"]
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["unit"] == "evals/s" and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    assert abs(line["ms_per_step"] * 1e-3 * line["value"] - 128) < 1e-6 * 128


def test_both_arms_name_the_workload_alike():
    sys.path.insert(0, ROOT)
    import bench
    w = bench.workload(7, 128, 1024)
    assert "7x128" in w and "batch 1024" in w and "configs[1]" in w
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count('"workload": workload(') >= 3                     # reference arm, product arm, scaled tower: one function


def test_product_arm_has_no_cpu_fallback():
    if torch.cuda.is_available():
        return
    r = _bench("--steps", "1", "--warmup", "1")
    assert r.returncode != 0 and "no CPU fallback" in r.stderr
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]
