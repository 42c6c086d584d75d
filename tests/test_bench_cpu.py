"""bench.py contract checks that need no GPU: the reference arm (CPU oracle port) prints the
JSON line the driver parses, and the product arm refuses to run without a CUDA device."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT,
                          capture_output=True, text=True, timeout=600)


def test_reference_arm_prints_the_contract_line():
    r = _run("--impl", "reference", "--workload", "tiny", "--steps", "2", "--warmup", "1",
             "--cpu-seconds", "0.3")
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "Mpixels/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["steps"] == 2 and line["warmup"] == 1
    assert line["config"]["workload"] == "tiny" and "sample" in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"]
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_product_arm_fails_loudly_without_cuda():
    if torch.cuda.is_available():
        return
    r = _run("--workload", "tiny", "--steps", "1", "--warmup", "0")
    assert r.returncode != 0
    assert "no CPU fallback" in r.stderr
