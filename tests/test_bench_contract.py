"""bench.py contract, the parts that run without a GPU: the reference arm (the oracle's C port timed on the host cores), rank != 0
of a reference-arm launch, and the product arm refusing to run without CUDA (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMALL = ["--reads", "600", "--read-len", "2000", "--num-hashes", "128", "--cpu-seconds", "0.2", "--steps", "1", "--warmup", "0"]


def _run(args, env=None):
    e = dict(os.environ, PYTHONPATH=ROOT)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=300, env=e, cwd=ROOT)


def test_reference_arm_prints_one_json_line():
    r = _run(["--impl", "reference"] + SMALL)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "gbases_per_s_sketched_and_overlapped" and d["unit"] == "Gbases/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["value"] > 0 and d["ms_per_step"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "reads" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["mode"] == "self" and d["config"]["num_hashes"] == 128 and "workload" in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    r = _run(["--impl", "reference", "--gpus", "2"] + SMALL, env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = _run(SMALL)
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)
