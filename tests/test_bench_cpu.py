"""CPU tests of bench.py's contract: the reference arm prints one JSON line with the agreed keys (timing the CPU
restatement on a bounded sample), ranks other than 0 stay silent, and our arm refuses to run without a GPU."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(args, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                          env=dict(os.environ, **(env or {})), timeout=300)


def test_reference_arm_prints_the_contract_line():
    res = run(["--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-sample-n", "300"])
    assert res.returncode == 0, res.stderr
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "evals/s" and d["higher_is_better"] is True
    assert d["metric"] == "NLL+grad evals/s at N=20k" and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "N=20000" in d["config"]["workload"]


def test_reference_arm_is_silent_on_other_ranks():
    res = run(["--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-sample-n", "300"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert res.returncode == 0 and res.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a box without a GPU")
def test_our_arm_has_no_cpu_fallback():
    res = run(["--steps", "1", "--warmup", "1"])
    assert res.returncode != 0
    assert "no CPU fallback" in (res.stderr + res.stdout)
