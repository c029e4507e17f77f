"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`: the oracle port on the host cores) prints ONE
JSON line with the keys the driver reads, ranks other than 0 print nothing, the product arm refuses to run without a GPU, and the
FLOP model is SURVEY.md section 8(d)'s."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, env=e, capture_output=True, text=True,
                          timeout=600)


def test_flop_model_is_the_surveys():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.flops_per_volume(32, "s", 224) == 391_968_033_792          # SURVEY.md 8(d), per 32-slice volume
    assert bench.flops_per_volume(64, "b", 252) == 3_801_683_466_240        # ViT-B/14 @252, 64 slices


def test_reference_arm_prints_one_json_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "volumes_per_sec" and d["unit"] == "volumes/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "volumes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_are_silent():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"], env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = _run(["--steps", "1", "--warmup", "1"])
    assert r.returncode != 0 and "no CPU fallback" in r.stderr
