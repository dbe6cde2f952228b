"""The parts of bench.py's contract that run without a GPU: the reference arm's JSON line (`--impl reference`: the unmodified
reference's RefineParticles on a bounded sample, all host threads), that only rank 0 prints it, and that the product arm refuses
to run without a B200 instead of falling back to anything."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*flags, env=None):
    e = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        e.pop(k, None)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *flags], cwd=ROOT, env=e, capture_output=True, text=True, timeout=900)


def test_reference_arm_line():
    out = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-sample", "150000")
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["unit"] == "particles/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["data"] == "synthetic" and "workload" in d["config"] and "cfg 2" in d["config"]["workload"]
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_stay_silent():
    out = run_bench("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--cpu-sample", "150000",
                    env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "29533"})
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_product_arm_has_no_cpu_path():
    import torch

    if torch.cuda.is_available():
        return  # on a GPU box the product arm is exercised by the driver itself
    out = run_bench("--steps", "1", "--warmup", "0")
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)
