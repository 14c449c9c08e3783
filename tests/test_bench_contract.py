"""CPU: the reference arm of bench.py (`--impl reference`: the CPU port of the reference path, the only part of
bench.py that needs no GPU) prints one JSON line with the contract's keys; under torchrun only rank 0 reports."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"}


def _run(cmd):
    p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, p.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_single_process():
    out = _run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0"])
    assert KEYS <= set(out), KEYS - set(out)
    assert out["impl"] == "reference" and out["n_gpus"] == 1 and out["gpu_launches"] == 0
    assert out["value"] > 0 and out["unit"] == "images/s" and out["higher_is_better"] is True
    assert out["cpu_baseline"]["kind"] == "port" and out["cpu_baseline"]["cores"] >= 1
    assert out["e2e"]["value"] == out["value"] and out["e2e"]["h2d_bytes_per_step"] == 0
    assert "workload" in out["config"] and "model" not in out["config"]


def test_reference_arm_under_torchrun_prints_once():
    out = _run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                "127.0.0.1", "--master-port", "29541", "bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1",
                "--warmup", "0"])
    assert out["impl"] == "reference" and out["n_gpus"] == 2
