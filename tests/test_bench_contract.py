"""bench.py's driver contract on the CPU side: the reference arm (`--impl reference`, the oracle port timed on the host
cores) prints exactly one JSON line with the agreed keys, alone and under a 2-rank torchrun (rank 0 prints, the other
rank exits 0 without work); the b200 arm refuses to run without a GPU instead of falling back."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e", "cpu_baseline"}


def _json_lines(out: str):
    return [json.loads(ln) for ln in out.splitlines() if ln.startswith("{")]


def _check_reference_line(d, n_gpus):
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "clips/sec (16x288^2 encoder forward)" and d["unit"] == "clips/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == n_gpus and d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None
    assert d["config"]["workload"].startswith("videoprism_public_v1_base encoder forward") and d["config"]["global_batch"] == 32
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "clip" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0"], cwd=ROOT,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = _json_lines(r.stdout)
    assert len(lines) == 1 and len([ln for ln in r.stdout.splitlines() if ln.strip()]) == 1
    _check_reference_line(lines[0], 1)


def test_reference_arm_under_torchrun_only_rank0_prints():
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29641", "bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = _json_lines(r.stdout)
    assert len(lines) == 1
    _check_reference_line(lines[0], 2)


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="a GPU is present: the b200 arm would run")
def test_b200_arm_has_no_cpu_fallback():
    r = subprocess.run([sys.executable, "bench.py", "--steps", "1", "--warmup", "0"], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and not _json_lines(r.stdout)
    assert "no CPU fallback" in (r.stderr + r.stdout)
