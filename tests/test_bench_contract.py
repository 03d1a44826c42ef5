"""bench.py's reference arm runs without a GPU: exactly one JSON line on stdout with the keys the driver reads."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("workload", ["landmarks15", "facefrontal", "single-psvm"])
def test_reference_arm_prints_one_json_line(built, workload):
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", workload, "--steps", "1", "--warmup", "0",
           "--ref-frames-per-core", "1", "--ref-frames-per-step", "1"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "classified_patches_per_s" and d["unit"] == "patches/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert "workload" in d["config"] and d["vs_baseline"] is None
    if workload != "single-psvm":  # the cascade arms report where the CPU time goes (pyramid | extract+hq64 | wvm | oe | svm+nms)
        assert set(cb["split_core_seconds"]) == {"pyramid", "extract+hq64", "wvm", "overlap_elimination", "svm+nms"}
        assert "pyramid" in cb


def test_b200_arm_refuses_to_run_without_a_gpu(built):
    """no CPU fallback: without a CUDA device the product arm exits with an error instead of producing a number"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"], capture_output=True, text=True,
                       timeout=600, cwd=ROOT)
    assert r.returncode != 0 and r.stdout.strip() == ""
    assert "no CUDA device" in r.stderr
