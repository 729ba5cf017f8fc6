"""bench.py's reference arm (the CPU restatement of the reference op sequence, timed on the host
cores) runs without a GPU and prints the contract's JSON line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--steps", "1", "--warmup", "1", "--cpu-batch", "256", "--cpu-row-cap", "20000"],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    r = json.loads(lines[0])
    assert r["impl"] == "reference" and r["metric"] == "DLRM train samples/sec"
    assert r["unit"] == "samples/s" and r["higher_is_better"] is True and r["value"] > 0
    assert r["config"]["workload"] == "dlrm_criteo_synthetic" and r["config"]["embed_dim"] == 128
    cb = r["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == r["value"] and cb["sample"]
    assert r["e2e"] == {"value": r["value"], "unit": "samples/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}
    assert r["vs_baseline"] is None and r["dtype"] == "f32" and r["data"] == "synthetic"


def test_b200_arm_refuses_to_run_without_cuda():
    """No CPU fallback: without a CUDA device the product arm exits with an error."""
    import torch
    if torch.cuda.is_available():
        return
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert res.returncode != 0
    assert "no CUDA device" in (res.stderr + res.stdout)
