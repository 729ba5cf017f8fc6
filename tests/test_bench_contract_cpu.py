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


def _profile_lines():
    import glob
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r2_bench_*.json"))):
        text = open(path).read().strip()
        try:
            rec = json.loads(text)
        except json.JSONDecodeError:
            rec = json.loads(text.splitlines()[-1])
        yield os.path.basename(path), rec


def test_committed_bench_lines_are_self_consistent():
    """The bench lines committed as evidence under profiles/ carry every key of the contract and
    their derived numbers agree with each other (value = batch / step time, roofline.frac =
    achieved / peak against the measured peak, e2e with its copy sizes, clocks without thermal
    throttling, the multi-GPU lines with a green parity self-check)."""
    seen = 0
    for name, r in _profile_lines():
        seen += 1
        assert r["unit"] == "samples/s" and r["higher_is_better"] is True and r["scaling"] == "weak", name
        assert r["dtype"] == "f32" and r["data"] == "synthetic" and r["vs_baseline"] is None, name
        assert r["config"]["workload"], name
        if r.get("impl") == "reference":
            assert r["cpu_baseline"]["kind"] == "port" and r["config"].get("same_config") is False, name
            continue
        gb = r["config"]["global_batch"]
        assert abs(r["value"] - gb / (r["ms_per_step"] * 1e-3)) <= 1e-3 * r["value"], name
        assert r["steps"] >= 1 and r["warmup"] >= 3 and r["gpu_launches"] > 0, name
        e = r["e2e"]
        assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0, name
        assert e["value"] != r["value"], name
        rf = r["roofline"]
        if r["n_gpus"] == 1:        # kernel timings are taken on the single-GPU run only
            assert rf["bound"] in ("hbm", "tensor", "fp32_fma") and rf["peak"] > 0, name
            assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 2e-3, name
            assert 0 < rf["frac"] < 1.2, name
        bad = {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        assert not bad & set(r["clocks"]["reasons"]), name
        if r["n_gpus"] > 1:
            pc = r["parity_check"]
            assert pc["ok"] is True and pc["world"] == r["n_gpus"] and pc["max_rel_err"] <= pc["tol"], name
        else:
            assert r["cpu_baseline"]["value"] > 0 and r["cpu_baseline"]["cores"] >= 1, name
    assert seen >= 10
