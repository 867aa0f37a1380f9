"""CPU: the reference arm of bench.py (`--impl reference`: the oracle port of the reference's CPU path) runs without a GPU
and prints ONE JSON line with the keys the driver contract names."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "quantized_gemm_gops" and d["unit"] == "GOPS"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert "BASELINE configs[1]" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "GOPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_cpu_reference_helpers_run():
    """The CPU legs bench.py reports beside the GPU numbers (oracle port on the host cores) work on a bounded sample."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    r = bench.cpu_linearbin_gops(batch=64, reps=1)
    assert r["gops"] > 0 and r["cores"] >= 1 and "LinearBin" in r["sample"]
    m = bench.cpu_reference_gops(32, reps=1, warmup=0)
    assert m["gops_best"] > 0 and m["batch"] == 32
