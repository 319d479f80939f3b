"""bench.py's output contract on the CPU: the reference arm prints exactly one JSON line with the agreed keys (native
chatter such as NCCL's banner goes to stderr), and the generated stream tables are reproducible byte for byte."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, SMM_BENCH_REF_BUDGET="4")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "objective-evals/sec (all chains)" and d["unit"] == "evals/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("C2:")


def test_b200_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3", "--warmup", "3"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
    assert r.stdout.strip() == ""


def test_stream_tables_are_reproducible():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_stream_tables.py")], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    with open(os.path.join(ROOT, "include", "smm_stream_tables.h")) as f:
        assert r.stdout == f.read()


def test_reference_arm_under_torchrun_prints_one_line_from_rank_0():
    """N > 1: the driver launches the reference arm like the GPU arm; rank 0 alone works and prints, the others exit 0"""
    env = dict(os.environ, SMM_BENCH_REF_BUDGET="2")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "bench.py"),
                        "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0


def test_committed_bench_lines_carry_the_contract_keys():
    """the lines the GPU arm printed on the B200 boxes (committed under profiles/) hold every key of the bench contract:
    whole-job value, e2e with the copies declared, launches, roofline against the measured peak, clocks, the parity leg,
    the CPU baseline with its core count (N = 1) and the secondary configurations"""
    for name, n in (("bench_r2n.json", 1), ("bench_r2q_driver.json", 1), ("bench_r2o_n2.json", 2), ("bench_r2o_n8_mode3.json", 8)):
        with open(os.path.join(ROOT, "profiles", name)) as f:
            d = json.loads(f.read())
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                  "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline", "parity"):
            assert k in d, (name, k)
        assert d["n_gpus"] == n and d["metric"] == "objective-evals/sec (all chains)" and d["dtype"] == "f64"
        assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
        assert "workload" in d["config"] and "model" not in d["config"]
        assert abs(d["value"] - d["config"]["n_chains"] * d["steps"] / (d["ms_per_step"] * d["steps"] * 1e-3)) < 1e-6 * d["value"]
        e = d["e2e"]
        assert e["unit"] == "evals/s" and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] <= d["value"] * 1.001
        r = d["roofline"]
        assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"] > 0
        assert d["gpu_launches"] >= 1 and d["clocks"]["reasons"] == [] and d["clocks"]["sm_mhz"] > 0.9 * d["clocks"]["sm_max_mhz"]
        p = d["parity"]
        assert p["ok"] and p["world"] == n and p["int_mismatches"] == 0 and p["params_bit_exact"] and p["max_rel_err"] <= 1e-6
        if n == 1:
            assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] > 0
        if d.get("secondary"):
            assert set(d["secondary"]) == {"c3", "c4", "c5"} and all("value" in v for v in d["secondary"].values())
