"""bench.py's reference arm runs on the CPU (it times the oracle port), so its JSON contract can be checked here:
one line, the keys the driver reads, the metric/unit of BASELINE.json, `impl: reference`, zero-byte e2e."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--frames", "200",
                                   "--beams", "100", "--steps", "2", "--warmup", "1"], text=True, timeout=600, cwd=ROOT)
    lines = [ln for ln in out.splitlines() if ln.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["value"] > 0 and d["unit"] == "residual evals/s" and "workload" in d["config"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and cb["cores"] >= 1 and cb["sample"]
    # the metric is the one BASELINE.json names
    assert "residual" in d["metric"].lower() and "residual" in json.dumps(base).lower()


def test_both_arms_print_the_same_config():
    """The driver compares the two arms' `config`: both come from one function (the reference arm's bounded sample is described
    in its cpu_baseline.sample, not in config)."""
    sys.path.insert(0, ROOT)
    import argparse

    import bench

    a = argparse.Namespace(frames=10_000, beams=1_000)
    c = bench.workload_config(a, 1)
    assert set(c) == {"workload"} and "10000 frames x 1000 points" in c["workload"]
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count("workload_config(args, world)") == 3  # the definition + run_reference + run_ours


def test_ground_truth_pose_error_is_accurate_near_zero():
    """bench.py's check block measures the distance to the generator's ground truth: it must resolve 1e-12 rad (an arccos of
    the trace would stop at 1e-8)."""
    sys.path.insert(0, ROOT)
    import numpy as np

    import bench
    from oracle import oracle as O

    gt = O.ground_truth()[1]
    ang, dt = bench.pose_error_vs_ground_truth(gt)
    assert ang < 1e-15 and dt < 1e-15
    x = O.pose_plus(gt, np.array([0, 0, 1e-9, 2e-12, 0, 0]))
    ang, dt = bench.pose_error_vs_ground_truth(x)
    assert abs(ang - 2e-12) < 1e-14 and abs(dt - 1e-9) < 1e-14


def test_reference_arm_under_torchrun_only_rank0_works():
    """`bench.py --impl reference` launched with N ranks: rank 0 alone runs and prints, the others exit 0 without work."""
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ""
