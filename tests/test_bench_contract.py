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
