"""bench.py's driver contract on the CPU side: the reference arm (the oracle timed on the host cores) prints exactly
ONE JSON line on stdout with the agreed keys, whatever the libraries write elsewhere."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--ref-sample", "400", "--workload", "C2"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("cells/sec") and d["unit"] == "cells/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0
    assert d["config"]["workload"].startswith("Xenium-scale") and "model" not in d["config"]
    assert "ref_sample" in d["config"] and "threads" in d
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and "sample" in cb and cb["value"] == d["value"]
    e2e = d["e2e"]
    assert e2e["value"] == d["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    assert d["gpu_launches"] == 0


def test_default_workload_is_the_metric_configuration():
    """BASELINE.json quotes the metric on 1M cells (configs[2]); bench.py must default to it."""
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert 'ap.add_argument("--workload", default="C3"' in src
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert "1M cells" in base["metric"] and "1M cells" in base["configs"][2]
