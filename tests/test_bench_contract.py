"""bench.py's contract on a box without a GPU (CPU suite): the reference arm falls back to the oracle port and prints ONE JSON line
with the keys the driver reads; the product arm refuses to run (there is no CPU path to fall back to)."""
import json
import os
import subprocess
import sys

from util import ROOT


def _run(args, timeout=300):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout, env=env)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = _run(["--impl", "reference", "--workload", "c1", "--steps", "1", "--warmup", "0"])
    assert r.returncode == 0, r.stderr[-1500:]
    lines = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-1500:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["unit"] == "Mpixels/sec" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["config"]["workload"].startswith("C1 ") and "model" not in d["config"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]


def test_product_arm_has_no_cpu_fallback():
    r = _run(["--workload", "c1", "--steps", "1", "--warmup", "0"], timeout=120)
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout)
