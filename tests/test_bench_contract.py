"""bench.py keeps the driver's contract: one JSON line on stdout with the agreed keys, for both arms."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e"}


def _run(args, timeout):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line_on_cpu():
    """--impl reference needs no GPU: the reference's compiled calc_dist (oracle/_ref) over rows of one BI scan."""
    d = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--n", "3000"], 300)
    assert BASE_KEYS <= set(d) and d["impl"] == "reference" and d["metric"] == "2opt_move_evals_per_sec" and d["unit"] == "evals/s"
    assert d["value"] > 1e6 and d["e2e"] == {"value": d["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb


@pytest.mark.gpu
def test_our_arm_line_on_gpu():
    d = _run(["--n", "6000", "--steps", "4", "--warmup", "3", "--no-tlo"], 600)
    assert BASE_KEYS | {"roofline", "cpu_baseline", "clocks", "gpu_launches"} <= set(d)
    assert d["n_gpus"] == 1 and d["steps"] == 4 and d["warmup"] == 3 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["gpu_launches"] >= 8 and d["value"] > 1e10 and d["e2e"]["value"] > 1e9
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    rf = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(rf) and 0 < rf["frac"] < 1.2
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
