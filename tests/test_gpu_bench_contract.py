"""GPU: bench.py's JSON line carries every key of the measurement contract -
run small (32^3, batch 2) so it takes seconds; the numbers themselves are not judged here."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _run(*extra):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--patch", "32", "--batch", "2",
                        "--steps", "3", "--warmup", "3", *extra],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "exactly ONE JSON line on stdout, got %d" % len(lines)
    return json.loads(lines[0])


def test_our_arm_line_has_the_contract_keys():
    j = _run()
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches",
              "roofline", "cpu_baseline", "extra"):
        assert k in j, k
    assert j["unit"] == "patches/s" and j["n_gpus"] == 1 and j["steps"] == 3 and j["value"] > 0
    assert j["higher_is_better"] is True and j["scaling"] == "weak" and j["vs_baseline"] is None
    assert "workload" in j["config"] and "model" not in j["config"]
    assert j["gpu_launches"] > 100
    e = j["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == 2 * (4 * 4 + 1) * 32 ** 3 and e["d2h_bytes_per_step"] == 4
    rf = j["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic", "kernel", "summed", "by_kernel", "top_shape"):
        assert k in rf, k
    assert rf["bound"] in ("hbm", "tensor", "fp32_fma") and 0 < rf["frac"] < 1.5 and 0 < rf["summed"] < 1.5
    assert rf["kernel"] == max(rf["by_kernel"].items(), key=lambda kv: kv[1]["ms_per_step"])[0]
    cb = j["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] > 0 and cb["sample"]
    assert set(j["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}


def test_supernet_workload_line():
    j = _run("--workload", "supernet", "--batch", "1", "--no-cpu-baseline", "--no-roofline")
    assert "supernet" in j["metric"] and j["value"] > 0 and j["e2e"]["value"] > 0
    assert j["e2e"]["h2d_bytes_per_step"] == 2 * (4 * 4 + 1) * 32 ** 3
