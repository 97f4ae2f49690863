"""bench.py prints exactly one JSON line with the keys the build brief's measurement contract names."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e"}


def run_bench(*args, timeout=900):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout[-2000:]          # stdout is exactly one JSON line
    return json.loads(lines[0])


def test_reference_arm_line():
    """`--impl reference`: the unmodified reference's CPU loop on this box's cores (oracle/_ref), same metric and config."""
    d = run_bench("--impl", "reference", "--workload", "box", "--steps", "1", "--warmup", "0")
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "visibility_rays_per_sec" and d["unit"] == "rays/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["n_gpus"] == 1 and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.gpu
def test_own_arm_line():
    d = run_bench("--workload", "box", "--steps", "2", "--warmup", "3")
    assert BASE_KEYS | {"gpu_launches", "clocks", "roofline", "cpu_baseline"} <= set(d)
    assert d["metric"] == "visibility_rays_per_sec" and d["unit"] == "rays/s" and d["dtype"] == "f32"
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["gpu_launches"] > 0 and d["warmup"] >= 3
    assert d["vs_baseline"] is None and d["scaling"] in ("strong", "weak") and "workload" in d["config"]
    e = d["e2e"]
    assert e["unit"] == d["unit"] and 0 < e["value"] < d["value"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] == "GB/s" and r["peak"] > 0
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and "traffic" in r
    cb = d["cpu_baseline"]
    assert cb["value"] > 0 and cb["cores"] >= 1 and cb["kind"] in ("reference", "port") and cb["sample"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
