#!/usr/bin/env python
"""Small driver for profiling: python tools/run_model.py church [runs] [shots] - converges the model `runs` times
(rewind in between) through the C ABI and prints the device time and the round counters of each run."""
import gzip, os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hw_radiosity_b200 as hr

name = sys.argv[1] if len(sys.argv) > 1 else "church"
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 3
shots = int(sys.argv[3]) if len(sys.argv) > 3 else -1
with tempfile.TemporaryDirectory() as tmp:
    tri = os.path.join(tmp, "m.tri")
    open(tri, "wb").write(gzip.open(os.path.join(ROOT, "tests", "golden", "models", name + ".tri.gz")).read())
    m = hr.Model(tri)
ctx = hr.Context(m)
for i in range(runs):
    ctx.rewind()
    t0 = ctx.device_ms
    done, conv = ctx.step(shots)
    st = ctx.stats()
    print(f"run {i}: {done} shots conv={conv} {ctx.device_ms - t0:.3f} ms rounds={st['rounds']} rows={st['ff_rows_filled']} "
          f"scheduled={st['shots_scheduled']} reapplied={st['rounds_reapplied']}", flush=True)
import hashlib
import numpy as np
accB, deltaB, parent = ctx.download()
print("digest", hashlib.sha256(np.ascontiguousarray(accB[parent >= 0]).tobytes()).hexdigest()[:12], "log tail", ctx.shooter_log()[-5:], flush=True)
