#!/usr/bin/env python
"""BASELINE config 3 (blocks.tri, adaptive splitting enabled, -adapt_area 5): device time of the split-enabled loop on
one GPU next to the split-enabled reference restatement (oracle, single thread - the reference's own split path races
with more).  Prints one JSON line.  Not the headline bench (bench.py); a parity-config timing for profiles/."""
import gzip, json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import hw_radiosity_b200 as hr
from oracle import pyoracle as po
from conftest import oracle_scene

name = sys.argv[1] if len(sys.argv) > 1 else "blocks"
with tempfile.TemporaryDirectory() as tmp:
    tri = os.path.join(tmp, name + ".tri")
    open(tri, "wb").write(gzip.open(os.path.join(ROOT, "tests", "golden", "models", name + ".tri.gz")).read())
    m = hr.Model(tri, prepartition=False, tri_limit=1000000, sample_area=-1.0)
times = []
for rep in range(4):
    ctx = hr.Context(m, split_enabled=1, area_limit=5.0, sample_area=m.sample_area, tri_limit=1000000)
    t0 = time.perf_counter()
    done, conv = ctx.step(-1)
    wall = time.perf_counter() - t0
    times.append((ctx.device_ms, wall, done, conv, ctx.trinum))
    st = ctx.stats()
    del ctx
dev_ms, wall, done, conv, trinum = min(times[1:])
t0 = time.perf_counter()
sc = oracle_scene(po, m, split_enabled=True, capacity=1000000, threads=1, area_limit=5.0, tri_limit=1000000)
seq = sc.run()
cpu_s = time.perf_counter() - t0
print(json.dumps({"workload": f"model/{name}.tri -adapt_area 5, adaptive splitting enabled, to convergence", "n_gpus": 1,
                  "device_ms": dev_ms, "wall_ms": wall * 1e3, "shots": done, "converged": conv, "slots": trinum,
                  "rays": st["rays"], "splits": st["splits"], "kernel_launches": st["kernel_launches"],
                  "cpu_oracle_split_single_thread_s": cpu_s, "cpu_shots": len(seq), "cpu_slots": sc.trinum}))
