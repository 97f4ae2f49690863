#!/usr/bin/env python
"""A/B harness for kernel variants: python tools/ff_variant.py [models...] (RADK_LIB selects the library build).
Per model: converges it 1 + RUNS times through the C ABI (profiled rounds: CUDA events around the kernel classes),
prints the best per-class device times, the plain step time, and whether the final state equals the committed golden
(bit-exact SHA-256 of accB / deltaB) - a variant that changes results is rejected on the spot."""
import gzip, hashlib, json, os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import hw_radiosity_b200 as hr

RUNS = int(os.environ.get("RUNS", "3"))
CAP = {"room": 200, "box": 500, "hall": 100}
gold = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))["models"]
sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
for name in (sys.argv[1:] or ["church"]):
    with tempfile.TemporaryDirectory() as tmp:
        tri = os.path.join(tmp, "m.tri")
        open(tri, "wb").write(gzip.open(os.path.join(ROOT, "tests", "golden", "models", name + ".tri.gz")).read())
        m = hr.Model(tri)
    ctx = hr.Context(m)
    cap = CAP.get(name, -1)
    best, step_ms = None, []
    for i in range(RUNS + 1):
        ctx.rewind()
        k = ctx.profile(cap)
        if i and (best is None or sum(k["ms"].values()) < sum(best["ms"].values())):
            best = k
        ctx.rewind()
        t0 = ctx.device_ms
        done, conv = ctx.step(cap)
        if i:
            step_ms.append(ctx.device_ms - t0)
    accB, deltaB, parent = ctx.download()
    live = parent >= 0
    ok = None
    g = gold.get(name, {}).get("full") if cap < 0 else None
    if g:
        ok = sha(accB[live]) == g["accB_sha"] and sha(deltaB[live]) == g["deltaB_sha"] and done == g["shots"]
    st = ctx.stats()
    print(f"{name:8s} step {min(step_ms):8.3f} ms  " + "  ".join(f"{a} {b:7.3f}" for a, b in best["ms"].items()) +
          f"  shots {done} rounds/run {st['rounds'] // (2 * (RUNS + 1))} rows/run {st['ff_rows_filled'] // (2 * (RUNS + 1))} parity {ok}"
          f"  digest {sha(accB[live])[:10]}", flush=True)
    ctx.close()
