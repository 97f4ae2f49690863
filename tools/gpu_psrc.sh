#!/bin/bash
# -parallel_src on the batched path: parity tests + timing of the reference's go.sh configuration next to the generic kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py tests/test_gpu_edges.py -x -q -m gpu 2>&1 | tail -15
python - <<'PY'
import gzip, os, sys, tempfile, time
sys.path.insert(0, os.getcwd())
import hw_radiosity_b200 as hr
with tempfile.TemporaryDirectory() as tmp:
    tri = os.path.join(tmp, "m.tri")
    open(tri, "wb").write(gzip.open("tests/golden/models/room.tri.gz").read())
    m = hr.Model(tri, tri_limit=40000)
    m2 = hr.Model(tri)
for model, tag in ((m, "room -triangle 40000 -parallel_src 10 (go.sh)"), (m2, "room -parallel_src 10")):
    for kw in (dict(parallel_src=10), dict(parallel_src=10, per_shot_kernels=1)):
        ctx = hr.Context(model, **kw)
        for rep in range(2):
            ctx.rewind()
            t0 = ctx.device_ms
            done, conv = ctx.step(-1 if model is m else 300)
            print(tag, kw, "iterations", done, "conv", conv, "device ms", round(ctx.device_ms - t0, 3), flush=True)
        ctx.close()
PY
