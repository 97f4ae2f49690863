#!/bin/bash
# one bench.py line at N GPUs of the box: N=2 [ARGS="--no-extra"] bash tools/gpu_bench_n.sh   (gpurun --gpus N)
N=${N:-2}; TAG=${TAG:-r02_bench_n$N}
mkdir -p gpurun_out
if [ "$N" = 1 ]; then LAUNCH="python"; else LAUNCH="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"; fi
timeout ${TO:-1400} $LAUNCH bench.py --gpus $N --steps ${STEPS:-5} --warmup 3 $ARGS > gpurun_out/$TAG.json 2> gpurun_out/$TAG.err; echo rc=$?
tail -4 gpurun_out/$TAG.err
python - <<PY
import json
d=json.loads(open("gpurun_out/$TAG.json").read().strip().splitlines()[-1])
print("N=$N ms_per_step", round(d["ms_per_step"],3), "e2e_ms", round(d["e2e"]["seconds_per_step"]*1e3,3), "frac", round(d["roofline"]["frac"],3), d["parity"], d["roofline"]["kernel_ms_per_step"], "launches", d["gpu_launches"])
for k, v in (d.get("workloads") or {}).items():
    print("  ", k, {a: (round(b,3) if isinstance(b,float) else b) for a, b in v.items() if a in ("device_ms","shots","parity","error","skipped","setup_s","wall_s_total")})
PY
