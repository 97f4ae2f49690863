#!/bin/bash
# A/B of the form-factor kernels on one B200 (run under gpurun): parity tests first, then bench lines.
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
for wl in church box room; do
  for k in ${KERNELS:-stream batch}; do
    RADK_FF_KERNEL=$k python bench.py --workload $wl --steps 5 --warmup 3 > gpurun_out/bench_${wl}_${k}.json 2> gpurun_out/bench_${wl}_${k}.err
    python - <<PY
import json
d=json.load(open("gpurun_out/bench_${wl}_${k}.json"))
print("${wl} ${k}", round(d["ms_per_step"],3), "ms", d["roofline"]["kernel_ms_per_step"], "e2e", round(d["e2e"]["seconds_per_step"]*1e3,3))
PY
  done
done
