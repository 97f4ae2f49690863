#!/bin/bash
# bench lines for several values of one environment variable: VAR=RADK_X VALS="a b" WLS="church box"
mkdir -p gpurun_out
for wl in ${WLS:-church box room}; do
  for k in $VALS; do
    env $VAR=$k python bench.py --workload $wl --steps 5 --warmup 3 > gpurun_out/bench_${wl}_${k}.json 2> gpurun_out/bench_${wl}_${k}.err
    python - <<PY
import json
d=json.load(open("gpurun_out/bench_${wl}_${k}.json"))
print("${wl} $VAR=${k}", round(d["ms_per_step"],3), "ms", {a: round(b,3) for a,b in d["roofline"]["kernel_ms_per_step"].items()}, "e2e", round(d["e2e"]["seconds_per_step"]*1e3,3))
PY
  done
done
