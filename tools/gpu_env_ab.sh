#!/bin/bash
# bench lines for several environment settings (run under gpurun):
#   CONFIGS="base: tagged:RADK_SHOTS_TAGGED=1 t2:RADK_SHOTS_TAGGED=1,RADK_SHOTS_RPT=2" WLS="church box" bash tools/gpu_env_ab.sh
mkdir -p gpurun_out
for wl in ${WLS:-church box room}; do
  for cfg in $CONFIGS; do
    name=${cfg%%:*}; envs=$(echo "${cfg#*:}" | tr ',' ' ')
    env $envs python bench.py --workload $wl --steps ${STEPS:-5} --warmup 3 > gpurun_out/bench_${wl}_${name}.json 2> gpurun_out/bench_${wl}_${name}.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${wl}_${name}.json"))
    print("${wl} ${name}", round(d["ms_per_step"],3), "ms", {a: round(b,3) for a,b in d["roofline"]["kernel_ms_per_step"].items()}, "e2e", round(d["e2e"]["seconds_per_step"]*1e3,3))
except Exception as e:
    print("${wl} ${name} FAILED", e); print(open("gpurun_out/bench_${wl}_${name}.err").read()[-600:])
PY
  done
done
