#!/bin/bash
# bench lines for prebuilt library variants hw-radiosity_b200/variants/libradk_<tag>.so (scratch copy on the GPU box)
mkdir -p gpurun_out
for wl in ${WLS:-church box room}; do
  for tag in $TAGS; do
    cp hw-radiosity_b200/variants/libradk_$tag.so hw-radiosity_b200/libradk.so
    python bench.py --workload $wl --steps 5 --warmup 3 > gpurun_out/bench_${wl}_${tag}.json 2> gpurun_out/bench_${wl}_${tag}.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${wl}_${tag}.json"))
    print("${wl} ${tag}", round(d["ms_per_step"],3), "ms", {a: round(b,3) for a,b in d["roofline"]["kernel_ms_per_step"].items()})
except Exception as e:
    print("${wl} ${tag} FAILED", e)
PY
  done
done
