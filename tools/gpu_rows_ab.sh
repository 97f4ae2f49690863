#!/bin/bash
# form-factor cache rows A/B (RADK_FF_ROWS): ROWS="32 64 128" MODELS="church box room" bash tools/gpu_rows_ab.sh
mkdir -p gpurun_out
for r in ${ROWS:-32 64 128}; do
  echo "== rows $r"
  RADK_FF_ROWS=$r timeout ${VT:-120} python tools/ff_variant.py ${MODELS:-church box room} 2>&1 | tail -6
done | tee gpurun_out/rows_${OUT:-ab}.txt
