#!/bin/bash
# per-kernel launch list of one converged run (ncu, serialized): ROWS="64 128" WL=church bash tools/gpu_launches.sh
mkdir -p gpurun_out
for r in ${ROWS:-128}; do
  echo "== rows $r"
  RADK_FF_ROWS=$r RADK_NO_GRAPH=1 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${WL:-church}_r$r.csv python tools/run_model.py ${WL:-church} 2 ${SHOTS:--1} 2>&1 | grep "^run"
  python tools/launch_summary.py gpurun_out/launches_${WL:-church}_r$r.csv 0.5
done
