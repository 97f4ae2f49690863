#!/bin/bash
# Round-2 evidence: launch list of one converged run + one `ncu --set full` capture of each kernel of the round
# (predict / ff_batch / chain / apply / apply-commit).  Kernels are launched by the host for the captures
# (RADK_NO_GRAPH=1: same kernels, same arguments as the graph loop).   WL=church bash tools/gpu_r02_profile.sh
WL=${WL:-church}
mkdir -p gpurun_out
RADK_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_${WL}.csv python tools/run_model.py $WL 2 -1 2>&1 | grep "^run"
python tools/launch_summary.py gpurun_out/r02_launches_${WL}.csv 0.5
# skip counts: the second round of the first run (full-size launches, warm caches)
for spec in "ff_batch_kernel:1" "predict_kernel:1" "chain_kernel:1" "apply_kernel:2" "apply_kernel:3"; do
  k=${spec%%:*}; s=${spec##*:}
  RADK_NO_GRAPH=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 1 -f -o gpurun_out/r02_ncu_${WL}_${k}_s$s python tools/run_model.py $WL 1 -1 > gpurun_out/r02_ncu_${WL}_${k}_s$s.log 2>&1
  tail -2 gpurun_out/r02_ncu_${WL}_${k}_s$s.log
done
ls -la gpurun_out/r02_ncu_${WL}_*.ncu-rep
