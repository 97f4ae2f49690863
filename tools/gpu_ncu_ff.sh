#!/bin/bash
# one ncu --set full capture (with source) of ff_batch_kernel's second launch of a converged run: WL=church TAG=x bash tools/gpu_ncu_ff.sh
WL=${WL:-church}; TAG=${TAG:-cur}
mkdir -p gpurun_out
RADK_NO_GRAPH=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:ff_batch_kernel -s 1 -c 1 -f -o gpurun_out/ncu_ff_${WL}_${TAG} python tools/run_model.py $WL 1 -1 > gpurun_out/ncu_ff_${WL}_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_ff_${WL}_${TAG}.log
python tools/ff_variant.py church box room 2>&1 | tail -3
