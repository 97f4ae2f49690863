#!/bin/bash
# one ncu --set full capture of the form-factor kernel (first full-size launch of a church run)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:${KREGEX:-ff_batch} -s 3 -c 1 -f -o gpurun_out/${OUT:-prof_ff} python bench.py --steps 1 --warmup 1 --workload ${WL:-church} > gpurun_out/ncu_run.log 2>&1
