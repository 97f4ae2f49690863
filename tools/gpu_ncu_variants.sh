#!/bin/bash
# ncu --set full capture of the form-factor kernel of each library variant: TAGS="base walk" KREGEX=ff_ bash tools/gpu_ncu_variants.sh
mkdir -p gpurun_out
for tag in $TAGS; do
  RADK_NO_GRAPH=1 RADK_LIB=$PWD/hw-radiosity_b200/variants/libradk_$tag.so timeout 200 ncu --set full --clock-control none --import-source on \
     -k regex:${KREGEX:-ff_} -s ${SKIP:-3} -c 1 -f -o gpurun_out/ncu_${tag} python tools/run_model.py ${WL:-church} 1 > gpurun_out/ncu_${tag}.log 2>&1
  tail -2 gpurun_out/ncu_${tag}.log
done
