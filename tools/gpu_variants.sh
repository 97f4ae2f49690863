#!/bin/bash
# A/B of library builds under hw-radiosity_b200/variants/libradk_<tag>.so:  TAGS="base x" MODELS="church box" bash tools/gpu_variants.sh
mkdir -p gpurun_out
for tag in $TAGS; do
  echo "== $tag"
  RADK_LIB=$PWD/hw-radiosity_b200/variants/libradk_$tag.so timeout ${VT:-90} python tools/ff_variant.py ${MODELS:-church box room} 2>&1 | tail -8
done | tee gpurun_out/variants_${OUT:-last}.txt
