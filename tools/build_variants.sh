#!/bin/bash
# builds library variants for A/B runs: tools/build_variants.sh "tag:-DFOO -DBAR=1" "tag2:..."
cd "$(dirname "$0")/../hw-radiosity_b200" || exit 1
for spec in "$@"; do
  tag=${spec%%:*}; defs=${spec#*:}; [ "$defs" = "$spec" ] && defs=""
  make variant TAG=$tag DEFS="$defs" > /dev/null || { echo "build of $tag failed"; exit 1; }
  grep -h -A2 "ff_stream_kernel\|ff_walk_kernel\|ff_batch_kernel" variants/ptxas_$tag.log | grep "Used" | head -3 | sed "s/^/$tag: /"
done
