#!/bin/bash
# full ncu capture of named kernels on cfg2. usage: bash tools/gpu_ncu.sh <tag> <kernel-regex>:<skip> [...]
tag=$1; shift; o=gpurun_out; mkdir -p $o
cp drone_image_stitch_cpp_b200/lib/libdronestitch_cuda.so.sources $o/${tag}_so.sources
B="python bench.py --workload ${WORKLOAD:-cfg2} --steps 3 --warmup 3 --no-cpu-baseline --no-parity"
for spec in "$@"; do
  k=${spec%%:*}; s=${spec##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 1 -f -o $o/${tag}_$k $B > $o/${tag}_ncu_$k.log 2>&1; echo "$k rc=$?"
done
