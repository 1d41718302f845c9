#!/bin/bash
# DRAM traffic of one launch of the dominant kernel on the bench's default workload (cfg4) and on cfg2, for roofline.traffic.
# usage (under gpurun): bash tools/gpu_traffic.sh <tag>;  then: python tools/merge_traffic.py gpurun_out/<tag>
tag=${1:-tr}; o=gpurun_out; mkdir -p $o
cp drone_image_stitch_cpp_b200/lib/libdronestitch_cuda.so.sources $o/${tag}_so.sources
for w in cfg4 cfg2; do
  timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:ds_mb_feed_l0 -s 3 -c 1 --csv \
    --log-file $o/${tag}_traffic_$w.csv python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --no-parity --no-also > $o/${tag}_traffic_$w.log 2>&1; echo "$w rc=$?"
done
