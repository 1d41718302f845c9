#!/bin/bash
# last visit of a round: randomised parity on the GPU, bench lines, the accumulate kernel's ncu capture and the dominant-kernel traffic
tag=${1:-fin}; o=gpurun_out; mkdir -p $o
cp drone_image_stitch_cpp_b200/lib/libdronestitch_cuda.so.sources $o/${tag}_so.sources
timeout 400 python tools/fuzz_parity.py --gpu 150 91000 > $o/${tag}_fuzz_gpu.log 2>&1; echo "fuzz rc=$?"; tail -1 $o/${tag}_fuzz_gpu.log
timeout 300 python bench.py --workload cfg2 --steps 20 --warmup 3 > $o/${tag}_bench_cfg2.json 2> $o/${tag}_bench_cfg2.err; echo "cfg2 rc=$?"
timeout 300 python bench.py --workload cfg1 --steps 20 --warmup 3 --no-cpu-baseline > $o/${tag}_bench_cfg1.json 2> $o/${tag}_bench_cfg1.err; echo "cfg1 rc=$?"
B="python bench.py --workload cfg2 --steps 3 --warmup 3 --no-cpu-baseline --no-parity"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ds_mb_accum -s 15 -c 1 -f -o $o/${tag}_accum_l1 $B > $o/${tag}_ncu_a1.log 2>&1; echo "accum rc=$?"
bash tools/gpu_traffic.sh $tag
