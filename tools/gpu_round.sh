#!/bin/bash
# One GPU visit: parity tests, cfg2 / cfg1 bench lines, ncu launch list and full captures of the named kernels.
# usage (under gpurun): bash tools/gpu_round.sh <tag>
tag=${1:-r2}
o=gpurun_out
mkdir -p $o
cp drone_image_stitch_cpp_b200/lib/libdronestitch_cuda.so.sources $o/${tag}_so.sources
timeout 900 python -m pytest tests -m gpu -x -q > $o/${tag}_gputests.log 2>&1; echo "tests rc=$?" | tee -a $o/${tag}_gputests.log
timeout 300 python bench.py --workload cfg2 --steps 20 --warmup 3 > $o/${tag}_bench_cfg2.json 2> $o/${tag}_bench_cfg2.err; echo "cfg2 rc=$?"
timeout 300 python bench.py --workload cfg1 --steps 20 --warmup 3 --no-cpu-baseline > $o/${tag}_bench_cfg1.json 2> $o/${tag}_bench_cfg1.err; echo "cfg1 rc=$?"
B="python bench.py --workload cfg2 --steps 3 --warmup 3 --no-cpu-baseline --no-parity"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $o/${tag}_launches_cfg2.csv $B > $o/${tag}_ncu_list.log 2>&1; echo "list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ds_mb_feed_l0 -s 4 -c 1 -f -o $o/${tag}_feed_l0 $B > $o/${tag}_ncu_l0.log 2>&1; echo "l0 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ds_mb_pyrdown -s 12 -c 1 -f -o $o/${tag}_pyrdown_l1 $B > $o/${tag}_ncu_p1.log 2>&1; echo "pyrdown rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ds_mb_accum -s 15 -c 1 -f -o $o/${tag}_accum_l1 $B > $o/${tag}_ncu_a1.log 2>&1; echo "accum rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ds_mb_collapse -s 19 -c 1 -f -o $o/${tag}_collapse_l0 $B > $o/${tag}_ncu_c0.log 2>&1; echo "collapse rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ds_feather_blend -s 4 -c 1 -f -o $o/${tag}_feather python bench.py --workload cfg1 --steps 3 --warmup 3 --no-cpu-baseline --no-parity > $o/${tag}_ncu_fe.log 2>&1; echo "feather rc=$?"
for m in affine affine_seam plane_seam plane_proj homography many; do timeout 300 python tools/affine_bench.py $m 2>&1 | tail -1 >> $o/${tag}_slow_paths.jsonl; done
tail -3 $o/${tag}_gputests.log
