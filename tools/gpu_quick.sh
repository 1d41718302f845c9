#!/bin/bash
# quick GPU visit: parity tests (optional) + cfg2 bench line. usage: bash tools/gpu_quick.sh <tag> [notests]
tag=${1:-q}; o=gpurun_out; mkdir -p $o
if [ "$2" != "notests" ]; then timeout 900 python -m pytest tests -m gpu -x -q > $o/${tag}_gputests.log 2>&1; echo "tests rc=$?"; tail -3 $o/${tag}_gputests.log; fi
timeout 300 python bench.py --workload cfg2 --steps 20 --warmup 3 --no-cpu-baseline > $o/${tag}_bench_cfg2.json 2> $o/${tag}_bench_cfg2.err; echo "cfg2 rc=$?"
tail -c 1500 $o/${tag}_bench_cfg2.err
python - <<P
import json
d=json.loads(open('$o/${tag}_bench_cfg2.json').read().strip().split('\n')[-1])
print(round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), d['parity'].get('identical'), [(k['kernel'], round(k['ms'],4)) for k in d['roofline']['kernels']])
P
