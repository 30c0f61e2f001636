#!/bin/bash
TAG=${1:-it}
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_gpu_$TAG.log
run() { echo "== $*" | tee -a $OUT/sweep_$TAG.log
  timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu "$@" 2>&1 | tail -1 | python -c "
import sys,json
l=sys.stdin.read().strip()
try:
    d=json.loads(l); print(d['value'],'Mrays/s', d['ms_per_step'],'ms', 'frac',d['roofline']['frac'], 'e2e', d['e2e']['value'])
except Exception as e: print('ERR',l[-600:])" | tee -a $OUT/sweep_$TAG.log; }
for wl in knot64 cluster4k soup4k; do
  run --workload $wl
  run --workload $wl --sort-rays 0
  run --workload $wl --smem-nodes 1024
done
run --workload cornell64
run --workload cornell64 --kernel 2
run --workload knot64 --pool-slots 96
run --workload cluster4k --pool-slots 96
run --workload cluster4k --tail-lanes 24
run --workload cluster4k --tail-lanes 8
