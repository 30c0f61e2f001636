#!/bin/bash
TAG=${1:-it}
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_gpu_$TAG.log
run() { echo "== $*" | tee -a $OUT/sweep_$TAG.log
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu "$@" 2>&1 | tail -1 | python -c "
import sys,json
l=sys.stdin.read().strip()
try:
    d=json.loads(l); print(d['value'],'Mrays/s', d['ms_per_step'],'ms', 'frac',d['roofline']['frac'])
except Exception as e: print('ERR',l[-600:])" | tee -a $OUT/sweep_$TAG.log; }
for ps in 64 128; do for tl in 0 4 8 16 24; do run --workload knot64 --kernel 2 --pool-slots $ps --tail-lanes $tl; done; done
for sm in 0 256 1024; do run --workload knot64 --kernel 2 --pool-slots 64 --tail-lanes 8 --smem-nodes $sm; done
run --workload cornell64 --kernel 2 --pool-slots 64
run --workload cornell64 --kernel 1
