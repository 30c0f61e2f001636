#!/bin/bash
TAG=${1:-it}
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_gpu_$TAG.log
run() { echo "== $*" | tee -a $OUT/sweep_$TAG.log
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu "$@" 2>&1 | tail -1 | python -c "
import sys,json
l=sys.stdin.read().strip()
try:
    d=json.loads(l); print(d['value'],'Mrays/s', d['ms_per_step'],'ms', 'frac',d['roofline']['frac'], 'e2e', d['e2e']['value'])
except Exception as e: print('ERR',l[-600:])" | tee -a $OUT/sweep_$TAG.log; }
P=ray_tracing_b200
for wl in knot64 cluster4k soup4k; do run --workload $wl; run --workload $wl --lib $P/librt_b200_pf.so; done
run --workload cornell64
