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
P=ray_tracing_b200
for v in "" _mb4 _mb8; do run --workload cornell64 --kernel 1 --lib $P/librt_b200$v.so; done
for v in "" _pw8 _pw20 _pw24 _pw32; do run --workload knot64 --kernel 2 --lib $P/librt_b200$v.so; run --workload cornell64 --kernel 2 --lib $P/librt_b200$v.so; done
for v in "" _pw24 _pw32; do run --workload cluster4k --kernel 2 --lib $P/librt_b200$v.so; done
run --workload knot64 --kernel 2 --tail-lanes 31
run --workload knot64 --kernel 2 --tail-lanes 12
