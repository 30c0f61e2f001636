#!/bin/bash
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
run() { echo "== $*" | tee -a $OUT/sweep_$TAG.log
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu "$@" 2>&1 | tail -1 | python -c "
import sys,json
l=sys.stdin.read().strip()
try:
    d=json.loads(l); print(d['value'],'Mrays/s', d['ms_per_step'],'ms', 'frac',d['roofline']['frac'])
except Exception as e: print('ERR',l[-600:])" | tee -a $OUT/sweep_$TAG.log; }
for wl in knot64 cluster4k soup4k; do for L in "$@"; do run --workload $wl --lib ray_tracing_b200/$L; done; done
