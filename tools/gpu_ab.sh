#!/bin/bash
# A/B of several builds in one session (same GPU): usage gpu_ab.sh tag lib...
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
run() { echo "== $*" | tee -a $OUT/sweep_$TAG.log
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu "$@" 2>&1 | tail -1 | python -c "
import sys,json
l=sys.stdin.read().strip()
try:
    d=json.loads(l); print(d['value'],'Mrays/s', d['ms_per_step'],'ms', 'frac',d['roofline']['frac'])
except Exception as e: print('ERR',l[-600:])" | tee -a $OUT/sweep_$TAG.log; }
for wl in cornell64 knot64; do for L in "$@"; do run --workload $wl --lib ray_tracing_b200/$L; done; done
