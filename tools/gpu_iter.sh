#!/bin/bash
# Iteration session: parity tests + kernel variant sweeps on both workloads.
TAG=${1:-it}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_gpu_$TAG.log
for wl in cornell64 knot64; do
  for k in 1 2; do
    for ps in 64 96 128; do
      if [ $k = 1 ] && [ $ps != 64 ]; then continue; fi
      echo "== $wl kernel=$k pool=$ps" | tee -a $OUT/sweep_$TAG.log
      timeout 300 python bench.py --steps 3 --warmup 3 --workload $wl --kernel $k --pool-slots $ps --no-cpu 2>&1 | tail -1 | python -c "
import sys,json
l=sys.stdin.read().strip()
try:
    d=json.loads(l); print(d['value'],'Mrays/s', d['ms_per_step'],'ms', 'frac',d['roofline']['frac'])
except Exception as e: print('ERR',l[-400:])" | tee -a $OUT/sweep_$TAG.log
    done
  done
done
