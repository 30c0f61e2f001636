#!/bin/bash
TAG=${1:-r01f}
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $OUT/pytest_gpu_$TAG.log
run() { echo "== $*" | tee -a $OUT/sweep_$TAG.log
  timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu "$@" 2>&1 | tail -1 | python -c "
import sys,json
l=sys.stdin.read().strip()
try:
    d=json.loads(l); print(d['value'],'Mrays/s', d['ms_per_step'],'ms', 'frac',d['roofline']['frac'], 'e2e', d['e2e']['value'])
except Exception as e: print('ERR',l[-600:])" | tee -a $OUT/sweep_$TAG.log; }
run --workload soup4k
run --workload soup4k --kernel 1
run --workload knot64
run --workload cornell64
bash tools/gpu_prof.sh ${TAG}_cornell cornell64 1
bash tools/gpu_prof.sh ${TAG}_knot knot64 2
bash tools/gpu_prof.sh ${TAG}_soup soup4k 2
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/ncu_launches_$TAG.log 2>&1
