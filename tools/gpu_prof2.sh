#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
bash tools/gpu_prof.sh k2b_knot knot64 2
bash tools/gpu_prof.sh k2b_cluster cluster4k 2
bash tools/gpu_prof.sh k1c_cornell cornell64 1
for v in _pw20 _pw28; do for wl in knot64 cluster4k cornell64; do echo "== $wl $v"; timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu --workload $wl --kernel 2 --lib ray_tracing_b200/librt_b200$v.so 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['value'],'Mrays/s', d['ms_per_step'],'ms', 'frac',d['roofline']['frac'])"; done; done 2>&1 | tee $OUT/sweep_it6.log
