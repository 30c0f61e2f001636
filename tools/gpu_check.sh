#!/bin/bash
# One GPU-box session: parity tests, smoke, the bench lines.  Usage (from the repo root, under gpurun): bash tools/gpu_check.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > $OUT/gpu_$TAG.txt; nproc >> $OUT/gpu_$TAG.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $OUT/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8 | tee $OUT/smoke_$TAG.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee $OUT/bench_ref_$TAG.log
timeout 600 python bench.py 2>&1 | tail -1 | tee $OUT/bench_$TAG.log
timeout 600 python bench.py --steps 5 --warmup 3 --workload knot64 --no-cpu 2>&1 | tail -1 | tee $OUT/bench_knot_$TAG.log
