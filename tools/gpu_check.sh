#!/bin/bash
# One GPU-box session: parity tests, smoke, bench, ncu launch list + one full capture of the dominant kernel.
# Usage (from the repo root, under gpurun):  bash tools/gpu_check.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > $OUT/gpu_$TAG.txt; nproc >> $OUT/gpu_$TAG.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee $OUT/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8 | tee $OUT/smoke_$TAG.log
timeout 600 python bench.py --steps 5 --warmup 3 2>&1 | tail -3 | tee $OUT/bench_$TAG.log
timeout 600 python bench.py --steps 5 --warmup 3 --kernel 0 --no-cpu 2>&1 | tail -1 | tee $OUT/bench_mega_$TAG.log
timeout 600 python bench.py --steps 3 --warmup 3 --workload knot64 --no-cpu 2>&1 | tail -1 | tee $OUT/bench_knot_$TAG.log
timeout 600 python bench.py --steps 3 --warmup 3 --workload knot64 --kernel 0 --no-cpu 2>&1 | tail -1 | tee $OUT/bench_knot_mega_$TAG.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee $OUT/bench_ref_$TAG.log
# launch list of the bench command (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/ncu_launches_$TAG.log 2>&1
# one full capture of the dominant kernel (skip the warm-up launches)
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_raytrace_wave -s 3 -c 1 -f -o $OUT/wave_cornell_$TAG \
    python bench.py --steps 1 --warmup 3 --no-cpu > $OUT/ncu_full_cornell_$TAG.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_raytrace_wave -s 3 -c 1 -f -o $OUT/wave_knot_$TAG \
    python bench.py --steps 1 --warmup 3 --workload knot64 --no-cpu > $OUT/ncu_full_knot_$TAG.log 2>&1
ls -la $OUT
