#!/bin/bash
# First GPU call of round 2 (one B200): everything written after the round-1 GPU budget was spent gets its first run on hardware.
# Before the call, here:  python tools/build_variants.py
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/r2_first_call.sh'
# Results land in gpurun_out/ (pytest / bench logs, sweep_r2.jsonl with one JSON line per configuration, sweep_r2_bvhbuild.log).
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > $OUT/gpu_r02.txt; nproc >> $OUT/gpu_r02.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $OUT/pytest_gpu_r02a.log    # parity suite incl. the round-2 file and the random searches
timeout 700 python tools/sweep.py --stage 1 2>&1 | tail -70 | tee $OUT/sweep_stage1_r02.log      # mesh kernels: 256-bit loads, vote weights, prefetches, stacks
timeout 500 python tools/sweep.py --stage 2 2>&1 | tail -50 | tee $OUT/sweep_stage2_r02.log      # census repeats, occupancy, layouts
timeout 300 python tools/sweep.py --stage 4 2>&1 | tail -20 | tee $OUT/sweep_stage4_r02.log      # TLAS against the linear model test (500 / 24 models)
timeout 300 python tools/sweep.py --stage 3 2>&1 | tail -30 | tee $OUT/sweep_stage3_r02.log      # config 2
timeout 300 python tools/bvh_build_bench.py 2>&1 | tee $OUT/sweep_r2_bvhbuild.log                # rtBuildBVH against the host builder
