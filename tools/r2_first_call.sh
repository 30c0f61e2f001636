#!/bin/bash
# First GPU call of round 2 (one B200, ~35 min): everything that was written after the round-1 GPU budget was spent gets its first
# run on hardware, in order of how much depends on it.  Before the call, here:  python tools/build_variants.py
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/r2_first_call.sh'
# Results land in gpurun_out/ (pytest / bench logs, sweep_r2.jsonl, sweep_r2_bvhbuild.log).
OUT=gpurun_out; mkdir -p $OUT
bash tools/gpu_check.sh r02                                               # parity suite incl. the round-2 file, smoke, bench lines
# the randomised searches against the real kernels (the tools take any library with the ABI): first random GPU-vs-oracle parity
timeout 600 python tools/simt_fuzz.py --cases 600 --seed 201 --far 0.05 --spheres 0.08 --odd 0.1 --lib ray_tracing_b200/librt_b200.so 2>&1 | tail -4 | tee $OUT/fuzz_gpu_r02.log
timeout 300 python tools/simt_fuzz_session.py --cases 150 --seed 202 --lib ray_tracing_b200/librt_b200.so 2>&1 | tail -3 | tee -a $OUT/fuzz_gpu_r02.log
timeout 300 python tools/simt_fuzz_bvh.py --cases 400 --seed 203 --lib ray_tracing_b200/librt_b200.so 2>&1 | tail -3 | tee -a $OUT/fuzz_gpu_r02.log
timeout 600 python tools/sweep.py --stage 1 2>&1 | tail -60 | tee $OUT/sweep_stage1_r02.log      # mesh kernels: 256-bit loads, vote weights, prefetches, stacks
timeout 300 python tools/sweep.py --stage 4 2>&1 | tail -20 | tee $OUT/sweep_stage4_r02.log      # TLAS against the linear model test (500 / 24 models)
timeout 300 python tools/sweep.py --stage 3 2>&1 | tail -20 | tee $OUT/sweep_stage3_r02.log      # config 2
timeout 300 python tools/bvh_build_bench.py 2>&1 | tee $OUT/sweep_r2_bvhbuild.log                # rtBuildBVH against the host builder
