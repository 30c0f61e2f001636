#!/bin/bash
# Second GPU call of round 2: the whole GPU suite with the round-2 ABI tests, the new bench.py (mesh default + extras + ncu probe),
# combinations of the first call's winners, a launch list of one device BVH build.
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/r2_second_call.sh'
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_gpu_r02b.log
timeout 900 python bench.py 2> $OUT/bench_r02b.err | tail -1 | tee $OUT/bench_r02b.json
tail -5 $OUT/bench_r02b.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee $OUT/bench_ref_r02b.json
timeout 900 python tools/sweep.py --stage 5 2>&1 | tail -50 | tee $OUT/sweep_stage5_r02.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/bvh_build_launches_r02.csv python tools/bvh_build_bench.py --only soup --gpu-only --repeat 1 2>&1 | tail -2
