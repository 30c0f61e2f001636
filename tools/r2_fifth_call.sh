#!/bin/bash
# Fifth GPU call of round 2 (one B200): one GPU's share of an 8-GPU run (rank 0's tile) under the tail fixes, sphere-tree shapes,
# the predicated push.
OUT=gpurun_out; mkdir -p $OUT
rm -f $OUT/tile_ab.jsonl
timeout 900 python tools/tile_ab.py --world 8 --workloads knot256 cornell64 knot64 2>&1 | tail -30 | tee $OUT/tile_ab_r02.log
timeout 600 python tools/sweep.py --stage 10 2>&1 | tail -8 | tee $OUT/sweep_stage10_r02.log
timeout 600 python tools/sweep.py --stage 9 2>&1 | tail -8 | tee $OUT/sweep_stage9_r02.log
