#!/bin/bash
# Eighth GPU call of round 2 (one B200): sample chunks of the one-path-per-lane kernel on rank 0's tile of 8 and 4 (and on the whole image).
OUT=gpurun_out; mkdir -p $OUT
rm -f $OUT/tile_ab.jsonl
timeout 600 python -m pytest tests/test_gpu_round2_abi.py -m gpu -x -q -k "sample_chunks" 2>&1 | tail -3
timeout 900 python tools/tile_ab.py --world 8 4 --workloads cornell64 knot256 2>&1 | tail -40 | tee $OUT/tile_ab_r02j.log
