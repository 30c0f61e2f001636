#!/bin/bash
# Ninth GPU call of round 2 (one B200): sample chunks in the pooled kernel on rank 0's tile of 4 and 8; the whole-image kernels must not have moved.
OUT=gpurun_out; mkdir -p $OUT
rm -f $OUT/tile_ab.jsonl
timeout 600 python -m pytest tests/test_gpu_round2_abi.py -m gpu -x -q -k "sample_chunks" 2>&1 | tail -3
timeout 900 python tools/tile_ab.py --world 4 8 --workloads knot256 2>&1 | tail -30 | tee $OUT/tile_ab_r02k.log
timeout 600 python tools/tile_ab.py --world 2 --workloads cornell64 knot64 --only "default" 2>&1 | tail -6 | tee -a $OUT/tile_ab_r02k.log
