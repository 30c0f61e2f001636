#!/bin/bash
# Sixth GPU call of round 2 (one B200): one GPU's share of a 2 / 4 / 8-GPU run per kernel (threshold of the automatic kernel choice),
# the GPU suite on the promoted defaults (predicated push, one-sphere leaves).
OUT=gpurun_out; mkdir -p $OUT
rm -f $OUT/tile_ab.jsonl
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $OUT/pytest_gpu_r02h.log
timeout 900 python tools/tile_ab.py --world 2 4 8 --workloads knot256 knot64 knot1 2>&1 | tail -60 | tee $OUT/tile_ab_r02h.log
