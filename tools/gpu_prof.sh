#!/bin/bash
# ncu --set full capture of one launch of the dominant kernel.  usage: gpu_prof.sh tag workload kernel [extra bench args]
TAG=$1; WL=$2; K=$3; shift 3
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_raytrace -s 3 -c 1 -f -o $OUT/prof_${TAG} \
    python bench.py --steps 1 --warmup 3 --workload $WL --kernel $K --no-cpu --no-probe --extra none "$@" > $OUT/prof_${TAG}.log 2>&1
tail -2 $OUT/prof_${TAG}.log
