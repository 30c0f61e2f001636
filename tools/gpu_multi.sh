#!/bin/bash
# usage: gpu_multi.sh N tag   (run under gpurun --gpus N)
N=$1; TAG=${2:-mg}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L | tee $OUT/gpus_$TAG.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/multigpu_check.py 2>&1 | grep -v "^W\|^\*\*\*" | tail -40 | tee $OUT/multigpu_check_$TAG.log
for wl in ${WORKLOADS:-cornell64 knot64}; do
for ex in allgather fused; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps ${STEPS:-5} --warmup 3 --workload $wl --exchange $ex 2>&1 | grep '^{\|Error\|error' | tail -2 | tee $OUT/bench_${wl}_n${N}_${ex}_$TAG.log
done
done
if [ "${SKIP_N1:-0}" = "0" ]; then
for wl in ${WORKLOADS:-cornell64 knot64}; do
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --workload $wl 2>&1 | tail -1 | tee $OUT/bench_${wl}_n1_$TAG.log
done
fi
