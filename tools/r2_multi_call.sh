#!/bin/bash
# Multi-GPU session of round 2 (N GPUs of one box, N = first argument, default 8): the multi-GPU ABI tests with the real NCCL
# all-gather inside rtDispatch, the scaling curve of the default bench line (with its extras: soup4k = the 1M-triangle scene), and the
# A/Bs of the tail fixes.      /usr/local/graft/bin/gpurun --gpus 8 --timeout 1500 -- 'bash tools/r2_multi_call.sh 8'
N=${1:-8}
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_round2_abi.py -m gpu -x -q -k "group_context or cpp_tiled" 2>&1 | tail -6 | tee $OUT/pytest_multigpu_r02.log
PORT=29500
for n in 1 2 4 8; do
  if [ $n -le $N ]; then
    PORT=$((PORT+1))
    if [ $n -eq 1 ]; then timeout 900 python bench.py --gpus 1 --no-cpu 2> $OUT/scale_n1.err | tail -1 > $OUT/scale_r02_n1.json
    else NCCL_DEBUG=WARN timeout 900 $TR --nproc-per-node $n --master-port $PORT bench.py --gpus $n 2> $OUT/scale_n$n.err | tail -1 > $OUT/scale_r02_n$n.json; fi
    python - <<PY
import json
try:
    d = json.load(open("$OUT/scale_r02_n$n.json"))
    print("N=$n", d["config"]["name"], d["value"], "Mrays/s", d["ms_per_step"], "ms  e2e", d["e2e"]["value"], " extras:", {k: (v.get("value"), v.get("e2e", {}).get("value")) for k, v in d.get("extra", {}).items()})
except Exception as e:
    print("N=$n failed", e); print(open("$OUT/scale_n$n.err").read()[-1500:])
PY
  fi
done
# tail fixes at N GPUs: 96 pool slots (all pixels of a 1080p tile resident at once), pooled kernel for the sphere scene, grid fit
if [ $N -ge 8 ]; then
  for cfg in "--workload cornell64 --extra none" "--workload cornell64 --extra none --kernel 2" "--workload cornell64 --extra none --kernel 2 --pool-slots 64" "--workload cornell64 --extra none --grid-fit 1" "--workload knot256 --extra none --pool-slots 64" "--workload knot256 --extra none --exchange torch" "--workload soup4k16 --extra none"; do
    PORT=$((PORT+1))
    echo "== $cfg" | tee -a $OUT/tail_ab_r02.log
    timeout 600 $TR --nproc-per-node $N --master-port $PORT bench.py --gpus $N --no-probe $cfg 2>> $OUT/tail_ab.err | tail -1 | tee -a $OUT/tail_ab_r02.jsonl | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], 'kernel', d['kernel_ms_per_launch_max_over_ranks'], 'exch', d['exchange_ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])" | tee -a $OUT/tail_ab_r02.log
  done
fi
