#!/bin/bash
# Multi-GPU sessions of round 2 (N GPUs of one box).
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 900  -- 'bash tools/r2_multi_call.sh 2'     tests with the real NCCL all-gather inside rtDispatch + the bench line at N = 2
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 1200 -- 'bash tools/r2_multi_call.sh 8'     the bench line at N = 8 (extras: the 1M-triangle scene) + A/Bs of the tail fixes
N=${1:-2}
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi --query-gpu=name --format=csv,noheader | sort | uniq -c > $OUT/gpus_n$N.txt
if [ $N -le 2 ]; then
  timeout 600 python -m pytest tests/test_gpu_round2_abi.py -m gpu -x -q -k "group_context or cpp_tiled" 2>&1 | tail -6 | tee $OUT/pytest_multigpu_r02.log
fi
summ() { python -c "
import sys,json
d=json.loads(open('$1').read().strip().splitlines()[-1])
print('N=%d %s: %.1f Mrays/s %.3f ms/frame kernel %.3f exch %.3f e2e %.1f (%.3f ms)' % (d['n_gpus'], d['config']['name'], d['value'], d['ms_per_step'], d['kernel_ms_per_launch_max_over_ranks'], d['exchange_ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']))
for k,v in d.get('extra',{}).items():
    print('   ', k, v.get('error') or '%.1f Mrays/s %.3f ms kernel %.3f exch %.3f e2e %.1f' % (v['value'], v['ms_per_step'], v['kernel_ms_per_launch_max_over_ranks'], v['exchange_ms_per_step'], v['e2e']['value']))
"; }
PORT=29510
SECONDS=0
NCCL_DEBUG=WARN timeout 870 $TR --nproc-per-node $N --master-port $PORT bench.py --gpus $N 2> $OUT/scale_n$N.err | tail -1 > $OUT/scale_r02_n$N.json
echo "bench --gpus $N took $SECONDS s" | tee $OUT/scale_n$N.time
summ $OUT/scale_r02_n$N.json 2>&1 | tee $OUT/scale_r02_n$N.txt || tail -20 $OUT/scale_n$N.err
if [ $N -ge 8 ]; then
  for cfg in "--workload cornell64" "--workload cornell64 --kernel 2" "--workload cornell64 --kernel 2 --pool-slots 64" "--workload knot256 --pool-slots 64" "--workload knot256 --exchange torch"; do
    PORT=$((PORT+1))
    echo "== $cfg" | tee -a $OUT/tail_ab_r02.log
    timeout 300 $TR --nproc-per-node $N --master-port $PORT bench.py --gpus $N --no-probe --extra none $cfg 2>> $OUT/tail_ab.err | tail -1 > $OUT/tail_one.json
    cat $OUT/tail_one.json >> $OUT/tail_ab_r02.jsonl
    summ $OUT/tail_one.json 2>&1 | tee -a $OUT/tail_ab_r02.log
  done
fi
