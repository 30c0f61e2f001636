#!/bin/bash
# 8-GPU A/B of the tail fixes after the fair-share cap (tools/r2_multi_call.sh has the full line).  gpurun --gpus 8 -- 'bash tools/r2_multi_ab.sh 8'
N=${1:-8}
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
PORT=29600
rm -f $OUT/tail_ab2_r02.jsonl
for cfg in "--workload knot256" "--workload knot256 --pool-slots 64" "--workload cornell64" "--workload cornell64 --kernel 2" "--workload knot64" "--workload cornell1" "--workload cornell1 --kernel 2"; do
  PORT=$((PORT+1))
  echo "== $cfg" | tee -a $OUT/tail_ab2_r02.log
  timeout 300 $TR --nproc-per-node $N --master-port $PORT bench.py --gpus $N --no-probe --extra none $cfg 2>> $OUT/tail_ab2.err | tail -1 > $OUT/tail_one.json
  cat $OUT/tail_one.json >> $OUT/tail_ab2_r02.jsonl
  python -c "
import json
d=json.loads(open('$OUT/tail_one.json').read().strip().splitlines()[-1])
print('N=%d %s: %.1f Mrays/s %.3f ms/frame kernel %.3f exch %.3f e2e %.1f (%.3f ms)' % (d['n_gpus'], d['config']['name'], d['value'], d['ms_per_step'], d['kernel_ms_per_launch_max_over_ranks'], d['exchange_ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']))" 2>&1 | tee -a $OUT/tail_ab2_r02.log
done
