#!/bin/bash
# Short 8-GPU confirmation of the default workload after the sample chunks (one bench line, no extras, no probe).
N=${1:-8}
OUT=gpurun_out; mkdir -p $OUT
python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N --master-port 29701 bench.py --gpus $N --no-probe --extra knot64 2> $OUT/confirm_n$N.err | tail -1 > $OUT/confirm_r02_n$N.json
python -c "
import json
d=json.load(open('$OUT/confirm_r02_n$N.json'))
print('N=%d %s: %.1f Mrays/s %.3f ms/frame kernel %.3f exch %.3f e2e %.1f' % (d['n_gpus'], d['config']['name'], d['value'], d['ms_per_step'], d['kernel_ms_per_launch_max_over_ranks'], d['exchange_ms_per_step'], d['e2e']['value']))
for k,v in d['extra'].items(): print('   ', k, v.get('value'), v.get('ms_per_step'))
" || tail -20 $OUT/confirm_n$N.err
