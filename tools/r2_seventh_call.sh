#!/bin/bash
# Seventh GPU call of round 2 (one B200): occupancy of the one-path-per-lane kernel (6 / 7 / 8 CTAs per SM) on the whole image and on
# rank 0's tile of 8; the final bench line and GPU suite; ncu --set full of the final pooled kernel on the 1M-triangle scene.
OUT=gpurun_out; mkdir -p $OUT
rm -f $OUT/tile_ab.jsonl
for lib in "" "--lib ray_tracing_b200/variants/librt_b200_mb7.so" "--lib ray_tracing_b200/variants/librt_b200_mb8.so"; do
  timeout 600 python tools/tile_ab.py --world 8 --workloads cornell64 cornell1 knot256 --only "kernel 1 (one" $lib 2>&1 | tail -12 | tee -a $OUT/tile_ab_r02i.log
done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $OUT/pytest_gpu_r02i.log
timeout 900 python bench.py 2> $OUT/bench_r02i.err | tail -1 > $OUT/bench_r02i.json; tail -3 $OUT/bench_r02i.err
python -c "
import json; d=json.load(open('$OUT/bench_r02i.json'))
print('main', d['value'], d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'])
for k,v in d['extra'].items(): print(k, v.get('value'), v.get('ms_per_step'), 'frac', v.get('roofline',{}).get('frac'), v.get('roofline',{}).get('bound'), 'e2e', v.get('e2e',{}).get('value'), v.get('error'))
"
bash tools/gpu_prof.sh r02i_soup4k soup4k 2
