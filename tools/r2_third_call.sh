#!/bin/bash
# Third GPU call of round 2: the promoted default kernels (rt_devmath.cuh "build defaults of round 2") through the whole GPU suite and
# bench.py, variations on top of them, the TLAS threshold, the BVH build with its time split, ncu --set full of the two mesh workloads.
#   /usr/local/graft/bin/gpurun --timeout 2700 -- 'bash tools/r2_third_call.sh'
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $OUT/pytest_gpu_r02c.log
timeout 900 python bench.py 2> $OUT/bench_r02c.err | tail -1 > $OUT/bench_r02c.json; tail -3 $OUT/bench_r02c.err
python -c "
import json; d=json.load(open('$OUT/bench_r02c.json'))
print('main', d['value'], d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'])
for k,v in d['extra'].items(): print(k, v.get('value'), v.get('ms_per_step'), 'frac', v.get('roofline',{}).get('frac'), v.get('roofline',{}).get('bound'), 'e2e', v.get('e2e',{}).get('value'), v.get('error'))
"
timeout 900 python tools/sweep.py --stage 7 2>&1 | tail -45 | tee $OUT/sweep_stage7_r02.log
timeout 400 python tools/sweep.py --stage 6 2>&1 | tail -8 | tee $OUT/sweep_stage6_r02.log
RT_B200_BVH_TIMING=1 timeout 400 python tools/bvh_build_bench.py 2>&1 | tee $OUT/bvh_build_r02c.log
bash tools/gpu_prof.sh r02c_soup4k soup4k 2
bash tools/gpu_prof.sh r02c_knot256 knot256 2
