#!/bin/bash
# Tenth GPU call of round 2 (one B200): the shade-phase slot fields in global memory (more L1 for the trace phase).
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python tools/sweep.py --stage 11 2>&1 | tail -14 | tee $OUT/sweep_stage11_r02.log
for lib in "" "--lib ray_tracing_b200/variants/librt_b200_cold.so"; do
  timeout 300 python bench.py --workload soup4k --extra none --no-cpu --steps 3 $lib 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); p = d['roofline'].get('physical', {})
print('soup4k', '$lib', d['ms_per_step'], 'L1 hit', p.get('l1_hit_pct'), 'L2', p.get('l2', {}).get('hit_pct'), p.get('l2', {}).get('bytes_per_launch'), 'issue', p.get('issue'))"
done
