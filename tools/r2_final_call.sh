#!/bin/bash
# Last GPU call of round 2 (one B200): the whole GPU suite, smoke(), the default bench line and the reference arm, and the ncu launch list
# of the bench command (kernel share of a step).
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $OUT/pytest_gpu_r02z.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -7 | tee $OUT/smoke_r02z.log
SECONDS=0
timeout 900 python bench.py 2> $OUT/bench_r02z.err | tail -1 > $OUT/bench_r02z.json; echo "bench.py took $SECONDS s"; tail -2 $OUT/bench_r02z.err
SECONDS=0
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > $OUT/bench_ref_r02z.json; echo "reference arm took $SECONDS s"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_bench_default_r02.csv python bench.py --steps 2 --warmup 3 --extra none --no-probe --no-cpu 2>&1 | tail -1 | cut -c1-200
python - <<PY
import json
d=json.load(open("$OUT/bench_r02z.json"))
print('main', d['value'], d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'clocks', d['clocks'])
for k,v in d['extra'].items(): print(k, v.get('value'), v.get('ms_per_step'), 'frac', v.get('roofline',{}).get('frac'), v.get('roofline',{}).get('bound'), 'e2e', v.get('e2e',{}).get('value'), v.get('error'))
print(open("$OUT/bench_ref_r02z.json").read()[:300])
PY
