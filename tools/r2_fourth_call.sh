#!/bin/bash
# Fourth GPU call of round 2 (one B200): how long the default bench line takes, the BVH build after the copy / early-out changes,
# TMA-staged tree tops on the new kernels (with the L1 / L2 hit rates of the ncu probe for the A/B), the whole GPU suite again.
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $OUT/pytest_gpu_r02e.log
SECONDS=0
timeout 900 python bench.py 2> $OUT/bench_r02e.err | tail -1 > $OUT/bench_r02e.json
echo "python bench.py took $SECONDS s" | tee $OUT/bench_r02e.time
tail -3 $OUT/bench_r02e.err
RT_B200_BVH_TIMING=1 timeout 400 python tools/bvh_build_bench.py 2>&1 | grep -v "^\[rtBuildBVH\] 87120" | tee $OUT/bvh_build_r02e.log
timeout 600 python tools/sweep.py --stage 8 2>&1 | tail -14 | tee $OUT/sweep_stage8_r02.log
for sn in 0 256; do
  timeout 300 python bench.py --workload knot64 --extra none --no-cpu --smem-nodes $sn 2>/dev/null | tail -1 > $OUT/probe_knot64_smem$sn.json
  timeout 300 python bench.py --workload soup4k --extra none --no-cpu --smem-nodes $sn --steps 3 2>/dev/null | tail -1 > $OUT/probe_soup4k_smem$sn.json
done
python - <<PY
import json
for wl in ("knot64", "soup4k"):
    for sn in (0, 256):
        d = json.load(open("$OUT/probe_%s_smem%d.json" % (wl, sn))); p = d["roofline"].get("physical", {})
        print(wl, "smemNodes", sn, d["ms_per_step"], "ms  L1 hit", p.get("l1_hit_pct"), "L2 hit", p.get("l2", {}).get("hit_pct"), "L2 bytes", p.get("l2", {}).get("bytes_per_launch"), "issue", p.get("issue"))
PY
