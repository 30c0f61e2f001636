#!/bin/bash
# SUPERSEDED for quick ranking by tools/build_variants.py (run locally) + tools/sweep.py --stage 1|2|3 (one process per workload, no
# compiling on the GPU box); this script is the long form that goes through bench.py for every configuration.
# A/B candidates prepared at the end of round 1 (compile-time variants, all off by default, none measured yet; the pool-kernel
# ones are bit-exact on the SIMT interpreter build: RT_SIMT_VARIANTS=1 python -m pytest tests/test_simt_kernels.py -k variants).
# Build them next to the default library and compare on one GPU:   bash tools/round2_sweep.sh   (under gpurun)
python - <<'PY'
from ray_tracing_b200 import build
import os
P = build.PKG_DIR
for name, defs in [("skipsqrt", ("RT_SPHERE_SKIP_SQRT",)), ("mb5", ("RT_WAVE_MINBLOCKS=5",)), ("ir1", ("RT_INNER_REPEAT=1",)), ("pw20", ("RT_POOL_WARPS=20",)),
                   ("stacktop", ("RT_STACK_TOP_REG",)), ("rayinv", ("RT_CACHE_RAYINV",)), ("leaf2", ("RT_LEAF_REPEAT=2",)),
                   ("stacktop_leaf2", ("RT_STACK_TOP_REG", "RT_LEAF_REPEAT=2")), ("tri_na", ("RT_TRI_LOAD_POLICY=1",)), ("tri_ef", ("RT_TRI_LOAD_POLICY=2",)), ("pf", ("RT_PREFETCH_NEXT_PAIR",)), ("treelet", ("RT_TREELET_PREFETCH",))]:
    build.build_cuda(force=True, defines=defs, out=os.path.join(P, f"librt_b200_{name}.so"))
PY
RT_B200_LIB=ray_tracing_b200/librt_b200_skipsqrt.so python -m pytest tests -m gpu -q -k "cornell or sphere or soup" 2>&1 | tail -3
bash tools/gpu_ab.sh r2 librt_b200.so librt_b200_skipsqrt.so librt_b200_mb5.so
bash tools/gpu_ab2.sh r2 librt_b200.so librt_b200_ir1.so librt_b200_pw20.so librt_b200_stacktop.so librt_b200_rayinv.so librt_b200_leaf2.so librt_b200_stacktop_leaf2.so librt_b200_tri_na.so librt_b200_tri_ef.so
# node-pair record order (host-side layout only; kernel unchanged): breadth-first (0, default) vs treelets laid out depth-first
OUT=gpurun_out; mkdir -p $OUT
for wl in knot64 cluster4k soup4k; do for po in 0 1 2 3 4 6; do
  echo "== $wl --pair-order $po" | tee -a $OUT/sweep_r2_pairorder.log
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --workload $wl --pair-order $po 2>&1 | tail -1 | python -c "
import sys,json
l=sys.stdin.read().strip()
try:
    d=json.loads(l); print(d['value'],'Mrays/s', d['ms_per_step'],'ms', 'frac',d['roofline']['frac'])
except Exception as e: print('ERR',l[-600:])" | tee -a $OUT/sweep_r2_pairorder.log
done; done
# gridFit (whole pixels per persistent lane): matters when a GPU has few pixels per lane, i.e. on multi-GPU tiles.  Single GPU
# check of the same regime: a quarter-size frame.  On N GPUs:  torchrun ... bench.py --gpus N --grid-fit 1  (tools/gpu_multi.sh).
for gf in 0 1; do
  echo "== cornell64 --grid-fit $gf" | tee -a $OUT/sweep_r2_gridfit.log
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --workload cornell64 --grid-fit $gf 2>&1 | tail -1 | cut -c1-200 | tee -a $OUT/sweep_r2_gridfit.log
done
# BVH builder row: host (1 thread / all threads) vs rtBuildBVH, identical buffers required
python tools/bvh_build_bench.py 2>&1 | tee -a $OUT/sweep_r2_bvhbuild.log
# persisting L2 window over the node-pair records
for wl in knot64 cluster4k soup4k; do for lp in 0 1; do
  echo "== $wl --l2-persist $lp" | tee -a $OUT/sweep_r2_l2persist.log
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --workload $wl --l2-persist $lp 2>&1 | tail -1 | cut -c1-160 | tee -a $OUT/sweep_r2_l2persist.log
done; done
# next-record prefetch needs the pre-order layout
for wl in knot64 cluster4k soup4k; do
  echo "== $wl prefetch-next-pair + --pair-order 1" | tee -a $OUT/sweep_r2_pairorder.log
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --workload $wl --pair-order 1 --lib ray_tracing_b200/librt_b200_pf.so 2>&1 | tail -1 | cut -c1-160 | tee -a $OUT/sweep_r2_pairorder.log
done
for wl in knot64 cluster4k soup4k; do
  echo "== $wl treelet prefetch" | tee -a $OUT/sweep_r2_pairorder.log
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --workload $wl --treelet-prefetch 1 --lib ray_tracing_b200/librt_b200_treelet.so 2>&1 | tail -1 | cut -c1-160 | tee -a $OUT/sweep_r2_pairorder.log
done
