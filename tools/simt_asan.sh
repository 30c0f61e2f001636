#!/bin/bash
# Memory check of the kernels without a GPU: the SIMT interpreter build (tests/simt) compiled with AddressSanitizer — device buffers
# are heap blocks with red zones, the dynamic shared memory is a static array with red zones — runs all three trace kernels (plain /
# instrumented / EXT / TLAS instantiations, staged tree tops, treelet order, sphere accelerator) and the device BVH build against the oracle.
# Needs the system g++ (libasan); ~2 min.   bash tools/simt_asan.sh
set -e
cd "$(dirname "$0")/.."
CXX=${ASAN_CXX:-/usr/bin/g++}
ASAN_LIB=$($CXX -print-file-name=libasan.so)
OUT=$(mktemp -d)/librt_b200_simt_asan.so
CXX=$CXX python - "$OUT" <<'PY'
import sys
sys.path.insert(0, "tests/simt")
import build as B
B.FLAGS = [f for f in B.FLAGS if f not in ("-O2", "-Wall")] + ["-O1", "-w", "-fsanitize=address", "-fno-omit-frame-pointer"]
B.build(force=True, out=sys.argv[1])
PY
LD_PRELOAD=$ASAN_LIB ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0:halt_on_error=1 RT_SIMT_WORKERS=1 python - "$OUT" <<'PY'
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from conftest import render, assert_bit_equal, ORACLE_LIB
import ray_tracing_b200 as rt
from ray_tracing_b200 import scenes, capi
LIB = sys.argv[1]
for sc, frames in ((scenes.cornell_spheres(48, 32, 4, 2), 2), (scenes.knot_room(48, 28, 4, 2, nu=60, nv=8, glass=True), 2),
                   (scenes.random_soup(40, 40, max_bounces=5, rays_per_pixel=2, triangles=6000, spheres=300), 1)):
    fo, ao = render(ORACLE_LIB, sc, frames=frames)
    for opts in ({"kernel": 0}, {"kernel": 1}, {"kernel": 2}, {"kernel": 2, "smemNodes": 200, "pairOrder": 2, "poolSlots": 96, "gridFit": 1},
                 {"kernel": 1, "smemNodes": 50}, {"kernel": 2, "countStats": 1, "extInstantiation": 1}):
        fg, ag = render(LIB, sc, frames=frames, options=opts)
        assert_bit_equal(ag, ao, f"{sc.name} {opts}")
    print(sc.name, "clean", flush=True)
# TLAS kernels: 70 models (automatic), 12 models (forced, two-level tree), 3 models (forced, the root is a leaf)
for sc, tlas in ((scenes.instanced_knots(48, 28, 4, 2, instances=68), -1), (scenes.instanced_knots(48, 28, 4, 2, instances=10), 1),
                 (scenes.knot_room(48, 28, 4, 2, nu=60, nv=8), 1)):
    fo, ao = render(ORACLE_LIB, sc, frames=2)
    for kernel in (1, 2):
        fg, ag = render(LIB, sc, frames=2, options={"kernel": kernel, "tlas": tlas, "poolSlots": 32 if kernel == 2 else 64})
        assert_bit_equal(ag, ao, f"{sc.name} tlas {tlas} kernel {kernel}")
print("TLAS kernels clean", flush=True)
gpu = capi.RtLib(LIB).create(0)
m = scenes.knot_mesh(nu=100, nv=8)
for q in (1, 0, 2):
    th, nh, _ = rt.build_bvh(m.vertices, m.indices, m.normals, q)
    tg, ng = gpu.build_bvh(m.vertices, m.indices, m.normals, q)
    assert np.array_equal(ng.view(np.uint8), nh.view(np.uint8)) and np.array_equal(tg.view(np.uint8), th.view(np.uint8))
print("device BVH build clean")
# round 2: sample chunks (hand-off buffers), the group context (rtCreateMulti: tile staging + the in-process exchange), pipelined readback,
# deep-tree refusal, a rank without rows
import test_round2_abi as R
R.sample_chunks_are_schedule_only(LIB, sizes=((33, 7),))
R.group_equals_single(LIB, [0, 1, 2], scenes.knot_room(40, 30, 3, 2, nu=30, nv=6, glass=True))
R.deep_trees(LIB)
R.model_count_alone(LIB)
R.root_bounds_are_never_read(LIB)
sc = scenes.knot_room(7, 1, 3, 2, nu=20, nv=6)
for rank in (0, 1, 2):
    render(LIB, sc, frames=1, options={"kernel": 1, "sampleChunks": 7}, tile=(rank, 3, 1))
print("round-2 ABI paths clean")
PY
