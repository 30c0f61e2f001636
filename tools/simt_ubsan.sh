#!/bin/bash
# Undefined-behaviour check of the kernels and the host side of rt_api.cu without a GPU: the SIMT interpreter build (tests/simt)
# compiled with -fsanitize=undefined,float-cast-overflow (signed overflow, out-of-range float -> int conversions, bad shifts,
# misaligned or null accesses), run through the randomised searches: scenes with non-finite numbers, sessions, BVH builds of
# degenerate and non-finite meshes.  Needs the system g++ (libubsan); ~3 min.   bash tools/simt_ubsan.sh
set -e
cd "$(dirname "$0")/.."
CXX=${UBSAN_CXX:-/usr/bin/g++}
UBSAN_LIB=$($CXX -print-file-name=libubsan.so)
OUT=$(mktemp -d)/librt_b200_simt_ubsan.so
CXX=$CXX python - "$OUT" <<'PY'
import sys
sys.path.insert(0, "tests/simt")
import build as B
B.FLAGS = [f for f in B.FLAGS if f not in ("-O2", "-Wall")] + ["-O1", "-w", "-fsanitize=undefined,float-cast-overflow", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer"]
B.build(force=True, out=sys.argv[1])
PY
export LD_PRELOAD=$UBSAN_LIB UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 RT_SIMT_WORKERS=1
python tools/simt_fuzz.py --cases 150 --seed 61 --odd 0.2 --lib "$OUT" | tail -1
python tools/simt_fuzz_session.py --cases 40 --seed 8 --lib "$OUT" | tail -1
python tools/simt_fuzz_bvh.py --cases 120 --seed 44 --lib "$OUT" | tail -1
echo "no undefined behaviour reported"
