"""TEST INFRASTRUCTURE ONLY.  Builds tests/simt/_build/librt_b200_simt.so: the product's own rt_api.cu (all kernels included)
compiled by g++ against the SIMT interpreter in tests/simt/shim/ — see shim/cuda_runtime.h for what it is and is not.
The product (ray_tracing_b200/, bench.py, __graft_entry__.py) never builds or loads this library."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(REPO, "ray_tracing_b200", "csrc")
SHIM = os.path.join(HERE, "shim")
OUT_DIR = os.path.join(HERE, "_build")
LIB_SIMT = os.path.join(OUT_DIR, "librt_b200_simt.so")

# the same arithmetic contract as the nvcc build (-fmad=false ...): one IEEE binary32 operation per operator, no contraction
FLAGS = ["-O2", "-g1", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-pthread", "-shared", "-Wl,-Bsymbolic",
         "-Wall", "-Wno-unknown-pragmas", "-Wno-unused-function", "-Wno-unused-variable", "-DRT_SIMT_EMU"]


def _sources() -> list[str]:
    out = []
    for d in (CSRC, SHIM, os.path.join(REPO, "include")):
        out += [os.path.join(d, f) for f in os.listdir(d)]
    return out


def build(force: bool = False, defines: tuple = (), out: str = "") -> str:
    """RT_SIMT_DEFINES=A,B=2 in the environment builds (and makes the test suite use) a variant with those macros defined."""
    env_defs = tuple(d for d in os.environ.get("RT_SIMT_DEFINES", "").split(",") if d)
    if env_defs and not defines and not out:
        defines = env_defs
        out = os.path.join(OUT_DIR, "librt_b200_simt_" + "_".join(d.replace("=", "") for d in env_defs) + ".so")
    target = out or LIB_SIMT
    os.makedirs(os.path.dirname(target), exist_ok=True)
    if not force and os.path.exists(target) and all(os.path.getmtime(s) <= os.path.getmtime(target) for s in _sources()):
        return target
    cmd = [os.environ.get("CXX", "g++")] + FLAGS + [f"-D{d}" for d in defines] + \
          ["-I", SHIM, "-I", os.path.join(REPO, "include"), "-x", "c++", os.path.join(CSRC, "rt_api.cu"), "-o", target]
    print("[simt build]", " ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("SIMT test build failed")
    return target


if __name__ == "__main__":
    print(build(force=True))
