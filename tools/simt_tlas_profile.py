#!/usr/bin/env python
"""Work of the model loop with and without the TLAS, counted WITHOUT a GPU on the SIMT interpreter's profile build
(RT_SIMT_PROFILE: rt_device.cuh counts every per-model box test, every TLAS box test and every TLAS walk).

The reference walks every Model for every ray segment (RayCommon.hlsl:347-371).  Kernels 1 and 2 test the ray against each
model's padded world box instead (option "modelSkip", linear in the model count); with the option "tlas" one walk of a tree over
those boxes per ray segment marks the candidates and only they are tested.  Counts are exact (they do not depend on the
schedule); every configuration renders the same bits as the oracle (asserted).  What the counts cannot tell is time: a TLAS box
test is two 64-byte record halves and two slab tests, a per-model test one 32-byte box.

    python tools/simt_tlas_profile.py [--quick]     -> markdown table on stdout
"""
import argparse
import ctypes as C
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
sys.path.insert(0, os.path.join(REPO, "tests", "simt"))
import build as simt_build                      # noqa: E402
from conftest import ORACLE_LIB, assert_bit_equal, render   # noqa: E402
from ray_tracing_b200 import scenes             # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    lib = simt_build.build(defines=("RT_SIMT_PROFILE",), out=os.path.join(simt_build.OUT_DIR, "librt_b200_simt_profile.so"))
    L = C.CDLL(lib)
    prof = (C.c_ulonglong * 64).in_dll(L, "simtProf")
    knob = (C.c_int * 8).in_dll(L, "simtKnob")
    knob[0], knob[1], knob[2], knob[3], knob[4] = 1, 1, 1, 2, 1          # the default schedule of kernel 2
    counts = [22, 118] if args.quick else [22, 62, 118, 498, 1998]
    print("| models | kernel | rays | linear: model box tests / ray | TLAS: box tests / ray (tree + per-model) | candidates marked / ray | ratio |")
    print("|---:|---|---:|---:|---:|---:|---:|")
    for n in counts:
        sc = scenes.instanced_knots(96, 54, max_bounces=6, rays_per_pixel=2, instances=n)
        ref, _ = render(ORACLE_LIB, sc, frames=1)
        for kernel in (2, 1):
            row = {}
            for tlas in (0, 1):
                for i in range(64):
                    prof[i] = 0
                f, _, st = render(lib, sc, frames=1, options={"kernel": kernel, "tlas": tlas}, want_stats=True)
                assert_bit_equal(f, ref, f"{n} models kernel {kernel} tlas {tlas}")
                row[tlas] = (st["rays"], int(prof[25]), int(prof[26]), int(prof[27]))
            rays, lin, _, _ = row[0]
            _, per, tree, walks = row[1]
            assert walks == rays, (walks, rays)
            print(f"| {n + 2} | {kernel} | {rays} | {lin / rays:.1f} | {(tree + per) / rays:.1f} ({tree / rays:.1f} + {per / rays:.1f}) | {per / rays:.1f} | "
                  f"{lin / max(tree + per, 1):.1f}x |", flush=True)


if __name__ == "__main__":
    main()
