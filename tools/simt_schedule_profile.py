#!/usr/bin/env python
"""Schedule profile of the pooled wavefront kernel WITHOUT a GPU: the SIMT interpreter build (tests/simt) compiled with
RT_SIMT_PROFILE counts, per configuration of the scheduling knobs (vote weights, inner / leaf visits per census, tail lanes),
the census iterations, the steps executed per kind and the lanes they ran with, the phases and the shade / camera batches.

Step counts are exact (the schedule of a warp is deterministic); the instruction cost per step kind is an ESTIMATE taken from the
SASS line profile of the round-1 kernel (profiles/r01_f_*): census 20, inner visit 75, leaf triangle 55, next-model 90, shade batch
420, camera batch 160, phase switch 120.  The proxy reproduces the two directions measured on the B200 in round 1 (2 inner visits
per census beat 1 and 3; 16 tail lanes beat 8), which is what it is trusted for: ranking schedules, not predicting milliseconds.
Every configuration renders the same bits (asserted).

    python tools/simt_schedule_profile.py [--quick]     -> markdown table on stdout
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
sys.path.insert(0, os.path.join(REPO, "tests", "simt"))
import build as simt_build                      # noqa: E402
from conftest import ORACLE_LIB, assert_bit_equal, render   # noqa: E402
from ray_tracing_b200 import scenes             # noqa: E402

COST = dict(census=20, inner=75, leaf=55, next=90, shade=420, gen=160, phase=120)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    lib = simt_build.build(force=True, defines=("RT_SIMT_PROFILE",), out=os.path.join(simt_build.OUT_DIR, "librt_b200_simt_profile.so"))
    L = C.CDLL(lib)
    prof = (C.c_ulonglong * 64).in_dll(L, "simtProf")
    knob = (C.c_int * 8).in_dll(L, "simtKnob")
    s = 0.5 if args.quick else 1.0
    cases = [("knot 87k + room, 8 bounces", scenes.knot_room(int(384 * s), int(216 * s), max_bounces=8, rays_per_pixel=2)),
             ("glass knot 87k, 10 bounces", scenes.knot_room(int(320 * s), int(180 * s), max_bounces=10, rays_per_pixel=2, glass=True)),
             ("soup 200k triangles, sky, 16 bounces", scenes.random_soup(int(192 * s), int(192 * s), max_bounces=16, rays_per_pixel=2, triangles=200000, spheres=16))]
    #          vote weights (inner, leaf, next), inner visits per census, leaf triangles per census, tail lanes
    configs = [((1, 1, 1), 2, 1, 16), ((1, 1, 1), 1, 1, 16), ((1, 1, 1), 3, 1, 16), ((1, 1, 1), 2, 1, 8), ((1, 1, 1), 2, 1, 24),
               ((1, 2, 1), 2, 1, 16), ((1, 3, 2), 2, 1, 16), ((1, 4, 2), 2, 1, 16), ((2, 7, 4), 2, 1, 16), ((1, 6, 3), 2, 1, 16),
               ((1, 3, 2), 2, 2, 16), ((1, 2, 2), 2, 2, 16), ((1, 3, 2), 3, 1, 16), ((1, 3, 2), 2, 1, 24)]
    print("| scene | weights i/l/n | inner visits | leaf tris | tail | census iterations | inner steps @ lanes | leaf steps @ lanes | next steps @ lanes | idle lanes | shade batches @ hits | est. instructions | vs first |")
    print("|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
    for name, sc in cases:
        ref, _ = render(ORACLE_LIB, sc, frames=1)
        base = None
        for w, R, LR, tail in configs:
            knob[0], knob[1], knob[2], knob[3], knob[4] = w[0], w[1], w[2], R, LR
            for i in range(64):
                prof[i] = 0
            f, _ = render(lib, sc, frames=1, options={"kernel": 2, "tailLanes": tail})
            assert_bit_equal(f, ref, f"{name} {w} {R} {LR} {tail}")
            p = [int(prof[i]) for i in range(64)]
            cost = (p[0] * COST["census"] + p[7] * COST["inner"] * R + p[9] * COST["leaf"] * LR + p[5] * COST["next"]
                    + p[15] * COST["shade"] + p[17] * COST["gen"] + p[14] * COST["phase"])
            base = base or cost
            d = lambda a, b: a / b if b else 0.0
            print(f"| {name} | {w[0]}/{w[1]}/{w[2]} | {R} | {LR} | {tail} | {p[0]} | {p[7]} @ {d(p[8], p[7]):.1f} | {p[9]} @ {d(p[10], p[9]):.1f} | {p[5]} @ {d(p[6], p[5]):.1f} | "
                  f"{d(p[1], p[0]):.1f} | {p[15]} @ {d(p[16], p[15]):.1f} | {cost / 1e6:.1f} M | {cost / base:.3f} |", flush=True)
    knob[0], knob[1], knob[2], knob[3], knob[4] = 1, 1, 1, 2, 1


if __name__ == "__main__":
    main()
