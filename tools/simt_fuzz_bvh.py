#!/usr/bin/env python
"""Randomised comparison of the device BVH build (rtBuildBVH, csrc/rt_bvh_build.cuh, on the SIMT interpreter build) with the host
builder (host/BVH.cpp, the reference's BVH.cs): Nodes and Triangles must be the same bytes.  Random meshes with the things a level-
synchronous build can get wrong: duplicated and coincident triangles, degenerate ones, exact zeros of both signs, grids (many equal
centres -> equal split costs), tiny and huge coordinates, 1..5 triangles, sorted / reversed / shuffled input orders, all three
quality modes.

    python tools/simt_fuzz_bvh.py [--cases 300] [--seed 1]
"""
import argparse
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests", "simt"))
import build as simt_build                      # noqa: E402
import ray_tracing_b200 as rt                   # noqa: E402
from ray_tracing_b200 import capi               # noqa: E402


def random_mesh(rng):
    n = int(rng.choice([1, 2, 3, 5, 17, 64, 300, 2000, 9000]))
    kind = int(rng.randint(0, 7))
    scale = float(rng.choice([1e-4, 1.0, 1.0, 1e4]))
    if kind == 0:                                   # soup
        c = rng.uniform(-1, 1, (n, 1, 3)); tri = c + rng.uniform(-0.05, 0.05, (n, 3, 3))
    elif kind == 1:                                 # regular grid of identical quads' halves: many equal centres and costs
        k = int(np.ceil(np.sqrt(n))); gx, gy = np.meshgrid(np.arange(k), np.arange(k)); g = np.stack([gx.ravel(), gy.ravel(), np.zeros(k * k)], 1)[:n]
        tri = g[:, None, :] + np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], dtype=np.float64)[None]
    elif kind == 2:                                 # every triangle the same
        tri = np.repeat(rng.uniform(-1, 1, (1, 3, 3)), n, axis=0)
    elif kind == 3:                                 # degenerate: points and segments mixed in
        c = rng.uniform(-1, 1, (n, 1, 3)); tri = c + rng.uniform(-0.05, 0.05, (n, 3, 3)); tri[::3, 1] = tri[::3, 0]; tri[::5, 2] = tri[::5, 0]
    elif kind == 4:                                 # exact zeros of both signs, axis-aligned planes
        tri = rng.choice([-1.0, -0.0, 0.0, 1.0, 0.5], (n, 3, 3))
    elif kind == 5:                                 # a line of triangles along one axis (two flat axes)
        tri = np.zeros((n, 3, 3)); tri[:, :, 0] = np.arange(n)[:, None] + rng.uniform(0, 0.9, (n, 3))
    else:                                           # clusters far apart
        c = rng.choice([-100.0, 0.0, 100.0], (n, 1, 3)) + rng.uniform(-1, 1, (n, 1, 3)); tri = c + rng.uniform(-0.2, 0.2, (n, 3, 3))
    tri = tri * scale
    odd = rng.rand()
    if odd < 0.06:                                  # non-finite vertices: the reference builds *something*; the device build must build the same
        for _ in range(int(rng.randint(1, 4))):
            tri[int(rng.randint(0, n)), int(rng.randint(0, 3)), int(rng.randint(0, 3))] = rng.choice([np.nan, np.inf, -np.inf])
    elif odd < 0.12:                                # extents whose surface area overflows FP32: every candidate costs inf, the reference
        tri = tri * (1e20 / scale)                  # then splits by "axis 0, position 0" (ChooseSplit's defaults)
    order = int(rng.randint(0, 4))
    if order == 1:
        tri = tri[::-1]
    elif order == 2:
        tri = tri[rng.permutation(n)]
    elif order == 3:
        tri = tri[np.argsort(tri[:, :, int(rng.randint(0, 3))].mean(1), kind="stable")]
    v = np.ascontiguousarray(tri.reshape(-1, 3), dtype=np.float32)
    nrm = np.ascontiguousarray(rng.uniform(-1, 1, v.shape), dtype=np.float32)
    return v, np.arange(3 * n, dtype=np.int32), nrm


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=300)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--lib", default=None)
    args = ap.parse_args()
    gpu = capi.RtLib(args.lib or simt_build.build()).create(0)
    bad = 0
    for case in range(args.cases):
        seed = args.seed * 100003 + case
        rng = np.random.RandomState(seed)
        v, idx, nrm = random_mesh(rng)
        q = int(rng.choice([1, 1, 0, 2]))
        try:
            th, nh, _ = rt.build_bvh(v, idx, nrm, q)
        except ValueError:                          # more than 2n + 1 nodes (chains of empty children): both sides must refuse
            try:
                gpu.build_bvh(v, idx, nrm, q)
                bad += 1
                print(f"case seed {seed}: the host refuses (node capacity), the device build does not", flush=True)
            except capi.RtError:
                pass
            continue
        tg, ng = gpu.build_bvh(v, idx, nrm, q)
        same = len(ng) == len(nh) and np.array_equal(ng.view(np.uint8), nh.view(np.uint8)) and np.array_equal(tg.view(np.uint8), th.view(np.uint8))
        if not same:
            bad += 1
            where = "node count" if len(ng) != len(nh) else ("nodes" if not np.array_equal(ng.view(np.uint8), nh.view(np.uint8)) else "triangle order")
            print(f"case seed {seed}: {len(idx) // 3} triangles, quality {q}: {where} differ ({len(nh)} vs {len(ng)} nodes)", flush=True)
        if (case + 1) % 50 == 0:
            print(f"  {case + 1} cases, {bad} findings", flush=True)
    gpu.destroy()
    print(f"{args.cases} cases, {bad} findings")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
