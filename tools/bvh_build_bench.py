#!/usr/bin/env python
"""Build-time comparison for the BVH builder row: host builder with 1 thread (the reference's recursion), host
builder with all threads, rtBuildBVH on the GPU.  Prints one JSON line per mesh; the three builds must return identical buffers.

    python tools/bvh_build_bench.py            (needs a GPU for the third column; --lib to point at another build)
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ray_tracing_b200 as rt                      # noqa: E402
from ray_tracing_b200 import build as b, capi, scenes   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=b.LIB_CUDA)
    ap.add_argument("--repeat", type=int, default=3)
    ap.add_argument("--only", default=None, help="substring of the mesh name: run that mesh only (for an ncu launch list of one build)")
    ap.add_argument("--gpu-only", action="store_true", help="skip the host builds (identical = null)")
    args = ap.parse_args()
    meshes = [("knot 87k", scenes.knot_mesh()),
              ("cluster 871k", max(scenes.knot_cluster(64, 36, 2, 1).meshes, key=lambda m: m.triangle_count)),
              ("soup 333k (one of the three models of config 5)", max(scenes.random_soup(16, 16, 2, 1, triangles=1_000_000, spheres=1).meshes, key=lambda m: m.triangle_count))]
    gpu = capi.RtLib(args.lib).create(0)
    if args.only:
        meshes = [(n, m) for n, m in meshes if args.only in n]
    for name, m in meshes:
        if args.gpu_only:
            row = {"mesh": name, "triangles": int(m.triangle_count)}
            for q, qn in ((1, "High"),):
                best = 1e9
                for _ in range(args.repeat):
                    t0 = time.perf_counter(); tg, ng = gpu.build_bvh(m.vertices, m.indices, m.normals, q); best = min(best, time.perf_counter() - t0)
                row[qn] = {"nodes": int(len(ng)), "gpu_ms_incl_copies": round(1e3 * best, 1)}
            print(json.dumps(row), flush=True)
            continue
        row = {"mesh": name, "triangles": int(m.triangle_count)}
        for q, qn in ((1, "High"), (0, "Low")):
            rt.set_build_threads(1)
            t0 = time.perf_counter(); th, nh, _ = rt.build_bvh(m.vertices, m.indices, m.normals, q); t_host1 = time.perf_counter() - t0
            rt.set_build_threads(0)
            t0 = time.perf_counter(); ta, na, _ = rt.build_bvh(m.vertices, m.indices, m.normals, q); t_hostn = time.perf_counter() - t0
            best = 1e9
            for _ in range(args.repeat):
                t0 = time.perf_counter(); tg, ng = gpu.build_bvh(m.vertices, m.indices, m.normals, q); best = min(best, time.perf_counter() - t0)
            same = (np.array_equal(ng.view(np.uint8), nh.view(np.uint8)) and np.array_equal(tg.view(np.uint8), th.view(np.uint8))
                    and np.array_equal(na.view(np.uint8), nh.view(np.uint8)) and np.array_equal(ta.view(np.uint8), th.view(np.uint8)))
            row[qn] = {"nodes": int(len(nh)), "host_1_thread_ms": round(1e3 * t_host1, 1), "host_all_threads_ms": round(1e3 * t_hostn, 1),
                       "gpu_ms_incl_copies": round(1e3 * best, 1), "identical": bool(same)}
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
