#!/usr/bin/env python
"""A/B sweep of kernel variants and options in ONE process per workload (the scene is generated once; every configuration gets a
fresh context on it).  Kernel time = CUDA events around the dispatches (rtGetStats kernelMs), after warm-up; a 256 MiB write
between frames evicts L2.  For ranking candidates; the numbers that count are bench.py's.

    python tools/build_variants.py                      (here, before gpurun)
    python tools/sweep.py --stage 1                     (on the GPU box; writes gpurun_out/sweep_r2.jsonl)
"""
import argparse
import hashlib
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tools"))
import bench                                         # noqa: E402  (workload table, scene factory, algorithmic bytes)
from build_variants import path as variant           # noqa: E402

# (label, variant library or None, options)
STAGES = {
    1: (["knot64", "soup4k"], [
        ("default", None, {}),
        ("256-bit pair loads", "ldg256", {}),
        ("256-bit pair loads + 64-byte triangles", "ldg256_tri64", {}),
        ("256-bit pair loads + 64-byte triangles, no_allocate", "ldg256_tri64_na", {}),
        ("256-bit pair loads + vote 1/3/2", "ldg256_vote132", {}),
        ("256-bit loads + 64-byte triangles + vote 1/3/2 + prefetch cur", "ldg256_tri64_vote132_pfcur", {}),
        ("vote 1/3/2", "vote132", {}),
        ("vote 1/4/2", "vote142", {}),
        ("vote 2/7/4", "vote274", {}),
        ("vote 1/6/3", "vote163", {}),
        ("vote 1/3/2 + prefetch cur", "vote132_pfcur", {}),
        ("vote 1/3/2 + prefetch cur + treelet", "vote132_pfcur_treelet", {"treeletPrefetch": 1}),
        ("treelet", "treelet", {"treeletPrefetch": 1}),
        ("treelet+stacktop", "treelet_stacktop", {"treeletPrefetch": 1}),
        ("stacktop", "stacktop", {}),
        ("smem stack 4", "smemstack4", {}),
        ("c3 + branch-free pop from the shared-memory stack", "c8", {}),
        ("smem stack 8", "smemstack8", {}),
        ("prefetch cur", "pfcur", {}),
        ("prefetch cur, 1 inner visit per census", "pfcur_ir1", {}),
        ("prefetch cur + treelet", "pfcur_treelet", {"treeletPrefetch": 1}),
        ("tri no_allocate", "tri_na", {}),
        ("tri evict_first", "tri_ef", {}),
        ("pairOrder 1", None, {"pairOrder": 1}),
        ("pairOrder 2", None, {"pairOrder": 2}),
        ("pairOrder 3", None, {"pairOrder": 3}),
        ("pairOrder 1 + next-pair prefetch", "pf", {"pairOrder": 1}),
        ("all_mesh", "all_mesh", {"treeletPrefetch": 1}),
    ]),
    2: (["knot64", "cluster4k", "soup4k"], [
        ("default", None, {}),
        ("leaf2", "leaf2", {}), ("ir1", "ir1", {}), ("ir3", "ir3", {}), ("rayinv", "rayinv", {}), ("pw20", "pw20", {}), ("pw16", "pw16", {}),
        ("glass branch out of line (instruction footprint)", "glassool", {}),
        ("sphere tree: surface-area sweep on the top 14 levels", "sphsah", {}),
        ("l2Persist", None, {"l2Persist": 1}), ("pairOrder 4", None, {"pairOrder": 4}), ("pairOrder 6", None, {"pairOrder": 6}),
        ("poolSlots 96 + treelet", "treelet", {"treeletPrefetch": 1, "poolSlots": 96}),
    ]),
    4: (["instances500", "instances16"], [       # TLAS over the models' world boxes (automatic above 64 models) against the linear per-model test
        ("default", None, {}), ("tlas off", None, {"tlas": 0}), ("tlas on", None, {"tlas": 1}), ("tlas on, kernel 1", None, {"tlas": 1, "kernel": 1}),
        ("tlas off, kernel 1", None, {"tlas": 0, "kernel": 1}), ("no model skipping at all", None, {"modelSkip": 0}),
    ]),
    5: (["soup4k", "cluster4k", "knot64"], [       # round 2, second call: combinations of the first call's winners
        ("default", None, {}),
        ("ldg256 + vote 1/3/2 + smem stack 8", "c1", {}),
        ("ldg256 + vote 1/3/2 + smem stack 8 + leaf2", "c2", {}),
        ("ldg256 + vote 1/3/2 + smem stack 8 + leaf2 + sphere SAH", "c3", {}),
        ("vote 1/3/2 + smem stack 8 + leaf2 + sphere SAH (128-bit loads)", "c4", {}),
        ("ldg256 + vote 2/7/4 + smem stack 8 + leaf2 + sphere SAH", "c5", {}),
        ("ldg256 + vote 1/3/2 + smem stack 16 + leaf2 + sphere SAH", "c6", {}),
        ("ldg256 + vote 1/3/2 + stack top in registers + leaf2 + sphere SAH", "c7", {}),
        ("c3 + branch-free pop from the shared-memory stack", "c8", {}),
        ("smem stack 8", "smemstack8", {}),
        ("smem stack 16", "smemstack16", {}),
        ("vote 1/3/2", "vote132", {}),
        ("28 warps per SM (72 registers)", "pw28", {}),
        ("32 warps per SM (64 registers)", "pw32", {}),
        ("28 warps + ldg256 + vote 1/3/2 + leaf2 + sphere SAH", "pw28_c", {}),
        ("32 warps + ldg256 + vote 1/3/2 + leaf2 + sphere SAH", "pw32_c", {}),
    ]),
    7: (["soup4k", "cluster4k", "knot64"], [        # round 2, third call: on top of the promoted defaults (rt_devmath.cuh)
        ("default (round-2 defaults)", None, {}),
        ("round-1 kernels (RT_DEFAULTS_R1)", "r1", {}),
        ("32 pool slots per warp (more L1)", None, {"poolSlots": 32}),
        ("28 warps x 32 slots", "pw28_m32", {"poolSlots": 32}),
        ("stack ring 4", "stack4", {}),
        ("stack ring 16", "stack16", {}),
        ("vote 2/7/4", "vote274", {}),
        ("3 leaf primitives per census", "leaf3", {}),
        ("64-byte triangle records", "tri64", {}),
        ("l2Persist", None, {"l2Persist": 1}),
        ("sortRays", None, {"sortRays": 1}),
        ("tailLanes 8", None, {"tailLanes": 8}),
        ("tailLanes 24", None, {"tailLanes": 24}),
    ]),
    8: (["soup4k", "cluster4k", "knot64"], [        # tree tops staged in shared memory by TMA bulk copy, now that the pools + the stack ring leave L1 ~20 KB
        ("default", None, {}),
        ("smemNodes 128 (TMA-staged tree tops)", None, {"smemNodes": 128}),
        ("smemNodes 256 (TMA-staged tree tops)", None, {"smemNodes": 256}),
        ("smemNodes 300 (TMA-staged tree tops)", None, {"smemNodes": 300}),
    ]),
    10: (["soup4k", "cluster4k", "knot64"], [ ("default", None, {}), ("predicated push", "pushpred", {}) ]),
    13: (["soup4k", "cluster4k", "knot64", "cornell64"], [ ("default", None, {}), ("streaming loads / stores for the render targets", "stream", {}), ("the same + l2Persist", "stream", {"l2Persist": 1}) ]),
    12: (["soup4k", "cluster4k", "knot64", "knot256"], [ ("default", None, {}), ("smemNodes 256 (extension instantiation)", None, {"smemNodes": 256}) ]),
    11: (["soup4k", "cluster4k", "knot64"], [ ("default", None, {}), ("shade-phase slot fields in global memory (66 KB more L1)", "cold", {}),
                                              ("the same + stack ring of 16", "cold_ring16", {}), ("the same + 96 slots per warp", "cold", {"poolSlots": 96}) ]),
    9: (["soup4k"], [        # sphere accelerator shape (config 5: 10,000 spheres)
        ("default (leaves of 4, SAH top 14)", None, {}),
        ("leaves of 1, SAH 20 levels", "sphleaf1", {}),
        ("leaves of 2, SAH 18 levels", "sphleaf2", {}),
        ("leaves of 8", "sphleaf8", {}),
        ("leaves of 4, SAH 20 levels", "sphsah20", {}),
    ]),
    6: (["instances62", "instances126", "instances250"], [       # where the TLAS starts to pay (automatic threshold)
        ("tlas off", None, {"tlas": 0}), ("tlas on", None, {"tlas": 1}),
    ]),
    3: (["cornell64", "cornell1"], [
        ("default", None, {}), ("glass out of line", "glassool", {}), ("zero-defocus + glass out of line + skipsqrt", "cornell_all", {}), ("zero-defocus shortcut", "zerodefocus", {}), ("zero-defocus + skipsqrt", "zerodefocus_skipsqrt", {}), ("skipsqrt", "skipsqrt", {}), ("mb5", "mb5", {}), ("gridFit", None, {"gridFit": 1}), ("kernel 2", None, {"kernel": 2}),
    ]),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stage", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--workloads", nargs="*", default=None)
    args = ap.parse_args()
    import torch
    import ray_tracing_b200 as rt
    from ray_tracing_b200 import build as b, scenes
    workloads, configs = STAGES[args.stage]
    os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
    log = open(os.path.join(REPO, "gpurun_out", "sweep_r2.jsonl"), "a")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for wl in (args.workloads or workloads):
        w = bench.WORKLOADS[wl]
        sc = bench.make_scene(w)
        base, base_digest = None, None
        for label, var, opts in configs:
            lib = variant(var) if var else b.LIB_CUDA
            if not os.path.exists(lib):
                print(f"-- {wl}: {label}: {lib} not built", flush=True)
                continue
            try:
                mgr = rt.RayComputeManager(lib)
                scenes.apply(sc, mgr)
                ctx = mgr.context
                for k, v in opts.items():
                    ctx.set_option(k, v)
                mgr.OnEnable()
                for i in range(args.warmup):
                    ctx.set_int("Frame", 1 + i); ctx.dispatch_full(0)
                ctx.synchronize(); ctx.reset_stats()
                for i in range(args.steps):
                    flush.zero_(); torch.cuda.synchronize()
                    ctx.set_int("Frame", 1 + args.warmup + i); ctx.dispatch_full(0)
                st = ctx.stats()
                ms = st["kernelMs"] / args.steps
                mrays = st["rays"] / (st["kernelMs"] * 1e-3) / 1e6
                digest = hashlib.sha256(mgr.accumulatedResult.tobytes()).hexdigest()[:16]     # same frames in every configuration: same bits expected
                mgr.OnDestroy()
            except Exception as e:                      # a variant that fails must not stop the sweep
                print(f"-- {wl}: {label}: {type(e).__name__}: {str(e)[:200]}", flush=True)
                continue
            if base is None:
                base, base_digest = ms, digest
            row = {"workload": wl, "config": label, "identical_to_default": digest == base_digest, "kernel_ms": round(ms, 4), "Mrays_s": round(mrays, 1), "vs_default": round(base / ms, 4), "options": opts, "lib": os.path.basename(lib)}
            print(json.dumps(row), flush=True)
            log.write(json.dumps(row) + "\n"); log.flush()


if __name__ == "__main__":
    main()
