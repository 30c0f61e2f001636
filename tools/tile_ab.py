#!/usr/bin/env python
"""What one GPU of an N-GPU run has to do, measured on ONE GPU: the trace kernel over rank 0's row bands (rtSetTile(0, N, 8)) for a
list of option sets, against the whole image / N.  Kernel time from CUDA events; no exchange.  For choosing the tail fixes without an
8-GPU box (the 8-GPU lines of bench.py are the numbers that count).

    python tools/tile_ab.py --world 8 --workloads knot256 cornell64
"""
import argparse
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench                                         # noqa: E402

CONFIGS = [("default (automatic)", {}), ("kernel 2, whole pixels", {"kernel": 2, "sampleChunks": 0}), ("kernel 2, 2 sample chunks", {"kernel": 2, "sampleChunks": 2}),
           ("kernel 2, 4 sample chunks", {"kernel": 2, "sampleChunks": 4}), ("kernel 2, 8 sample chunks", {"kernel": 2, "sampleChunks": 8}), ("kernel 2, 16 sample chunks", {"kernel": 2, "sampleChunks": 16}),
           ("kernel 1, whole pixels", {"kernel": 1, "sampleChunks": 0}), ("kernel 1, 2 sample chunks", {"kernel": 1, "sampleChunks": 2}), ("kernel 1, 4 sample chunks", {"kernel": 1, "sampleChunks": 4}),
           ("kernel 1, 8 sample chunks", {"kernel": 1, "sampleChunks": 8}), ("kernel 1, 16 sample chunks", {"kernel": 1, "sampleChunks": 16}),
           ("kernel 1, 32 sample chunks", {"kernel": 1, "sampleChunks": 32})]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, nargs="*", default=[8])
    ap.add_argument("--workloads", nargs="*", default=["knot256", "cornell64"])
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--lib", default=None, help="alternative build of librt_b200.so")
    ap.add_argument("--only", nargs="*", default=None, help="substrings of the configuration labels to run")
    args = ap.parse_args()
    import torch
    import ray_tracing_b200 as rt
    from ray_tracing_b200 import build as b, scenes
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    out = open(os.path.join(REPO, "gpurun_out", "tile_ab.jsonl"), "a")
    for wl in args.workloads:
        w = bench.WORKLOADS[wl]
        sc = bench.make_scene(w)
        full = None
        for world in [1] + list(args.world):
            for label, opts in (CONFIGS if world > 1 else CONFIGS[:1]) if not args.only else [c for c in CONFIGS if any(o in c[0] for o in args.only)]:
                mgr = rt.RayComputeManager(args.lib or b.LIB_CUDA)
                scenes.apply(sc, mgr)
                ctx = mgr.context
                for k, v in opts.items():
                    ctx.set_option(k, v)
                ctx.set_tile(0, world, 8)
                mgr.OnEnable()
                for i in range(2):
                    ctx.set_int("Frame", 1 + i); ctx.dispatch_full(0)
                ctx.synchronize(); ctx.reset_stats()
                for i in range(args.steps):
                    flush.zero_(); torch.cuda.synchronize()
                    ctx.set_int("Frame", 3 + i); ctx.dispatch_full(0)
                st = ctx.stats()
                ms = st["kernelMs"] / args.steps
                mgr.OnDestroy()
                if world == 1:
                    full = ms
                row = {"lib": os.path.basename(args.lib or b.LIB_CUDA), "workload": wl, "tile": f"rank 0 of {world}", "config": label, "kernel_ms": round(ms, 4), "ideal_ms": round(full / world, 4), "efficiency": round(full / world / ms, 3)}
                print(json.dumps(row), flush=True)
                out.write(json.dumps(row) + "\n"); out.flush()


if __name__ == "__main__":
    main()
