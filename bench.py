#!/usr/bin/env python
"""bench.py — the headline measurement: Mrays/s and ms/frame of the path-tracing hot path at 1920x1080, 8 bounces.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run, one rank per GPU)
    python bench.py --impl reference ...                     the reference's CPU statement of the path (oracle) on the host cores

A "step" is one frame: one RayTrace dispatch over the whole image with NumRaysPerPixel = 64 (BASELINE.json configs[1]:
9-sphere Cornell box, 1920x1080, 8 bounces, 64 spp on 1 x B200; the (R = spp, F = 1) split of SURVEY.md §8d).
A "ray" is one CalculateRayCollision call (RayCommon.hlsl:487), counted by the kernel itself.

  value     whole-job Mrays/s with everything resident in HBM: per step only the Frame uniform changes (no buffer
            upload), the dispatch, and for N > 1 the per-frame all-gather of finished tiles.
  e2e       the same metric through the public host API (RayComputeManager.RenderFrame -> C-ABI) with HOST buffers:
            every step re-uploads ModelInfo / Spheres / uniforms from host memory (as the reference does each frame,
            RayComputeManager.cs:192-204) and reads the accumulated float4 image back into pinned host memory.
  roofline  HBM roofline of the dominant kernel (k_raytrace_wave) on ALGORITHMIC bytes:
            32*boxTests + 72*triTests + 224*rays*modelCount + 104*sphereTests + 48*W*H per frame, counts taken from an
            instrumented replay of the very frames that were timed; peak from MEASURED_PEAKS.json.
  cpu_baseline  the oracle (CPU port of the reference shader) on this box's host cores, bounded pixel sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

WORKLOADS = {
    # name: (scene factory kwargs, description)
    "cornell64": dict(kind="cornell", width=1920, height=1080, bounces=8, spp=64,
                      desc="9-sphere Cornell box, 1920x1080, 8 bounces, 64 spp per frame (BASELINE.json configs[1])"),
    "knot64": dict(kind="knot", width=1920, height=1080, bounces=8, spp=16,
                   desc="87,132-triangle knot in a Cornell room (3 models, BVH), 1920x1080, 8 bounces, 16 spp per frame (configs[2] shape)"),
    # the reference's own usage: one sample per pixel per frame, many accumulated frames (all five shipped scenes)
    "cornell1": dict(kind="cornell", width=1920, height=1080, bounces=8, spp=1,
                     desc="9-sphere Cornell box, 1920x1080, 8 bounces, 1 spp per frame (the (R = 1, F = spp) split of configs[1])"),
    "knot1": dict(kind="knot", width=1920, height=1080, bounces=8, spp=1,
                  desc="87,132-triangle knot in a Cornell room (3 models, BVH), 1920x1080, 8 bounces, 1 spp per frame"),
    "instances16": dict(kind="instances", width=1920, height=1080, bounces=8, spp=16,
                        desc="24 models sharing two meshes (a 7,680-triangle knot instanced 23 times, glass / opaque / emissive, + room), "
                             "1920x1080, 8 bounces, 16 spp per frame (the many-Model shape of the reference's shipped scenes)"),
    "instances500": dict(kind="instances", width=1920, height=1080, bounces=8, spp=8, instances=498,
                         desc="500 models sharing two meshes (a 7,680-triangle knot instanced 498 times + room + light), 1920x1080, 8 bounces, "
                              "8 spp per frame (many-Model shape: the TLAS over the models' world boxes, option tlas, is on automatically above 64 models)"),
    "cluster4k": dict(kind="cluster", width=3840, height=2160, bounces=12, spp=4,
                      desc="871,212-triangle glass knot cluster (one mesh, one deep BVH) in a room, 3840x2160, 12 bounces, 4 spp per frame (configs[3] shape)"),
    "soup4k": dict(kind="soup", width=4096, height=4096, bounces=16, spp=2,
                   desc="1,000,000 random triangles in 3 models + 10,000 spheres, sky on, 4096x4096, 16 bounces, 2 spp per frame (configs[4] shape)"),
}
METRIC = "Mrays/s at 1920x1080, 8 bounces (ray = one CalculateRayCollision call)"
FALLBACK_HBM_GBS = 6650.0


def make_scene(w):
    from ray_tracing_b200 import scenes
    if w["kind"] == "cornell":
        return scenes.cornell_spheres(w["width"], w["height"], w["bounces"], w["spp"])
    if w["kind"] == "instances":
        return scenes.instanced_knots(w["width"], w["height"], w["bounces"], w["spp"], instances=w.get("instances", 23))
    if w["kind"] == "cluster":
        return scenes.knot_cluster(w["width"], w["height"], w["bounces"], w["spp"])
    if w["kind"] == "soup":
        return scenes.random_soup(w["width"], w["height"], w["bounces"], w["spp"], triangles=1_000_000, spheres=10_000)
    return scenes.knot_room(w["width"], w["height"], w["bounces"], w["spp"])


def algorithmic_bytes(st, model_count, width, height, frames):
    return (32 * (st["boxTests"] + st.get("sphereBoxTests", 0)) + 72 * st["triTests"] + 224 * st["rays"] * model_count + 104 * st["sphereTests"]
            + 48 * width * height * frames)


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle on the host cores (cpu_baseline leg and --impl reference)
# ---------------------------------------------------------------------------------------------------------------------

def effective_cores():
    """Host threads this process can really use: the affinity mask, capped by a cgroup CPU quota if one is set."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = max(1, min(n, int(float(quota) / float(period) + 0.5)))
    except Exception:
        pass
    return n


def cpu_sample(w, steps, warmup, pixels=262144):
    """Times the oracle on a seeded sparse pixel sample of the workload (exact: pixels are independent and seeded from
    their global index).  Returns (Mrays/s, ms per step, cores, sample description)."""
    import ctypes as C
    import ray_tracing_b200 as rt
    from ray_tracing_b200 import build as b, scenes
    sc = make_scene(w)
    mgr = rt.RayComputeManager(b.LIB_ORACLE)
    scenes.apply(sc, mgr)
    mgr.OnEnable()
    L = C.CDLL(b.LIB_ORACLE)
    rng = np.random.RandomState(11)
    xy = np.stack([rng.randint(0, w["width"], pixels), rng.randint(0, w["height"], pixels)], axis=1).astype(np.int32)
    out = np.empty((pixels, 4), dtype=np.float32)
    ctx = mgr.context
    ctx.set_option("threads", effective_cores())
    handle = C.c_void_p(ctx.handle.value)
    times, rays = [], 0
    for i in range(warmup + steps):
        ctx.set_int("Frame", 1 + i)
        ctx.reset_stats()
        t0 = time.perf_counter()
        rc = L.orRenderPixels(handle, xy.ctypes.data_as(C.c_void_p), pixels, out.ctypes.data_as(C.c_void_p))
        dt = time.perf_counter() - t0
        assert rc == 0
        if i >= warmup:
            times.append(dt)
            rays += ctx.stats()["rays"]
    cores = effective_cores()
    total = sum(times)
    sample = f"{pixels} seeded random pixels x {w['spp']} spp of the {w['width']}x{w['height']} frame per step, {cores} threads"
    return rays / total / 1e6, 1e3 * total / max(len(times), 1), cores, sample


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, ms, cores, sample = cpu_sample(w, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["desc"], "note": "CPU restatement of RayCommon.hlsl (oracle/), the reference HLSL/C# cannot run here"},
        "cpu_baseline": {"value": round(value, 3), "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": round(value, 3), "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------------

class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.proc = index, [], None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.samples.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        rows = [r for r in self.samples if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[3 + k].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "reasons": reasons, "samples": len(rows)}


def ncu_traffic(workload, kernel_name):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu capture of this
    workload (profiles/ncu_traffic.json, written from `ncu --set full` reports by tools/ncu_traffic.py); None if not captured."""
    p = os.path.join(REPO, "profiles", "ncu_traffic.json")
    try:
        d = json.load(open(p)).get(workload, {})
        return d.get(kernel_name.split(" ")[0])
    except Exception:
        return None


def measured_hbm_peak():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def run_gpu(args, w):
    import torch
    import torch.distributed as dist
    import ray_tracing_b200 as rt
    from ray_tracing_b200 import build as b, multigpu, scenes

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    sc = make_scene(w)
    W, H = w["width"], w["height"]
    mgr = rt.RayComputeManager(args.lib or b.LIB_CUDA, device=local)   # raises without the CUDA library / a GPU
    scenes.apply(sc, mgr)
    tiled = multigpu.TiledRenderer(mgr, rank, world, band_rows=args.band_rows, device=dev, fused=args.exchange == "fused")   # also puts the context on a torch stream
    ctx, stream = tiled.ctx, tiled.stream
    if args.kernel is not None:
        ctx.set_option("kernel", args.kernel)
    if args.pool_slots is not None:
        ctx.set_option("poolSlots", args.pool_slots)
    if args.smem_nodes is not None:
        ctx.set_option("smemNodes", args.smem_nodes)
    if args.model_skip is not None:
        ctx.set_option("modelSkip", args.model_skip)
    if args.sort_rays is not None:
        ctx.set_option("sortRays", args.sort_rays)
    if args.tail_lanes is not None:
        ctx.set_option("tailLanes", args.tail_lanes)
    if args.pair_order is not None:
        ctx.set_option("pairOrder", args.pair_order)
    if args.grid_fit is not None:
        ctx.set_option("gridFit", args.grid_fit)
    if args.l2_persist is not None:
        ctx.set_option("l2Persist", args.l2_persist)
    if args.treelet_prefetch is not None:
        ctx.set_option("treeletPrefetch", args.treelet_prefetch)
    if args.tlas is not None:
        ctx.set_option("tlas", args.tlas)
    mgr.OnEnable()
    if tiled.fused:
        with torch.cuda.stream(stream):
            tiled._connect_peers()
            dist.barrier()
    model_count = len(sc.models)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)    # > 126 MB L2
    pinned = torch.empty((H, W, 4), dtype=torch.float32).pin_memory()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def step_resident(frame_no):
        with torch.cuda.stream(stream):
            ctx.set_int("Frame", frame_no)
            if tiled.fused:
                tiled.frame_fence()
                ctx.dispatch_full(0)
                tiled.frame_fence()
            else:
                ctx.dispatch_full(0)
            if world > 1 and not tiled.fused:
                send, recv = tiled._views()
                ctx.pack_tile()
                dist.all_gather_into_tensor(recv, send)
                ctx.unpack_tiles()
            flush.zero_()                                             # evict L2 between steps

    def step_e2e():
        tiled.render_frame()                                          # RenderFrame(): host -> device uploads + dispatch (+ all-gather)
        with torch.cuda.stream(stream):
            mgr.read_accumulated_into(pinned.data_ptr(), pinned.numel() * 4)   # device -> pinned host, synchronises

    # ---- device-resident timing ------------------------------------------------------------------------------------
    frame_no = 1
    sampler = ClockSampler(local); sampler.start()                   # samples cover warm-up + timed region (clocks under load)
    for _ in range(args.warmup):
        step_resident(frame_no); frame_no += 1
    barrier()
    ctx.reset_stats()
    first_timed = frame_no
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with torch.cuda.stream(stream):
        e0.record(stream)
    for _ in range(args.steps):
        step_resident(frame_no); frame_no += 1
    with torch.cuda.stream(stream):
        e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    st = ctx.stats()
    rays_local, kernel_ms = st["rays"], st["kernelMs"]

    # ---- instrumented replay of the timed frames: exact traversal counts for the roofline -----------------------
    ctx.set_option("countStats", 1)
    ctx.reset_stats()
    for f in range(first_timed, first_timed + args.steps):
        with torch.cuda.stream(stream):
            ctx.set_int("Frame", f)
            ctx.dispatch_full(0)
    cst = ctx.stats()
    ctx.set_option("countStats", 0)
    assert cst["rays"] == rays_local, "instrumented replay traced different rays"

    # ---- end to end through the host API ---------------------------------------------------------------------------------
    mgr.ResetAccumulatedRender()
    for _ in range(args.warmup):
        step_e2e()
    barrier()
    ctx.reset_stats()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_rays_local = ctx.stats()["rays"]
    clocks = sampler.stop()

    # ---- reduce over ranks --------------------------------------------------------------------------------------------------
    if world > 1:
        t = torch.tensor([ms_total, e2e_s, kernel_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, e2e_s, kernel_ms_max = [float(x) for x in t.tolist()]
        r = torch.tensor([rays_local, e2e_rays_local, cst["boxTests"], cst["triTests"], cst["sphereTests"]], dtype=torch.int64, device=dev)
        dist.all_reduce(r, op=dist.ReduceOp.SUM)
        rays, e2e_rays, box, tri, sph = [int(x) for x in r.tolist()]
    else:
        kernel_ms_max = kernel_ms
        rays, e2e_rays, box, tri, sph = rays_local, e2e_rays_local, cst["boxTests"], cst["triTests"], cst["sphereTests"]

    if rank == 0:
        value = rays / (ms_total * 1e-3) / 1e6
        e2e_value = e2e_rays / e2e_s / 1e6
        peak, peak_src = measured_hbm_peak()
        # roofline of the dominant kernel on THIS rank's launches (per launch = per frame)
        alg = algorithmic_bytes({"boxTests": cst["boxTests"], "triTests": cst["triTests"], "rays": cst["rays"], "sphereTests": cst["sphereTests"],
                                 "sphereBoxTests": cst.get("sphereBoxTests", 0)},
                                model_count, W, H // world if world > 1 else H, args.steps)
        achieved = alg / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else 0.0
        kernel_label = {2: "k_raytrace_pool (persistent wavefront, per-warp path pools)",
                        1: "k_raytrace_wave (persistent threads, one path per lane)", 0: "k_raytrace_mega (reference-shaped)"}[
                            args.kernel if args.kernel is not None else (2 if model_count > 0 else 1)]
        h2d = 224 * model_count + 104 * len(sc.spheres) + 4 * 40
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_total / args.steps, 4), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["desc"], "rays_per_frame": rays // args.steps, "spp_per_frame": w["spp"],
                       "tiling": (f"row bands of {args.band_rows} rows round-robin over {world} GPU(s), " +
                                  ("finished pixels stored into the peers' images by the trace kernel (CUDA-IPC over NVLink), 4-byte all-reduce as frame fence"
                                   if tiled.fused else "one NCCL all-gather of finished tiles per frame")) if world > 1 else "single GPU",
                       "l2": "flushed between steps (256 MiB write inside the timed region)",
                       "kernel": kernel_label,
                       "pool_slots": args.pool_slots, "smem_nodes": args.smem_nodes, "pair_order": args.pair_order, "grid_fit": args.grid_fit, "tlas": args.tlas},
            "ms_per_frame": round(ms_total / args.steps, 4),
            "clocks": clocks,
            "e2e": {"value": round(e2e_value, 2), "unit": "Mrays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": W * H * 16,
                    "ms_per_step": round(1e3 * e2e_s / args.steps, 4)},
            "gpu_launches": args.steps * (1 + (2 if (world > 1 and not tiled.fused) else 0)),
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                         "traffic": ncu_traffic(args.workload, kernel_label) if world == 1 else None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg // args.steps, "kernel_ms_per_launch": round(kernel_ms / args.steps, 4),
                         "counts_per_launch": {"rays": cst["rays"] // args.steps, "boxTests": cst["boxTests"] // args.steps,
                                               "triTests": cst["triTests"] // args.steps, "sphereTests": cst["sphereTests"] // args.steps},
                         "note": "algorithmic bytes per SURVEY.md 8(d) from the reference's own test counters; sphere-only scenes keep their spheres in shared memory, so their DRAM traffic is far below this and frac can exceed 1"},
        }
        if world == 1 and not args.no_cpu:
            v, ms, cores, sample = cpu_sample(w, 1, 0)
            line["cpu_baseline"] = {"value": round(v, 3), "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cornell64", choices=sorted(WORKLOADS))
    ap.add_argument("--band-rows", type=int, default=8)
    ap.add_argument("--exchange", default="allgather", choices=["allgather", "fused"],
                    help="N > 1: one NCCL all-gather of finished tiles per frame, or pixels stored into the peers' images by the trace kernel (CUDA-IPC over NVLink)")
    ap.add_argument("--kernel", type=int, default=None, help="0 = megakernel, 1 = persistent threads, 2 = pooled wavefront (default)")
    ap.add_argument("--pool-slots", type=int, default=None, help="paths per warp pool of kernel 2 (32, 64, 96)")
    ap.add_argument("--model-skip", type=int, default=None, help="kernels 1/2: skip models the ray cannot reach (1 default / 0)")
    ap.add_argument("--tlas", type=int, default=None, help="kernels 1/2: tree over the models' world boxes: -1 = automatic (above 64 models, default), 0 = linear test, 1 = on")
    ap.add_argument("--sort-rays", type=int, default=None, help="kernel 2: group the ray queue by direction octant (1/0)")
    ap.add_argument("--tail-lanes", type=int, default=None, help="kernel 2: leave the trace phase when this few lanes still trace")
    ap.add_argument("--treelet-prefetch", type=int, default=None, help="1 = flagged two-level treelets + L1 prefetch of both next records (needs --lib built with RT_TREELET_PREFETCH)")
    ap.add_argument("--l2-persist", type=int, default=None, help="1 = persisting L2 window over the node-pair records")
    ap.add_argument("--grid-fit", type=int, default=None, help="1 = size the persistent grid for a whole number of pixels per lane (multi-GPU tail), 0 = default")
    ap.add_argument("--pair-order", type=int, default=None, help="node-pair record order: 0 = breadth-first (default), d = treelets of d levels, depth-first")
    ap.add_argument("--smem-nodes", type=int, default=None, help="node pairs staged in shared memory (-1 = auto)")
    ap.add_argument("--lib", default=None, help="alternative build of librt_b200.so (A/B experiments)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_gpu(args, w)


if __name__ == "__main__":
    main()
