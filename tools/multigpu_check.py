"""Run under torchrun on N GPUs: every rank traces its row bands, ONE NCCL all-gather per frame hands every rank the
whole frame; rank 0 checks the gathered result bit-for-bit against an untiled single-GPU render of the same frames."""
import os, sys
import numpy as np, torch, torch.distributed as dist
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import ray_tracing_b200 as rt
from ray_tracing_b200 import build as b, multigpu, scenes

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
ok = True
cases = [(n, s, f) for f in (False, True) for n, s in (("cornell", scenes.cornell_spheres(400, 300, 5, 3)), ("knot", scenes.knot_room(320, 180, 5, 2, nu=200, nv=12)))]
for name, sc, fused in cases:
    name = name + ("/fused-p2p" if fused else "/all-gather")
    mgr = rt.RayComputeManager(b.LIB_CUDA, device=local)
    scenes.apply(sc, mgr)
    tiled = multigpu.TiledRenderer(mgr, rank, world, band_rows=8, device=dev, fused=fused)
    mgr.OnEnable()
    for _ in range(3):
        tiled.render_frame()
    if fused:
        tiled.frame_fence()
    torch.cuda.synchronize()
    frame, accum = mgr.raytraceFrameTex, mgr.accumulatedResult
    if rank == 0:
        ref = rt.RayComputeManager(b.LIB_CUDA, device=local)
        scenes.apply(sc, ref)
        ref.OnEnable()
        for _ in range(3):
            ref.RenderFrame()
        f1, a1 = ref.raytraceFrameTex, ref.accumulatedResult
        same = np.array_equal(frame.view(np.uint32), f1.view(np.uint32)) and np.array_equal(accum.view(np.uint32), a1.view(np.uint32))
        print(f"multigpu_check {name}: world={world} bitwise_equal_to_single_gpu={same}", flush=True)
        ok = ok and same
    dist.barrier()
    mgr.OnDestroy()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
