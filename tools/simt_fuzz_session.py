#!/usr/bin/env python
"""Randomised SESSIONS on the SIMT interpreter build: not one frame of a scene but a sequence of host calls — frames, camera moves
(near, far, back), models that move / change material, Spheres buffers that shrink and grow across the accelerator's threshold,
resizes, accumulation resets, settings and kernel options that change between frames — driven identically through the host
manager against the oracle and against the kernels, compared bit for bit after every frame.  What it exercises is the host-side
state of csrc/rt_api.cu: dirty flags, cached repacks, the regions the padded boxes were built for, kernel selection.

    python tools/simt_fuzz_session.py [--cases 100] [--seed 1] [--steps 8]
"""
import argparse
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
sys.path.insert(0, os.path.join(REPO, "tests", "simt"))
sys.path.insert(0, os.path.join(REPO, "tools"))
import build as simt_build                      # noqa: E402
from conftest import ORACLE_LIB                 # noqa: E402
import ray_tracing_b200 as rt                   # noqa: E402
from ray_tracing_b200 import scenes             # noqa: E402
from simt_fuzz import random_material, random_scene   # noqa: E402


def random_spheres(rng, n):
    sph = np.zeros(n, dtype=scenes.SPHERE_DTYPE)
    for i in range(n):
        sph[i] = scenes._sphere(tuple(rng.uniform(-2.5, 2.5, 3) + np.array([0, 2.0, 0])), float(np.exp(rng.uniform(np.log(0.02), np.log(0.6)))), random_material(rng))
    return sph


def make_script(rng, sc, steps):
    """A list of (action, arguments) — applied to both managers; ('option', ...) only to the kernels' side."""
    script = []
    nm = len(sc.models)
    for _ in range(steps):
        for _ in range(int(rng.randint(0, 4))):
            a = int(rng.randint(0, 11))
            if a == 10:
                # round 2: the ABI below the manager — modelCount set on its own (no ModelInfo re-send), then a dispatch
                script.append(("rawframe", (int(rng.randint(0, nm + 1)), int(rng.randint(1, 1000)))))
            elif a == 0:
                d = float(rng.choice([0.0, 3.0, 5.67, 60.0, 4000.0, 3e5]))
                script.append(("camera", (float(np.clip(np.degrees(2 * np.arctan(3.0 / max(d, 3.0))), 1e-3, 90.0)),
                                          scenes.trs(position=(float(rng.uniform(-1, 1)), 1.9, -d), euler_deg=(float(rng.uniform(-5, 5)), float(rng.uniform(-5, 5)), 0.0))[0])))
            elif a == 1 and nm:
                scale = np.exp(rng.uniform(np.log(0.02), np.log(1.2), 3))
                pos = tuple(rng.uniform(-2.5, 2.5, 3) * float(rng.choice([1.0, 1.0, 50.0])) + np.array([0, 2.0, 0]))
                script.append(("model", (int(rng.randint(0, nm)),) + scenes.trs(position=pos, euler_deg=tuple(rng.uniform(0, 360, 3)), scale=tuple(scale))))
            elif a == 2 and nm:
                script.append(("material", (int(rng.randint(0, nm)), random_material(rng))))
            elif a == 3:
                script.append(("spheres", (random_spheres(rng, int(rng.choice([0, 3, 64, 65, 150]))),)))
            elif a == 4:
                script.append(("screen", (int(rng.choice([8, 24, 40, 64])), int(rng.choice([6, 18, 36])))))
            elif a == 5:
                script.append(("reset", ()))
            elif a == 6:
                script.append(("setting", (str(rng.choice(["maxBounceCount", "numRaysPerPixel", "useSky", "accumulate", "divergeStrength", "defocusStrength"])), rng.rand())))
            else:
                name = str(rng.choice(["kernel", "tlas", "modelSkip", "poolSlots", "pairOrder", "smemNodes", "tailLanes", "sortRays", "gridFit", "extInstantiation", "countStats", "sampleChunks"]))
                value = {"kernel": [0, 1, 2, -1], "tlas": [-1, 0, 1], "modelSkip": [0, 1], "poolSlots": [0, 32, 64, 96], "pairOrder": [0, 1, 3], "smemNodes": [0, 9, 200, -1],
                         "tailLanes": [0, 5, 16, 31], "sortRays": [0, 1], "gridFit": [0, 1], "extInstantiation": [0, 1], "countStats": [0, 1], "sampleChunks": [-1, 0, 2, 3]}[name]
                script.append(("option", (name, int(rng.choice(value)))))
        script.append(("frame", ()))
    return script


def apply(mgr, action, args, kernels_side):
    if action == "camera":
        mgr.set_camera(*args)
    elif action == "model":
        mgr.set_model_transform(*args)
    elif action == "material":
        mgr.set_model_material(*args)
    elif action == "spheres":
        mgr.set_spheres(*args)
    elif action == "screen":
        mgr.set_screen(*args)
    elif action == "reset":
        mgr.ResetAccumulatedRender()
    elif action == "setting":
        name, u = args
        value = {"maxBounceCount": int(u * 7), "numRaysPerPixel": 1 + int(u * 3), "useSky": u < 0.5, "accumulate": u < 0.7,
                 "divergeStrength": float(u * 2), "defocusStrength": float(u < 0.3) * 30.0}[name]
        setattr(mgr, name, value)
    elif action == "option" and kernels_side:
        mgr.context.set_option(*args)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=100)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--lib", default=None)
    args = ap.parse_args()
    lib = args.lib or simt_build.build()
    bad = 0
    for case in range(args.cases):
        seed = args.seed * 100003 + case
        rng = np.random.RandomState(seed)
        sc = random_scene(rng, odd=0.0)
        script = make_script(rng, sc, args.steps)
        mgrs = []
        for backend in (ORACLE_LIB, lib):
            m = rt.RayComputeManager(backend)
            scenes.apply(sc, m)
            m.OnEnable()
            mgrs.append(m)
        frame_no, failed, resized = 0, None, False
        try:
            for action, a in script:
                if action == "screen":
                    resized = True                      # the textures follow at the next RenderFrame (InitFrame): no raw dispatch in between
                if action == "rawframe" and resized:
                    continue
                if action == "frame":
                    resized = False
                if action in ("frame", "rawframe"):
                    for m in mgrs:
                        if action == "frame":
                            m.RenderFrame()
                        else:
                            m.context.set_int("modelCount", a[0]); m.context.set_int("Frame", a[1]); m.context.dispatch_full(0)
                    frame_no += 1
                    for what in ("accumulatedResult", "raytraceFrameTex"):
                        x, y = getattr(mgrs[1], what), getattr(mgrs[0], what)
                        diff = (x.view(np.uint32) != y.view(np.uint32)) & ~(np.isnan(x) & np.isnan(y))
                        if x.shape != y.shape or diff.any():
                            failed = f"{what} differs after frame {frame_no} ({int(diff.sum()) if x.shape == y.shape else 'shape'})"
                            break
                    if failed:
                        break
                else:
                    apply(mgrs[0], action, a, False)
                    apply(mgrs[1], action, a, True)
        except Exception as e:
            failed = f"{type(e).__name__}: {e}"
        for m in mgrs:
            m.OnDestroy()
        if failed:
            bad += 1
            brief = [(act, a if act in ("option", "setting", "screen") else "...") for act, a in script]
            print(f"case seed {seed}: {failed}; {len(sc.models)} models, {len(sc.spheres)} spheres; script {brief}", flush=True)
        if (case + 1) % 25 == 0:
            print(f"  {case + 1} cases, {bad} findings", flush=True)
    print(f"{args.cases} cases, {bad} findings")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
