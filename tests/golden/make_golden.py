"""Regenerates the committed fixtures in tests/golden/.

  pcg_kat.json        integer known-answer vectors of the PCG hash (RayCommon.hlsl:127-137), computed here with
                      arbitrary-precision Python integers — independent of both the oracle and the CUDA code.
                      (They equal the hand-derived table in SURVEY.md Appendix A.)
  cornell_c1.json     config 1 (9-sphere Cornell box, 256x256, 4 bounces, 1 spp, renderSeed 12345, 2 frames) as
                      rendered by the CPU oracle: SHA-256 of the float32 SUM buffer + 64 probe pixels.
                      A regression fixture of the oracle itself — the reference has no golden images.

  soup_small.json     the config-5 shape at a size the oracle renders in a second (20,000 random triangles in three models — diffuse,
                      glass, emissive — + 300 spheres, sky and sun, 96x96, 8 bounces, 2 spp, 2 frames): SHA-256 of both buffers, the
                      reference's traversal counters and 32 probe pixels.  Pins the BVH builder, the traversal order (counters), the
                      model loop, glass and sky against drift of the oracle; its inputs come from exact arithmetic on a Mersenne-
                      twister stream (no transcendental numpy call), so the scene is the same on every host.

Run from the repo root:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))


def pcg_next(state: int):
    state = (state * 747796405 + 2891336453) & 0xFFFFFFFF
    result = (((state >> ((state >> 28) + 4)) ^ state) * 277803737) & 0xFFFFFFFF
    result = (result >> 22) ^ result
    return state, result


def make_pcg():
    out = []
    for s0 in (0, 1, 731738, 4294967295, 12345, 0x80000000, 2891336453):
        s, seq = s0, []
        for _ in range(8):
            s, r = pcg_next(s)
            # RandomValue: float(r) / 2^32, with r rounded to nearest-even float32 first
            seq.append({"state": s, "result": r, "value_bits": int(np.float32(np.float32(r) / np.float32(4294967296.0)).view(np.uint32))})
        out.append({"seed": s0, "draws": seq})
    json.dump(out, open(os.path.join(HERE, "pcg_kat.json"), "w"), indent=1)


def make_cornell():
    from conftest import render, ORACLE_LIB
    from ray_tracing_b200 import scenes
    sc = scenes.cornell_spheres(256, 256, 4, 1)
    frame, accum, st = render(ORACLE_LIB, sc, frames=2, want_stats=True)
    rng = np.random.RandomState(1)
    probes = [(int(y), int(x)) for y, x in zip(rng.randint(0, 256, 64), rng.randint(0, 256, 64))]
    fix = {
        "config": "cornell_spheres(256,256,max_bounces=4,rays_per_pixel=1), renderSeed 12345, 2 frames",
        "accum_sha256": hashlib.sha256(accum.tobytes()).hexdigest(),
        "frame_sha256": hashlib.sha256(frame.tobytes()).hexdigest(),
        "rays": int(st["rays"]), "sphereTests": int(st["sphereTests"]),
        "probes": [{"y": y, "x": x, "accum_bits": [int(v) for v in accum[y, x].view(np.uint32)]} for y, x in probes],
    }
    json.dump(fix, open(os.path.join(HERE, "cornell_c1.json"), "w"), indent=1)


def soup_scene():
    from ray_tracing_b200 import scenes
    return scenes.random_soup(96, 96, max_bounces=8, rays_per_pixel=2, triangles=20000, spheres=300)


def make_soup():
    from conftest import render, ORACLE_LIB
    frame, accum, st = render(ORACLE_LIB, soup_scene(), frames=2, want_stats=True)
    rng = np.random.RandomState(2)
    probes = [(int(y), int(x)) for y, x in zip(rng.randint(0, 96, 32), rng.randint(0, 96, 32))]
    fix = {
        "config": "random_soup(96,96,max_bounces=8,rays_per_pixel=2,triangles=20000,spheres=300), renderSeed 12345, 2 frames",
        "accum_sha256": hashlib.sha256(accum.tobytes()).hexdigest(),
        "frame_sha256": hashlib.sha256(frame.tobytes()).hexdigest(),
        "rays": int(st["rays"]), "boxTests": int(st["boxTests"]), "triTests": int(st["triTests"]), "sphereTests": int(st["sphereTests"]),
        "probes": [{"y": y, "x": x, "accum_bits": [int(v) for v in accum[y, x].view(np.uint32)]} for y, x in probes],
    }
    json.dump(fix, open(os.path.join(HERE, "soup_small.json"), "w"), indent=1)


if __name__ == "__main__":
    make_pcg()
    make_cornell()
    make_soup()
    print("golden fixtures written to", HERE)
