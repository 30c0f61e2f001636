"""Shared fixtures.  Markers: `gpu` = needs a B200 (run by the driver with `-m gpu` on the GPU box)."""
import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

ORACLE_LIB = os.path.join(REPO, "oracle", "liboracle.so")
CUDA_LIB = os.environ.get("RT_B200_LIB") or os.path.join(REPO, "ray_tracing_b200", "librt_b200.so")   # RT_B200_LIB: test an experimental build
GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the native libraries exist (they travel pre-built to the GPU box; here we build on demand)."""
    from ray_tracing_b200 import build
    if not (os.path.exists(ORACLE_LIB) and os.path.exists(build.LIB_HOST)):
        build.build_host()
        build.build_oracle()
    if not os.path.exists(CUDA_LIB):
        build.build_cuda()


@pytest.fixture(scope="session")
def oracle_path():
    return ORACLE_LIB


@pytest.fixture(scope="session")
def cuda_path():
    return CUDA_LIB


def render(backend, scene, frames=1, width=None, height=None, options=None, want_stats=False, tile=None):
    """Drive `frames` RenderFrame() calls of the host manager against a backend library; returns (frame, accumulated[, stats])."""
    import ray_tracing_b200 as rt
    from ray_tracing_b200 import scenes
    mgr = rt.RayComputeManager(backend)
    scenes.apply(scene, mgr, width, height)
    ctx = mgr.context
    for k, v in (options or {}).items():
        ctx.set_option(k, v)
    if tile:
        ctx.set_tile(*tile)
    mgr.OnEnable()
    ctx.reset_stats()
    for _ in range(frames):
        mgr.RenderFrame()
    out = (mgr.raytraceFrameTex, mgr.accumulatedResult)
    if want_stats:
        out = out + (ctx.stats(),)
    mgr.OnDestroy()
    return out


def assert_bit_equal(a: np.ndarray, b: np.ndarray, what=""):
    """Bitwise comparison with a readable report.  A NaN must be a NaN on both sides; its sign / payload bits are not compared
    (x86 and the GPU generate different default NaNs for the same invalid operation)."""
    ai, bi = a.view(np.uint32), b.view(np.uint32)
    if np.array_equal(ai, bi):
        return
    both_nan = np.isnan(a) & np.isnan(b)
    bad = np.argwhere((ai != bi) & ~both_nan)
    if len(bad) == 0:
        return
    diff = np.abs(a.astype(np.float64) - b.astype(np.float64))
    raise AssertionError(f"{what}: {len(bad)} of {ai.size} values differ bitwise; max |Δ| = {np.nanmax(diff):.3e}; first at {bad[0].tolist()} "
                         f"({a[tuple(bad[0])]!r} vs {b[tuple(bad[0])]!r})")


def seed_forcing_draw(post_step_state, draw, pixel_index, frame=1):
    """renderSeed that makes the `draw`-th NextRandom of a pixel leave the generator in `post_step_state` (RayCommon.hlsl:127-133,552:
    rngState = pixelIndex + Frame * 719393 + renderSeed, then state = state * 747796405 + 2891336453 per draw — an invertible LCG)."""
    a_inv, c, m = pow(747796405, -1, 1 << 32), 2891336453, 1 << 32
    st = post_step_state
    for _ in range(draw):
        st = ((st - c) * a_inv) % m
    seed = (st - pixel_index - frame * 719393) % m
    return seed - m if seed >= (1 << 31) else seed
