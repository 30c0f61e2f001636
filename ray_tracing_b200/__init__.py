"""ray_tracing_b200 — a Blackwell-native (sm_100a) replacement for the path-tracing hot path of
SebLague/Ray-Tracing, behind the reference's host API.

    capi     ctypes binding of the C-ABI (include/rt_b200.h) implemented by librt_b200.so (CUDA only)
    manager  RayComputeManager mirror + BVH builder (C++ host in librt_host.so)
    scenes   deterministic synthetic scenes for the BASELINE.json configurations
    unity_scene  reader for the reference's serialized .unity scenes (manager settings, camera, models, meshes by guid)
    display  RayTraceDisplay (Display.shader path) and PNG screenshots
    build    in-tree build of the native libraries

There is no CPU fallback anywhere in this package: without librt_b200.so or without a B200-class GPU the
entry points raise.
"""
from . import capi, scenes          # noqa: F401
from .capi import RtLib, RtContext, RtError   # noqa: F401
from .manager import RayComputeManager, build_bvh, load_obj, set_build_threads   # noqa: F401
from .display import RayTraceDisplay, write_png, write_exr, read_exr   # noqa: F401

__all__ = ["capi", "scenes", "RtLib", "RtContext", "RtError", "RayComputeManager", "build_bvh", "load_obj", "RayTraceDisplay", "write_png", "write_exr", "read_exr"]
