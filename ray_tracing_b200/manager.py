"""Python face of the C++ host (librt_host.so): the BVH builder and the RayComputeManager mirror.

`RayComputeManager` keeps the reference's public surface (Assets/Scripts/Tracer/RayComputeManager.cs:9-95):
the inspector fields by the same names, `ResetAccumulatedRender()`, `RenderFrame()`, and the two render
textures (`raytraceFrameTex`, `accumulatedResult`) as readbacks.  The work is done by the C++ class of the
same name in host/RayComputeManager.cpp, which drives the C-ABI (include/rt_b200.h).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

from . import capi
from .capi import MATERIAL_DTYPE, NODE_DTYPE, SPHERE_DTYPE, TRIANGLE_DTYPE

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
HOST_LIB = os.path.join(PKG_DIR, "librt_host.so")

QUALITY = {"Low": 0, "High": 1, "Disabled": 2}

_host = None


def host_lib() -> C.CDLL:
    global _host
    if _host is None:
        if not os.path.exists(HOST_LIB):
            raise FileNotFoundError(f"{HOST_LIB} not found — run `python -m ray_tracing_b200.build`")
        L = C.CDLL(HOST_LIB)
        vp, cp, ci, cf = C.c_void_p, C.c_char_p, C.c_int, C.c_float
        L.rthBuildBVH.argtypes = [vp, ci, vp, ci, vp, ci, vp, vp, ci, vp]
        L.rthBuildBVH.restype = ci
        L.rthLastError.restype = cp
        L.rthObjLoad.argtypes = [cp, ci, C.POINTER(ci), C.POINTER(ci)]
        L.rthObjCopy.argtypes = [vp, vp, vp]
        L.rthObjLastError.restype = cp
        L.rcmCreate.argtypes = [cp, ci, C.POINTER(vp)]
        L.rcmCreateMulti.argtypes = [cp, C.POINTER(ci), ci, C.POINTER(vp)]
        L.rcmDestroy.argtypes = [vp]
        L.rcmLastError.argtypes = [vp]
        L.rcmLastError.restype = cp
        L.rcmContext.argtypes = [vp]
        L.rcmContext.restype = vp
        L.rcmSetInt.argtypes = [vp, cp, ci]
        L.rcmGetInt.argtypes = [vp, cp, C.POINTER(ci)]
        L.rcmSetFloat.argtypes = [vp, cp, cf]
        L.rcmSetSun.argtypes = [vp, vp, vp]
        L.rcmSetScreen.argtypes = [vp, ci, ci]
        L.rcmSetCamera.argtypes = [vp, cf, vp]
        L.rcmAddMesh.argtypes = [vp, vp, ci, vp, ci, vp]
        L.rcmAddModel.argtypes = [vp, ci, vp, vp, vp]
        L.rcmSetModelTransform.argtypes = [vp, ci, vp, vp]
        L.rcmSetModelMaterial.argtypes = [vp, ci, vp]
        L.rcmSetSpheres.argtypes = [vp, vp, ci]
        for n in ("rcmOnEnable", "rcmResetAccumulatedRender", "rcmRenderFrame"):
            getattr(L, n).argtypes = [vp]
        L.rcmReadFrame.argtypes = [vp, vp, C.c_size_t]
        L.rcmReadAccumulated.argtypes = [vp, vp, C.c_size_t]
        L.rcmGetBVHStats.argtypes = [vp, vp, ci]
        _host = L
    return _host


STAT_NAMES = ("TimeMs", "TriangleCount", "TotalNodeCount", "LeafNodeCount", "LeafDepthMin", "LeafDepthMax",
              "LeafDepthSum", "LeafMinTriCount", "LeafMaxTriCount")


def build_bvh(vertices: np.ndarray, indices: np.ndarray, normals: np.ndarray, quality: str | int = "High"):
    """BVH(verts, indices, normals, quality) of the reference (BVH.cs:26).  Returns (triangles, nodes, stats)."""
    L = host_lib()
    v = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 3)
    n = np.ascontiguousarray(normals, dtype=np.float32).reshape(-1, 3)
    idx = np.ascontiguousarray(indices, dtype=np.int32).reshape(-1)
    ntri = idx.size // 3
    tris = np.zeros(max(ntri, 1), dtype=TRIANGLE_DTYPE)
    nodes = np.zeros(2 * max(ntri, 1) + 1, dtype=NODE_DTYPE)
    stats = np.zeros(9, dtype=np.int32)
    q = QUALITY[quality] if isinstance(quality, str) else int(quality)
    rc = L.rthBuildBVH(v.ctypes.data, v.shape[0], idx.ctypes.data, idx.size, n.ctypes.data, q,
                       tris.ctypes.data, nodes.ctypes.data, nodes.shape[0], stats.ctypes.data)
    if rc < 0:
        raise ValueError("rthBuildBVH: " + L.rthLastError().decode())
    return tris[:ntri].copy(), nodes[:rc].copy(), dict(zip(STAT_NAMES, (int(s) for s in stats)))


def set_build_threads(threads: int) -> int:
    """Host threads a BVH build may use (0 = all, 1 = single-threaded like the reference); the result does not depend on it."""
    return int(host_lib().rthSetBuildThreads(int(threads)))


def load_obj(path: str, unity_handedness: bool = True):
    """Wavefront OBJ -> (vertices (n,3) f32, indices (3t,) i32, normals (n,3) f32) as Unity's importer hands them to the
    reference's manager (host/ObjLoader.cpp)."""
    L = host_lib()
    nv, ni = C.c_int(), C.c_int()
    if L.rthObjLoad(path.encode(), 1 if unity_handedness else 0, C.byref(nv), C.byref(ni)) != 0:
        raise ValueError("load_obj: " + L.rthObjLastError().decode())
    v = np.empty((nv.value, 3), dtype=np.float32)
    n = np.empty((nv.value, 3), dtype=np.float32)
    idx = np.empty(ni.value, dtype=np.int32)
    L.rthObjCopy(v.ctypes.data, n.ctypes.data, idx.ctypes.data)
    return v, idx, n


def column_major(m) -> np.ndarray:
    """4x4 (row, col) matrix -> 16 floats in Unity Matrix4x4 memory order."""
    return np.ascontiguousarray(np.asarray(m, dtype=np.float32).reshape(4, 4).T).reshape(16)


class RayComputeManager:
    """Mirror of the reference's RayComputeManager (public fields + ResetAccumulatedRender / RenderFrame)."""

    _INT_FIELDS = ("maxBounceCount", "numRaysPerPixel", "renderSeed", "numAccumulatedFrames", "bvhQuality",
                   "rayTracingEnabled", "accumulate", "useSky", "randomizeSeedOnEnable", "buildBVHOnDevice")
    _FLOAT_FIELDS = ("defocusStrength", "divergeStrength", "focusDistance", "sunFocus", "sunIntensity")

    def __init__(self, backend_library: Optional[str] = None, device: int = 0, devices: Optional[Sequence[int]] = None):
        """device: the CUDA ordinal to render on; devices: several GPUs of this process behind ONE manager (rtCreateMulti: the
        image is row-tiled over them and every RenderFrame ends with one NCCL all-gather of the frame's tiles)."""
        object.__setattr__(self, "_h", None)
        L = host_lib()
        lib_path = backend_library or capi.DEFAULT_LIB
        if not os.path.exists(lib_path):
            raise FileNotFoundError(f"{lib_path} not found (librt_b200 is CUDA-only; there is no CPU fallback)")
        h = C.c_void_p()
        if devices is not None:
            arr = (C.c_int * len(devices))(*[int(d) for d in devices])
            rc = L.rcmCreateMulti(lib_path.encode(), arr, len(devices), C.byref(h))
        else:
            rc = L.rcmCreate(lib_path.encode(), device, C.byref(h))
        if rc != 0:
            raise capi.RtError(rc, L.rcmLastError(None).decode())
        object.__setattr__(self, "_L", L)
        object.__setattr__(self, "_h", h)
        object.__setattr__(self, "_lib", capi.RtLib(lib_path))
        object.__setattr__(self, "_size", (0, 0))
        object.__setattr__(self, "_floats", {})

    # -- field access by the reference's names
    def __setattr__(self, name, value):
        if name in self._INT_FIELDS:
            if name == "bvhQuality" and isinstance(value, str):
                value = QUALITY[value]
            self._ck(self._L.rcmSetInt(self._h, name.encode(), int(value)))
        elif name in self._FLOAT_FIELDS:
            self._floats[name] = float(value)
            self._ck(self._L.rcmSetFloat(self._h, name.encode(), float(value)))
        else:
            object.__setattr__(self, name, value)

    def __getattr__(self, name):
        if name in RayComputeManager._INT_FIELDS:
            v = C.c_int()
            self._ck(self._L.rcmGetInt(self._h, name.encode(), C.byref(v)))
            return v.value
        if name in RayComputeManager._FLOAT_FIELDS:
            return self._floats.get(name)
        raise AttributeError(name)

    def _ck(self, rc: int):
        if rc != 0:
            raise capi.RtError(rc, self._L.rcmLastError(self._h).decode())

    # -- scene set-up (what Unity's scene graph provides in the reference)
    def set_screen(self, width: int, height: int):
        self._ck(self._L.rcmSetScreen(self._h, width, height))
        object.__setattr__(self, "_size", (width, height))

    def set_camera(self, field_of_view: float, local_to_world):
        m = column_major(local_to_world)
        self._ck(self._L.rcmSetCamera(self._h, float(field_of_view), m.ctypes.data))

    def set_sun(self, color=(1, 1, 1, 1), forward: Optional[Sequence[float]] = None):
        c = np.asarray(color, dtype=np.float32)
        f = np.asarray(forward, dtype=np.float32) if forward is not None else None
        self._ck(self._L.rcmSetSun(self._h, c.ctypes.data, f.ctypes.data if f is not None else None))

    def add_mesh(self, vertices, indices, normals) -> int:
        v = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 3)
        n = np.ascontiguousarray(normals, dtype=np.float32).reshape(-1, 3)
        idx = np.ascontiguousarray(indices, dtype=np.int32).reshape(-1)
        rc = self._L.rcmAddMesh(self._h, v.ctypes.data, v.shape[0], idx.ctypes.data, idx.size, n.ctypes.data)
        if rc < 0:
            self._ck(rc)
        return rc

    def add_model(self, mesh_id: int, local_to_world, world_to_local, material: np.ndarray) -> int:
        a, b = column_major(local_to_world), column_major(world_to_local)
        mat = np.ascontiguousarray(material, dtype=MATERIAL_DTYPE).reshape(1)
        rc = self._L.rcmAddModel(self._h, mesh_id, a.ctypes.data, b.ctypes.data, mat.ctypes.data)
        if rc < 0:
            self._ck(rc)
        return rc

    def set_model_transform(self, model: int, local_to_world, world_to_local):
        a, b = column_major(local_to_world), column_major(world_to_local)
        self._ck(self._L.rcmSetModelTransform(self._h, model, a.ctypes.data, b.ctypes.data))

    def set_model_material(self, model: int, material: np.ndarray):
        mat = np.ascontiguousarray(material, dtype=MATERIAL_DTYPE).reshape(1)
        self._ck(self._L.rcmSetModelMaterial(self._h, model, mat.ctypes.data))

    def set_spheres(self, spheres: np.ndarray):
        s = np.ascontiguousarray(spheres, dtype=SPHERE_DTYPE)
        self._ck(self._L.rcmSetSpheres(self._h, s.ctypes.data if s.size else None, int(s.shape[0])))

    # -- the reference's methods
    def OnEnable(self):
        self._ck(self._L.rcmOnEnable(self._h))

    def ResetAccumulatedRender(self):
        self._ck(self._L.rcmResetAccumulatedRender(self._h))

    def RenderFrame(self):
        self._ck(self._L.rcmRenderFrame(self._h))

    Update = RenderFrame

    def OnDestroy(self):
        if self._h:
            self._L.rcmDestroy(self._h)
            object.__setattr__(self, "_h", None)

    def __del__(self):
        try:
            self.OnDestroy()
        except Exception:
            pass

    # -- render textures
    def _read(self, fn, out):
        w, h = self._size
        if out is None:
            out = np.empty((h, w, 4), dtype=np.float32)
        self._ck(fn(self._h, out.ctypes.data, out.nbytes))
        return out

    @property
    def raytraceFrameTex(self) -> np.ndarray:
        return self._read(self._L.rcmReadFrame, None)

    @property
    def accumulatedResult(self) -> np.ndarray:
        return self._read(self._L.rcmReadAccumulated, None)

    def read_accumulated_into(self, ptr: int, nbytes: int):
        self._ck(self._L.rcmReadAccumulated(self._h, C.c_void_p(ptr), nbytes))

    def bvh_stats(self) -> list[dict]:
        buf = np.zeros(9 * 64, dtype=np.int32)
        n = self._L.rcmGetBVHStats(self._h, buf.ctypes.data, 64)
        return [dict(zip(STAT_NAMES, (int(x) for x in buf[9 * i:9 * i + 9]))) for i in range(min(n, 64))]

    # -- the underlying C-ABI context (extensions: options, stats, tiles, streams)
    @property
    def context(self) -> capi.RtContext:
        ctx = capi.RtContext(self._lib, C.c_void_p(self._L.rcmContext(self._h)))
        ctx.destroy = lambda: None          # owned by the C++ manager
        ctx.width, ctx.height = self._size
        return ctx
