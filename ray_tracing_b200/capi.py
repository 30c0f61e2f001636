"""ctypes binding of the C-ABI declared in include/rt_b200.h.

`RtLib()` loads the product library, librt_b200.so (CUDA, sm_100a).  There is no CPU fallback: if the
library is missing it raises, and `RtLib.create()` raises when no B200-class GPU is present.  Tests and
bench.py may pass another path implementing the same ABI (the CPU oracle) — the product never does.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(PKG_DIR, "librt_b200.so")

RT_OK = 0
RT_E_INVALID, RT_E_UNKNOWN_NAME, RT_E_NO_DEVICE, RT_E_CUDA, RT_E_STATE = -1, -2, -3, -4, -5
RT_KERNEL_RAYTRACE, RT_KERNEL_RESET_ACCUMULATED = 0, 1

# numpy dtypes of the blittable structs in include/rt_types.h
MATERIAL_DTYPE = np.dtype([
    ("diffuseCol", "<f4", 4), ("emissionCol", "<f4", 4), ("specularCol", "<f4", 4), ("absorption", "<f4", 4),
    ("absorptionStrength", "<f4"), ("emissionStrength", "<f4"), ("smoothness", "<f4"),
    ("specularProbability", "<f4"), ("ior", "<f4"), ("flag", "<i4")])
NODE_DTYPE = np.dtype([("boundsMin", "<f4", 3), ("boundsMax", "<f4", 3), ("startIndex", "<i4"), ("triangleCount", "<i4")])
TRIANGLE_DTYPE = np.dtype([("posA", "<f4", 3), ("posB", "<f4", 3), ("posC", "<f4", 3),
                           ("normA", "<f4", 3), ("normB", "<f4", 3), ("normC", "<f4", 3)])
MODEL_DTYPE = np.dtype([("nodeOffset", "<i4"), ("triOffset", "<i4"), ("worldToLocal", "<f4", 16),
                        ("localToWorld", "<f4", 16), ("material", MATERIAL_DTYPE)])
SPHERE_DTYPE = np.dtype([("centre", "<f4", 3), ("radius", "<f4"), ("material", MATERIAL_DTYPE)])
assert MATERIAL_DTYPE.itemsize == 88 and NODE_DTYPE.itemsize == 32 and TRIANGLE_DTYPE.itemsize == 72
assert MODEL_DTYPE.itemsize == 224 and SPHERE_DTYPE.itemsize == 104

BUFFER_DTYPES = {"Triangles": TRIANGLE_DTYPE, "Nodes": NODE_DTYPE, "ModelInfo": MODEL_DTYPE, "Spheres": SPHERE_DTYPE}


class RtStats(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("boxTests", C.c_uint64), ("triTests", C.c_uint64), ("sphereTests", C.c_uint64),
                ("dispatches", C.c_uint64), ("kernelMs", C.c_double), ("sphereBoxTests", C.c_uint64), ("exchangeMs", C.c_double)]


class RtError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"rt error {code}: {msg}")
        self.code = code


class RtLib:
    """A loaded implementation of the rt_b200.h ABI."""

    def __init__(self, path: Optional[str] = None):
        self.path = path or DEFAULT_LIB
        if not os.path.exists(self.path):
            raise FileNotFoundError(
                f"{self.path} not found — build it with `python -m ray_tracing_b200.build` "
                "(librt_b200 is CUDA-only; there is no CPU fallback)")
        L = self.lib = C.CDLL(self.path)
        vp, cp, ci, cf = C.c_void_p, C.c_char_p, C.c_int, C.c_float
        sig = {
            "rtGetVersion": ([], ci),
            "rtCreate": ([C.POINTER(vp), ci], ci),
            "rtDestroy": ([vp], ci),
            "rtLastError": ([vp], cp),
            "rtSetBuffer": ([vp, cp, vp, ci, ci], ci),
            "rtSetInt": ([vp, cp, ci], ci),
            "rtSetInts": ([vp, cp, C.POINTER(ci), ci], ci),
            "rtSetFloat": ([vp, cp, cf], ci),
            "rtSetVector": ([vp, cp, C.POINTER(cf)], ci),
            "rtSetMatrix": ([vp, cp, C.POINTER(cf)], ci),
            "rtSetBool": ([vp, cp, ci], ci),
            "rtResize": ([vp, ci, ci], ci),
            "rtDispatch": ([vp, ci, ci, ci, ci], ci),
            "rtReadback": ([vp, cp, vp, C.c_size_t], ci),
            "rtSynchronize": ([vp], ci),
            "rtDisplay": ([vp, ci, ci, vp, C.c_size_t], ci),
            "rtSetStream": ([vp, vp], ci),
            "rtSetTile": ([vp, ci, ci, ci], ci),
            "rtPackTile": ([vp], ci),
            "rtUnpackTiles": ([vp], ci),
            "rtGetDevicePointer": ([vp, cp, C.POINTER(vp), C.POINTER(C.c_size_t)], ci),
            "rtSetOption": ([vp, cp, ci], ci),
            "rtGetIpcHandles": ([vp, vp, C.c_size_t], ci),
            "rtSetPeers": ([vp, ci, vp, C.c_size_t], ci),
            "rtGetStats": ([vp, C.POINTER(RtStats)], ci),
            "rtResetStats": ([vp], ci),
            "rtBuildBVH": ([vp, vp, ci, vp, ci, vp, ci, vp, vp, ci, C.POINTER(ci)], ci),
            "rtCreateMulti": ([C.POINTER(vp), C.POINTER(ci), ci], ci),
            "rtGetUniqueId": ([vp, C.c_size_t], ci),
            "rtCommInit": ([vp, vp, C.c_size_t, ci, ci], ci),
            "rtCommDestroy": ([vp], ci),
            "rtExchangeTiles": ([vp], ci),
            "rtReadbackAsync": ([vp, cp, vp, C.c_size_t], ci),
            "rtDisplayAsync": ([vp, ci, ci, vp, C.c_size_t], ci),
            "rtReadbackWait": ([vp], ci),
        }
        for name, (args, res) in sig.items():
            if not hasattr(L, name):
                if name in ("rtGetIpcHandles", "rtSetPeers", "rtBuildBVH", "rtCreateMulti", "rtGetUniqueId", "rtCommInit", "rtCommDestroy", "rtExchangeTiles", "rtReadbackAsync", "rtDisplayAsync", "rtReadbackWait"):      # absent from older experimental builds used in A/B runs
                    continue
                raise AttributeError(f"{self.path} does not export {name}")
            fn = getattr(L, name)
            fn.argtypes, fn.restype = args, res

    def create(self, device: int = 0) -> "RtContext":
        h = C.c_void_p()
        rc = self.lib.rtCreate(C.byref(h), device)
        if rc != RT_OK:
            raise RtError(rc, (self.lib.rtLastError(None) or b"").decode())
        return RtContext(self, h)

    def create_multi(self, devices) -> "RtContext":
        """rtCreateMulti: ONE context that renders on all `devices` of this process (row-band tiles + the all-gather inside rtDispatch)."""
        devices = list(devices)
        arr = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        rc = self.lib.rtCreateMulti(C.byref(h), arr, len(devices))
        if rc != RT_OK:
            raise RtError(rc, (self.lib.rtLastError(None) or b"").decode())
        return RtContext(self, h)

    def unique_id(self) -> bytes:
        """rtGetUniqueId: the 128 bytes rank 0 hands to every rank for rtCommInit."""
        buf = C.create_string_buffer(128)
        rc = self.lib.rtGetUniqueId(buf, 128)
        if rc != RT_OK:
            raise RtError(rc, (self.lib.rtLastError(None) or b"").decode())
        return buf.raw


class RtContext:
    def __init__(self, lib: RtLib, handle: C.c_void_p):
        self._lib, self._L, self._h = lib, lib.lib, handle
        self.width = self.height = 0

    # -- plumbing
    def _ck(self, rc: int):
        if rc != RT_OK:
            raise RtError(rc, (self._L.rtLastError(self._h) or b"").decode())

    @property
    def handle(self) -> C.c_void_p:
        return self._h

    def destroy(self):
        if self._h:
            self._L.rtDestroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    # -- ComputeShader-like surface
    def set_buffer(self, name: str, data: np.ndarray):
        dt = BUFFER_DTYPES.get(name)
        if dt is not None and data.dtype != dt:
            raise TypeError(f"{name} expects dtype {dt}")
        a = np.ascontiguousarray(data)
        self._ck(self._L.rtSetBuffer(self._h, name.encode(), a.ctypes.data_as(C.c_void_p) if a.size else None,
                                     int(a.shape[0]) if a.ndim else 0, a.dtype.itemsize))

    def set_buffer_raw(self, name: str, ptr, count: int, stride: int):
        self._ck(self._L.rtSetBuffer(self._h, name.encode(), ptr, count, stride))

    def set_int(self, name: str, v: int):
        self._ck(self._L.rtSetInt(self._h, name.encode(), int(v)))

    def set_ints(self, name: str, vals):
        arr = (C.c_int * len(vals))(*[int(v) for v in vals])
        self._ck(self._L.rtSetInts(self._h, name.encode(), arr, len(vals)))

    def set_float(self, name: str, v: float):
        self._ck(self._L.rtSetFloat(self._h, name.encode(), float(v)))

    def set_vector(self, name: str, v):
        vv = list(v) + [0.0] * (4 - len(v))
        arr = (C.c_float * 4)(*[float(x) for x in vv])
        self._ck(self._L.rtSetVector(self._h, name.encode(), arr))

    def set_matrix(self, name: str, m):
        """m: 4x4 array in ordinary (row, col) indexing; sent column-major like Unity's Matrix4x4."""
        a = np.asarray(m, dtype=np.float32).reshape(4, 4)
        flat = np.ascontiguousarray(a.T).reshape(16)
        arr = (C.c_float * 16)(*[float(x) for x in flat])
        self._ck(self._L.rtSetMatrix(self._h, name.encode(), arr))

    def set_bool(self, name: str, v: bool):
        self._ck(self._L.rtSetBool(self._h, name.encode(), 1 if v else 0))

    def resize(self, w: int, h: int):
        self._ck(self._L.rtResize(self._h, int(w), int(h)))
        self.width, self.height = int(w), int(h)

    def dispatch(self, kernel: int, gx: int, gy: int, gz: int = 1):
        self._ck(self._L.rtDispatch(self._h, kernel, gx, gy, gz))

    def dispatch_full(self, kernel: int = RT_KERNEL_RAYTRACE):
        """ComputeHelper.Dispatch(cs, width, height, kernelIndex) (ComputeHelper.cs:25-32)."""
        self.dispatch(kernel, (self.width + 7) // 8, (self.height + 7) // 8, 1)

    def readback(self, tex: str, out: Optional[np.ndarray] = None) -> np.ndarray:
        if out is None:
            out = np.empty((self.height, self.width, 4), dtype=np.float32)
        self._ck(self._L.rtReadback(self._h, tex.encode(), out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    def readback_into(self, tex: str, ptr: int, nbytes: int):
        self._ck(self._L.rtReadback(self._h, tex.encode(), C.c_void_p(ptr), nbytes))

    def display(self, use_accumulated: bool, frame: int) -> np.ndarray:
        """Display.shader: tex / Frame, encoded as the sRGB 8-bit back buffer.  Returns (H, W, 4) uint8, row 0 = bottom."""
        out = np.empty((self.height, self.width, 4), dtype=np.uint8)
        self._ck(self._L.rtDisplay(self._h, 1 if use_accumulated else 0, int(frame), out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    def synchronize(self):
        self._ck(self._L.rtSynchronize(self._h))

    # -- extensions
    def set_stream(self, stream_ptr: int):
        self._ck(self._L.rtSetStream(self._h, C.c_void_p(stream_ptr)))

    def set_tile(self, rank: int, world: int, band_rows: int):
        self._ck(self._L.rtSetTile(self._h, rank, world, band_rows))

    def readback_async(self, tex: str, host_ptr: int, nbytes: int):
        """rtReadbackAsync into caller-owned (pinned) host memory; valid after readback_wait()."""
        self._ck(self._L.rtReadbackAsync(self._h, tex.encode(), C.c_void_p(host_ptr), nbytes))

    def display_async(self, use_accumulated: bool, frame: int, host_ptr: int, nbytes: int):
        self._ck(self._L.rtDisplayAsync(self._h, 1 if use_accumulated else 0, int(frame), C.c_void_p(host_ptr), nbytes))

    def readback_wait(self):
        self._ck(self._L.rtReadbackWait(self._h))

    def comm_init(self, unique_id: bytes, rank: int, world: int):
        """rtCommInit (collective over all ranks): NCCL communicator inside the context; RayTrace dispatches then end with the
        all-gather of the frame's tiles, and the context renders tile `rank` of `world`."""
        assert len(unique_id) == 128
        self._ck(self._L.rtCommInit(self._h, unique_id, 128, rank, world))

    def comm_destroy(self):
        self._ck(self._L.rtCommDestroy(self._h))

    def exchange_tiles(self):
        self._ck(self._L.rtExchangeTiles(self._h))

    def pack_tile(self):
        self._ck(self._L.rtPackTile(self._h))

    def unpack_tiles(self):
        self._ck(self._L.rtUnpackTiles(self._h))

    def device_pointer(self, name: str) -> tuple[int, int]:
        p, n = C.c_void_p(), C.c_size_t()
        self._ck(self._L.rtGetDevicePointer(self._h, name.encode(), C.byref(p), C.byref(n)))
        return int(p.value), int(n.value)

    def ipc_handles(self) -> bytes:
        buf = C.create_string_buffer(144)
        self._ck(self._L.rtGetIpcHandles(self._h, buf, 144))
        return buf.raw

    def set_peers(self, handles: list):
        """handles: one 144-byte blob (ipc_handles()) per peer rank."""
        blob = b"".join(handles)
        self._ck(self._L.rtSetPeers(self._h, len(handles), blob if handles else None, len(blob)))

    def build_bvh(self, vertices, indices, normals, quality: int = 1):
        """rtBuildBVH: BVH(verts, indices, normals, quality) of the reference (BVH.cs:26), built by the backend -> (triangles, nodes)."""
        v = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 3)
        n = np.ascontiguousarray(normals, dtype=np.float32).reshape(-1, 3)
        idx = np.ascontiguousarray(indices, dtype=np.int32).reshape(-1)
        ntri = idx.size // 3
        tris = np.empty(max(ntri, 1), dtype=TRIANGLE_DTYPE)
        nodes = np.empty(2 * max(ntri, 1) + 1, dtype=NODE_DTYPE)
        count = C.c_int()
        self._ck(self._L.rtBuildBVH(self._h, v.ctypes.data, v.shape[0], idx.ctypes.data, idx.size, n.ctypes.data, int(quality),
                                    tris.ctypes.data, nodes.ctypes.data, nodes.shape[0], C.byref(count)))
        return tris[:ntri], nodes[:count.value]

    def set_option(self, name: str, v: int):
        self._ck(self._L.rtSetOption(self._h, name.encode(), int(v)))

    def stats(self) -> dict:
        s = RtStats()
        self._ck(self._L.rtGetStats(self._h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in RtStats._fields_}

    def reset_stats(self):
        self._ck(self._L.rtResetStats(self._h))
