"""RayTraceDisplay — mirror of the reference's display component (Assets/Scripts/Tracer/RayTraceDisplay.cs:9-23 +
Display.shader:42-47) and of its screenshot key (RayComputeManager.cs:106-111): the accumulated image divided by the
frame counter, shown / saved as an 8-bit sRGB picture.  The division and encoding run on the GPU (rtDisplay)."""
from __future__ import annotations

import struct
import zlib

import numpy as np


class RayTraceDisplay:
    def __init__(self, raytracer):
        self.raytracer = raytracer                      # a RayComputeManager

    def OnRenderImage(self) -> np.ndarray:
        """(H, W, 4) uint8, top row first (screen order).  Frame = numAccumulatedFrames when accumulating, else 1 —
        including the reference's off-by-one: the counter is one ahead of the number of summed frames."""
        m = self.raytracer
        acc = bool(m.accumulate)
        img = m.context.display(acc, m.numAccumulatedFrames if acc else 1)
        return img[::-1]

    def save_screenshot(self, path: str) -> None:
        write_png(path, self.OnRenderImage()[..., :3])


def write_png(path: str, rgb8: np.ndarray) -> None:
    h, w, _ = rgb8.shape
    raw = b"".join(b"\x00" + np.ascontiguousarray(rgb8[y]).tobytes() for y in range(h))

    def chunk(tag: bytes, data: bytes) -> bytes:
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0))
                + chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))
