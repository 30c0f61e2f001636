"""Row-band tiling of one frame over several GPUs.

The image is cut into bands of `band_rows` rows; band b belongs to rank b % world.  Every rank holds the whole
scene, traces only its bands (pixels keep their GLOBAL index, so seeds and therefore results equal the 1-GPU
run), and after each frame ONE all-gather over NCCL/NVLink hands every rank the finished tiles of all ranks;
a local scatter puts them back at their global position.

`TiledRenderer` is one rank of the one-process-per-GPU form.  The exchange lives INSIDE the C-ABI: rank 0 draws an
id (rtGetUniqueId), the launcher's rendezvous carries its 128 bytes to the other ranks (here: torch.distributed's
store; any transport does), every rank calls rtCommInit, and from then on each RayTrace dispatch ends with
pack -> ncclAllGather -> unpack on the context's stream — no Python, no torch in the data plane.  exchange="torch"
keeps round 1's form (rtPackTile -> torch.distributed all_gather_into_tensor on views of the context's device
buffers -> rtUnpackTiles) for A/B runs.  The host-side helpers (owned_rows, pack_rows, unpack_rows) state the tile
layout in numpy; the world_size-2 gloo test uses them with CPU tensors.
"""
from __future__ import annotations

import numpy as np


def owned_rows(height: int, rank: int, world: int, band_rows: int) -> np.ndarray:
    """Global y of every row rank owns, ascending (band b -> rank b % world)."""
    y = np.arange(height)
    return y[(y // band_rows) % world == rank]


def rows_per_rank(height: int, world: int, band_rows: int) -> int:
    """Rows in the all-gather contribution of every rank = the largest share (rank 0's); others pad."""
    return int(owned_rows(height, 0, world, band_rows).size)


def pack_rows(frame: np.ndarray, accum: np.ndarray, rank: int, world: int, band_rows: int) -> np.ndarray:
    """[frame bands | accumulated bands] of one rank, padded to rows_per_rank — the TileSend layout."""
    h, w = frame.shape[:2]
    n = rows_per_rank(h, world, band_rows)
    out = np.zeros((2, n, w, 4), dtype=np.float32)
    ys = owned_rows(h, rank, world, band_rows)
    out[0, :ys.size] = frame[ys]
    out[1, :ys.size] = accum[ys]
    return out


def unpack_rows(gathered: np.ndarray, height: int, world: int, band_rows: int):
    """Inverse of pack_rows over the rank-major all-gather result (world, 2, rows_per_rank, W, 4)."""
    w = gathered.shape[3]
    frame = np.zeros((height, w, 4), dtype=np.float32)
    accum = np.zeros((height, w, 4), dtype=np.float32)
    for r in range(world):
        ys = owned_rows(height, r, world, band_rows)
        frame[ys] = gathered[r, 0, :ys.size]
        accum[ys] = gathered[r, 1, :ys.size]
    return frame, accum


class _DevView:
    """Minimal __cuda_array_interface__ holder so torch can view a raw device pointer without copying."""

    def __init__(self, ptr: int, nfloats: int):
        self.__cuda_array_interface__ = {"shape": (nfloats,), "typestr": "<f4", "data": (ptr, False), "version": 3, "strides": None}


class TiledRenderer:
    """One rank of a row-tiled render: manager + per-frame all-gather of finished tiles (NCCL)."""

    def __init__(self, mgr, rank: int, world: int, band_rows: int = 8, device=None, fused: bool = False, exchange: str = "abi"):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.mgr, self.rank, self.world, self.band_rows = mgr, rank, world, band_rows
        self.ctx = mgr.context
        self.device = device
        self.exchange = "fused" if (fused and world > 1) else exchange
        if world > 1 and self.exchange == "abi":
            box = [mgr._lib.unique_id() if rank == 0 else None]
            dist.broadcast_object_list(box, src=0)                  # 128 bytes through the launcher's rendezvous
            self.ctx.comm_init(box[0], rank, world)                 # collective: NCCL communicator owned by the context
        self.ctx.set_tile(rank, world, band_rows)
        if world > 1 and self.exchange != "abi":
            self.ctx.set_option("exchange", 0)
        self._send = self._recv = None
        # one torch stream carries everything: trace -> pack -> all-gather -> unpack are ordered on it
        self.stream = torch.cuda.Stream(device=device)
        self.ctx.set_stream(self.stream.cuda_stream)
        # fused = the trace kernel stores finished pixels straight into the peers' images (CUDA-IPC over NVLink);
        # the per-frame collective shrinks to a 4-byte all-reduce used as a stream-ordered barrier
        self.fused = bool(fused) and world > 1
        self._peers_ready = False
        self._token = torch.zeros(1, dtype=torch.int32, device=device) if self.fused else None

    def frame_fence(self):
        """4-byte all-reduce on the renderer's stream = stream-ordered barrier across ranks."""
        self.dist.all_reduce(self._token)

    def _connect_peers(self):
        """Exchange IPC handles of the two images (after the manager has created them) and map the peers'."""
        mine = self.ctx.ipc_handles()
        gathered = [None] * self.world
        self.dist.all_gather_object(gathered, mine)
        self.ctx.set_peers([h for r, h in enumerate(gathered) if r != self.rank])
        self._peers_ready = True

    def _views(self):
        if self._send is None:
            torch = self.torch
            ps, ns = self.ctx.device_pointer("TileSend")
            pr, nr = self.ctx.device_pointer("TileRecv")
            self._send = torch.as_tensor(_DevView(ps, ns // 4), device=self.device)
            self._recv = torch.as_tensor(_DevView(pr, nr // 4), device=self.device)
        return self._send, self._recv

    def render_frame(self):
        """RenderFrame on this rank's bands, then the single all-gather of the frame's tiles."""
        with self.torch.cuda.stream(self.stream):
            if self.fused:
                if not self._peers_ready:
                    self._connect_peers()
                    self.dist.barrier()
                self.frame_fence()                                  # every rank has consumed the previous frame
                self.mgr.RenderFrame()                              # pixels land in every rank's images as they finish
                self.frame_fence()                                  # every rank's kernel is done: the images are complete
                return
            self.mgr.RenderFrame()                                  # exchange "abi": the dispatch itself ends with the all-gather
            if self.world == 1 or self.exchange == "abi":
                return
            send, recv = self._views()
            self.ctx.pack_tile()
            self.dist.all_gather_into_tensor(recv, send)
            self.ctx.unpack_tiles()
