"""NVLink peer-memory transport for the multi-GPU multiply (one process per GPU, one NVSwitch box).

The reference moves tiles with two Spark shuffles and one reduceByKey (matrix/BlockMatrix.scala:161-177).  Here every
rank maps the other ranks' tile buffers once (CUDA IPC), PULLS the tiles it needs with copy-engine DMA on a side
stream — so transfers overlap the GEMMs that do not depend on them — and the GEMM epilogue of a rank that holds a
non-owner partial stores it straight into a staging slot in the owner's HBM (P2P stores); the owner adds it.
Processes are ordered with stream-ordered flags (monotonic epochs) in exported device memory; no NCCL call sits on
the data path.  torch.distributed is used only to swap IPC handles (metadata) when a matrix is first seen.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Tuple

import torch

from . import _native as nat
from .runtime import Runtime, world

CH_READY, CH_DONE, CH_PARTIAL, CH_FREE = 0, 1, 2, 3
NUM_CHANNELS = 4


def transport() -> str:
    """'p2p' (default on GPUs) or 'nccl' (MARLIN_B200_TRANSPORT=nccl, CPU/gloo runs, or IPC unavailable)."""
    return os.environ.get("MARLIN_B200_TRANSPORT", "p2p").lower()


class PeerMesh:
    _instance: Optional["PeerMesh"] = None
    _failed = False

    @classmethod
    def get(cls) -> Optional["PeerMesh"]:
        """Collective on first use.  Returns None (on every rank) when P2P cannot be set up."""
        if cls._failed or transport() != "p2p" or not torch.cuda.is_available():
            return None
        if cls._instance is None:
            import torch.distributed as dist
            mesh, ok = None, 1
            try:
                mesh = PeerMesh()
            except Exception:
                ok = 0
            t = torch.tensor([ok], device="cuda", dtype=torch.int32)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            if int(t.item()) == 0:
                cls._failed = True
                return None
            cls._instance = mesh
        return cls._instance

    def __init__(self):
        import torch.distributed as dist
        self.rt = Runtime.get()
        self.rank, self.ws = world()
        self.epoch = 0
        self.copy_stream = torch.cuda.Stream(device=self.rt.device)
        flags = C.c_void_p()
        nat.check(self.rt.lib.mb_flags_alloc(self.rt.ctx, NUM_CHANNELS * self.ws, C.byref(flags)))
        self.flags_local = flags.value
        rec = self.export_ptr(self.flags_local)
        recs = [None] * self.ws
        dist.all_gather_object(recs, rec)
        self.flags_peer = [self.flags_local if r == self.rank else self.open(recs[r]) for r in range(self.ws)]
        self.staging: Optional[torch.Tensor] = None
        self.staging_peer: List[Optional[int]] = [None] * self.ws
        self.last_partial_epoch: Dict[int, int] = {}       # owner rank -> last epoch in which I wrote into its staging

    # ---- IPC ----
    def export_ptr(self, ptr: int) -> Tuple[bytes, int]:
        h = C.create_string_buffer(64)
        off, size = C.c_int64(), C.c_int64()
        nat.check(self.rt.lib.mb_ipc_export(self.rt.ctx, C.c_void_p(ptr), h, C.byref(off), C.byref(size)))
        return (h.raw, int(off.value))

    def open(self, rec: Tuple[bytes, int]) -> int:
        base = C.c_void_p()
        nat.check(self.rt.lib.mb_ipc_open(self.rt.ctx, rec[0], C.byref(base)))
        return int(base.value) + rec[1]

    # ---- flags (stream-ordered on torch's current stream) ----
    def _flag_addr(self, on_rank: int, channel: int, src: int) -> int:
        return self.flags_peer[on_rank] + 8 * (channel * self.ws + src)

    def signal(self, on_rank: int, channel: int, value: int) -> None:
        """Set flag[channel][me] on `on_rank` to `value` once everything queued so far on the current stream is done."""
        self.rt.sync_stream()
        nat.check(self.rt.lib.mb_flag_signal(self.rt.ctx, C.c_void_p(self._flag_addr(on_rank, channel, self.rank)), value))

    def wait(self, channel: int, src: int, value: int) -> None:
        """Hold the current stream until flag[channel][src] in MY array reaches `value`."""
        self.rt.sync_stream()
        nat.check(self.rt.lib.mb_flag_wait(self.rt.ctx, C.c_void_p(self._flag_addr(self.rank, channel, src)), value))

    def memcpy(self, dst_ptr: int, src_ptr: int, nbytes: int) -> None:
        self.rt.sync_stream()
        nat.check(self.rt.lib.mb_memcpy_async(self.rt.ctx, C.c_void_p(dst_ptr), C.c_void_p(src_ptr), nbytes))

    # ---- staging for partial C tiles (collective growth: every rank passes the same byte count) ----
    def ensure_staging(self, nbytes: int) -> None:
        import torch.distributed as dist
        have = self.staging.numel() if self.staging is not None else 0
        if nbytes <= have:
            return
        torch.cuda.synchronize()
        dist.barrier()
        self.staging = torch.empty(int(nbytes * 1.25) + 256, dtype=torch.uint8, device=self.rt.device)
        recs = [None] * self.ws
        dist.all_gather_object(recs, self.export_ptr(self.staging.data_ptr()))
        self.staging_peer = [self.staging.data_ptr() if r == self.rank else self.open(recs[r]) for r in range(self.ws)]


def tile_directory(mesh: PeerMesh, tag: str, blocks: dict) -> Optional[dict]:
    """Exchange {(tag, row, col): (ipc handle, offset)} for the packed tiles every rank holds.  Collective.
    Returns None on every rank if some tile is a strided view (the NCCL path handles those)."""
    import torch.distributed as dist
    mine, packed = {}, 1
    for (r, c), s in blocks.items():
        if not s.is_packed():
            packed = 0
            break
        mine[(tag, r, c)] = mesh.export_ptr(s.buf.data_ptr())
    gathered = [None] * mesh.ws
    dist.all_gather_object(gathered, (packed, mine))
    if not all(p for p, _ in gathered):
        return None
    out = {}
    for rnk, (_, d) in enumerate(gathered):
        for key, rec in d.items():
            out[key] = (rnk, rec)
    return out
