"""The per-box communicator behind the multi-GPU multiply (one process per GPU, one NVSwitch box).

The reference moves tiles with two Spark shuffles and one reduceByKey (matrix/BlockMatrix.scala:161-177).  Here the
whole exchange lives in libmarlin_b200.so (csrc/dist.cu, C ABI `mb_comm_*` / `mb_matmul_blocked_dist`): ranks
rendezvous over POSIX shared memory, pull tiles over NVLink peer memory (CUDA IPC) band by band while the persistent
DMMA kernel already runs, and reduce the k partials inside the GEMM epilogue (fused reduce-scatter) or through staged
adds.  This module only creates the communicator: torch.distributed is used ONCE, to agree on a session name and on
whether every rank could set the communicator up; no torch.distributed call sits on the data path.
"""
from __future__ import annotations

import ctypes as C
import os
import uuid
from typing import Optional

import torch

from . import _native as nat
from .runtime import Runtime, world


def transport() -> str:
    """'p2p' (default on GPUs: the C-ABI engine over NVLink peer memory) or 'nccl' (MARLIN_B200_TRANSPORT=nccl: the
    Python plan with grouped NCCL send/recv — also what the CPU/gloo tests exercise)."""
    return os.environ.get("MARLIN_B200_TRANSPORT", "p2p").lower()


class PeerMesh:
    """Process-wide mb_comm.  `PeerMesh.get()` is collective on first use and returns None on EVERY rank when the
    communicator cannot be created somewhere (no peer access, IPC unavailable): callers then use the NCCL transport."""
    _instance: Optional["PeerMesh"] = None
    _failed = False

    @classmethod
    def get(cls) -> Optional["PeerMesh"]:
        if cls._failed or transport() != "p2p" or not torch.cuda.is_available():
            return None
        if cls._instance is None:
            import torch.distributed as dist
            rank, ws = world()
            box = [uuid.uuid4().hex[:16] if rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            mesh, ok = None, 1
            try:
                mesh = PeerMesh(box[0])
            except Exception:
                ok = 0
            t = torch.tensor([ok], device="cuda", dtype=torch.int32)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            if int(t.item()) == 0:
                cls._failed = True
                return None
            cls._instance = mesh
        return cls._instance

    def __init__(self, session: str):
        self.rt = Runtime.get()
        self.rank, self.ws = world()
        # tiles are torch allocations: they must be exportable through CUDA IPC (not the case with the allocator's
        # expandable_segments / VMM mode).  Probing here makes EVERY rank fall back together instead of one rank failing
        # in the middle of a collective multiply.
        probe = torch.empty(1024, dtype=torch.float64, device=self.rt.device)
        h, off, size = C.create_string_buffer(64), C.c_int64(), C.c_int64()
        nat.check(self.rt.lib.mb_ipc_export(self.rt.ctx, C.c_void_p(probe.data_ptr()), h, C.byref(off), C.byref(size)))
        self.comm = C.c_void_p()
        nat.check(self.rt.lib.mb_comm_init(self.rt.ctx, self.rank, self.ws, session.encode(), C.byref(self.comm)))

    def plan(self, m: int, k: int, n: int):
        """(product seq -> rank, C tile -> owner) exactly as the engine will run it (mb_dist_plan)."""
        pr = (C.c_int32 * (m * k * n))()
        co = (C.c_int32 * (m * n))()
        nat.check(self.rt.lib.mb_dist_plan(m, k, n, self.ws, pr, co))
        return list(pr), list(co)

    def check(self) -> None:
        nat.check(self.rt.lib.mb_comm_check(self.comm))
