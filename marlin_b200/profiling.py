"""Tiny CUDA-event phase recorder (the reference only has System.currentTimeMillis() brackets,
examples/BLAS3.scala:30-56).  Disabled by default; bench.py enables it to split a step into
exchange / gemm / reduce time on the launching stream."""
from __future__ import annotations

from contextlib import contextmanager
from typing import Dict, List, Tuple

import torch

_enabled = False
_events: List[Tuple[str, "torch.cuda.Event", "torch.cuda.Event"]] = []


def enable(flag: bool = True) -> None:
    global _enabled
    _enabled = flag
    _events.clear()


@contextmanager
def phase(name: str):
    if not _enabled or not torch.cuda.is_available():
        yield
        return
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    try:
        yield
    finally:
        e1.record()
        _events.append((name, e0, e1))


def collect() -> Dict[str, Tuple[float, int]]:
    """{phase: (total ms, count)}; synchronises the device and clears the log."""
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    out: Dict[str, Tuple[float, int]] = {}
    for name, e0, e1 in _events:
        ms, n = out.get(name, (0.0, 0))
        out[name] = (ms + e0.elapsed_time(e1), n + 1)
    _events.clear()
    return out
