import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


_MULTI_PROCESS_LAST = ("test_gpu_dist_cabi.py", "test_gpu_multi.py")


def pytest_collection_modifyitems(config, items):
    # the multi-process GPU suites (subprocess ranks, minutes each; on a one-GPU box the ranks time-slice one device) run
    # after everything else, so the single-process kernel parity tests are never queued behind them
    # ... and before them, after every other single-process test, the check of WHICH generator kernel ran
    def rank(it):
        if it.name.startswith("test_compiled_host_mirror_multiplies_across_two_ranks"):
            return 3                              # new this round and never run on hardware: last of all
        if Path(str(it.fspath)).name in _MULTI_PROCESS_LAST:
            return 2
        return 1 if it.name.startswith("test_fill_uniform_fast_kernel_is_the_one_that_ran") else 0
    items.sort(key=rank)
    # gpu tests fail loudly on a box without a GPU only if explicitly selected; otherwise they are deselected by -m "not gpu"
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container (GPU tests run under gpurun)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import reference_model as rm
    rm.build()
    return rm


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
