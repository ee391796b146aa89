import os
import time


def start():
    """`new SparkContext(conf)`: one process per GPU; under torchrun join the NCCL group."""
    import torch
    import marlin_b200 as mb
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    if ws > 1:
        import torch.distributed as dist
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    mb.Runtime.get()
    return mb, int(os.environ.get("RANK", "0"))


def stop():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()


def millis() -> float:
    """System.currentTimeMillis() with the device drained first (the reference's clock brackets lazy RDD actions)."""
    import torch
    torch.cuda.synchronize()
    return time.time() * 1000.0
