"""Data-parallel plumbing: one process per GPU over NCCL/NVLink (mirrors virtex/utils/distributed.py:82-160, with
torchrun-style environment rendezvous instead of mp.spawn + tcp://)."""
import datetime
import os

import torch
import torch.distributed as dist


def init_from_env(backend: str = None):
    """Join the process group described by RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (no-op for a single process)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        timeout = datetime.timedelta(seconds=int(os.environ.get("VTX_DIST_TIMEOUT_S", "180")))
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local), timeout=timeout)
        else:
            dist.init_process_group(backend, timeout=timeout)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local)
    return rank, world, local


def get_world_size() -> int:
    return dist.get_world_size() if dist.is_initialized() else 1


def get_rank() -> int:
    return dist.get_rank() if dist.is_initialized() else 0


def is_master_process() -> bool:
    return get_rank() == 0


def synchronize() -> None:
    if dist.is_initialized():
        dist.barrier()


def average_across_processes(t):
    """All-reduce (SUM) then divide by the world size; tensors or dicts of tensors (virtex/utils/distributed.py:140-160)."""
    if not dist.is_initialized():
        return t
    if isinstance(t, torch.Tensor):
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        t /= get_world_size()
    elif isinstance(t, dict):
        for k in t:
            dist.all_reduce(t[k], op=dist.ReduceOp.SUM)
            t[k] /= dist.get_world_size()
    return t


def shard_batch_size(global_batch: int, world: int) -> int:
    """Per-rank batch: OPTIM.BATCH_SIZE // world_size (scripts/pretrain_virtex.py:79)."""
    return global_batch // world
