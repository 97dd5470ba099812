"""Multi-GPU plumbing for the path-tracing hot path: one process per GPU, scene replicated, the
samples of every pixel sliced across ranks, ONE sum-reduce of the per-GPU accumulation buffers (NCCL over
NVLink/NVSwitch on GPUs; gloo in the CPU tests). There is no exchange during a pass.

The reference's analogue is the mutex-protected framebuffer its worker goroutines write tiles into
(rt/bucket_renderer.go:291-300)."""
from __future__ import annotations

import os
from typing import Tuple


def slice_samples(spp: int, rank: int, world: int) -> Tuple[int, int]:
    """Samples [base, base+count) of every pixel rendered by `rank`. Slices are disjoint, cover [0, spp) and differ by
    at most one sample; with spp < world the trailing ranks get an empty slice (the reference's 1-spp preview pass)."""
    if world <= 0 or not (0 <= rank < world) or spp < 0:
        raise ValueError(f"bad slice request spp={spp} rank={rank} world={world}")
    q, r = divmod(spp, world)
    count = q + (1 if rank < r else 0)
    base = rank * q + min(rank, r)
    return base, count


class _DevicePtr:
    """Exposes a raw device allocation through __cuda_array_interface__ so torch can view it without a copy."""

    def __init__(self, ptr: int, n_floats: int):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (ptr, False), "version": 3, "strides": None}


def accum_tensor(ctx, device_index: int):
    """torch view (no copy) of the context's accumulation buffer: float32[4*W*H] = (sum R, sum G, sum B, n) per pixel."""
    import torch
    ptr, _, n = ctx.accum_device_ptr()
    return torch.as_tensor(_DevicePtr(ptr, n), device=torch.device("cuda", device_index))


def reduce_sum(tensor, dst: int = 0):
    """The one collective of the path: sum the accumulation buffers onto `dst`."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(tensor, dst=dst, op=dist.ReduceOp.SUM)
    return tensor


def init_from_env(backend: str):
    """torch.distributed rendezvous from RANK / WORLD_SIZE / MASTER_* (torchrun). Returns (rank, world, local_rank)."""
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29513")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local
