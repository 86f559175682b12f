"""Data-parallel plumbing: the path shards by image / sequence with no data-path collective --
every (cond, uncond) pair and every MMU sequence is independent, weights are replicated (2.9 GB bf16 + 0.19 GB VQ).
One process per GPU (torch.distributed, NCCL over NVLink); the only exchange is the final gather of uint8 images
(or of the [B, N] int64 code grids).  The reference has no multi-GPU inference at all (inference_t2i.py:55)."""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist


def env_world() -> Tuple[int, int, int]:
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init(backend: str | None = None, device: torch.device | None = None) -> Tuple[int, int]:
    """Initialise the default process group from the torchrun environment (no-op for a single process)."""
    rank, _, world = env_world()
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, **kw)
    return rank, world


def shard_rows(n_rows: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [begin, end) of the global batch owned by `rank` (CFG pairs stay together because the
    partition is over images, not over cond/uncond rows)."""
    base, rem = divmod(n_rows, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def gather_rows(local: torch.Tensor, world: int) -> torch.Tensor:
    """All-gather equally sized per-rank results (uint8 images [b, H, W, 3] or code grids [b, N]) along dim 0."""
    if world == 1:
        return local
    local = local.contiguous()
    out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    if dist.get_backend() == "nccl":
        dist.all_gather_into_tensor(out, local)
    else:
        parts = list(out.chunk(world, 0))
        dist.all_gather(parts, local)
    return out


def max_over_ranks(value: float, device: torch.device) -> float:
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
