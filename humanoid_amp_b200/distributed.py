"""Multi-GPU plumbing of the AMP path: one process per GPU, envs sharded, one gradient all-reduce.

The reference scales exactly this way through skrl (``train.py:53-58, 184-196``: ``--distributed`` +
``torch.distributed.run``; each rank owns ``num_envs`` envs on ``cuda:{local_rank}``); motion sampling, AMP observations
and the style reward need no communication.  The only collective on the path is skrl's ``Model.reduce_parameters``:
all-reduce(SUM) of the flattened gradients divided by the world size, once per model per mini-batch.  Here the three
models' gradients travel as ONE flat fp32 buffer per call (one NCCL launch over NVLink instead of three).
"""

from __future__ import annotations

from typing import Iterable, Sequence

import torch
import torch.distributed as dist


def shard_envs(total_envs: int, rank: int, world_size: int) -> tuple[int, int]:
    """``[begin, end)`` of the envs owned by ``rank``; the remainder goes to the lowest ranks."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of size {world_size}")
    base, extra = divmod(total_envs, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def reduce_parameters(parameter_groups: Iterable[Sequence[torch.nn.Parameter]], group=None, flat: torch.Tensor | None = None) -> torch.Tensor | None:
    """Average the gradients of every parameter in ``parameter_groups`` (e.g. policy, value, discriminator) over ranks.

    Semantics of skrl ``Model.reduce_parameters`` (missing grads count as zeros; SUM then divide by world size), but one
    flat buffer and one collective for all models.  ``flat`` may be a preallocated buffer of the right size (returned for
    reuse).  No-op when torch.distributed is not initialised or the world has one rank.
    """
    if not dist.is_available() or not dist.is_initialized():
        return flat
    world = dist.get_world_size(group)
    params = [p for grp in parameter_groups for p in grp]
    if not params or world == 1:
        return flat
    total = sum(p.numel() for p in params)
    ref = params[0]
    if flat is None or flat.numel() != total or flat.device != ref.device:
        flat = torch.empty(total, dtype=torch.float32, device=ref.device)
    offset = 0
    for p in params:
        n = p.numel()
        if p.grad is None:
            flat[offset : offset + n].zero_()
        else:
            flat[offset : offset + n].copy_(p.grad.reshape(-1))
        offset += n
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(world)
    offset = 0
    for p in params:
        n = p.numel()
        if p.grad is not None:
            p.grad.copy_(flat[offset : offset + n].view_as(p.grad))
        offset += n
    return flat
