"""Multi-GPU plumbing.  Environment copies are independent, so the batch shards by copy
with NO collective on the data path; torch.distributed (NCCL on GPUs, gloo in the CPU
tests) only gathers rollout statistics and reduces timings.  Mirrors the role of the
process pool in utilities/Parallel_Experience_Generator.py:28-40 (n workers, results
gathered at the end)."""
from __future__ import annotations

import numpy as np


def shard_range(n_total: int, rank: int, world: int):
    """Contiguous balanced split of n_total environment copies: [lo, hi) of this rank."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_timing(ms: float, steps: int, device=None):
    """(max over ranks of the device time, sum over ranks of env steps)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(ms), int(steps)
    t = torch.tensor([float(ms)], dtype=torch.float64, device=device)
    s = torch.tensor([int(steps)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    return float(t[0]), int(s[0])


def gather_episode_stats(completion, tardiness, energy, device=None):
    """All ranks' per-copy episode objectives concatenated in rank order (equal shard sizes
    are not required)."""
    import torch
    import torch.distributed as dist
    x = torch.stack([torch.as_tensor(np.asarray(v), dtype=torch.int64) for v in (completion, tardiness, energy)], 1)
    if device is not None:
        x = x.to(device)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return x.cpu().numpy()
    world = dist.get_world_size()
    n = torch.tensor([x.shape[0]], dtype=torch.int64, device=x.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    m = int(max(int(s[0]) for s in sizes))
    pad = torch.zeros((m, 3), dtype=torch.int64, device=x.device)
    pad[: x.shape[0]] = x
    parts = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    return torch.cat([p[: int(s[0])] for p, s in zip(parts, sizes)], 0).cpu().numpy()


def allreduce_gradients(parameters, average=True):
    """Data-parallel gradient reduction for the agents' PPO / SAC / DQN updates: every rank
    computes its loss on its own shard of environment copies, the gradients are summed over the
    ranks in ONE flat bucket (NCCL on GPUs, gloo in the CPU tests) and averaged.  The reference
    trains single-process (agents/MPPPO/MPPPO.py:221-250 computes the loss of one episode and
    steps Adam); with the batch sharded over GPUs this call sits between loss.backward() and
    optimizer.step().  Returns the number of gradient elements reduced."""
    import torch
    import torch.distributed as dist
    params = [p for p in parameters if p.grad is not None]
    if not params:
        return 0
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return sum(p.grad.numel() for p in params)
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    if average:
        flat /= dist.get_world_size()
    o = 0
    for p in params:
        n = p.grad.numel()
        p.grad.copy_(flat[o:o + n].view_as(p.grad))
        o += n
    return o
