"""Multi-GPU plumbing of the batch evaluator: contiguous point shards, one process per GPU, no collective on
the data path.  The only communication is control-plane: a barrier, a MAX reduction of
the per-rank timings, and (optionally) the host gather of the per-rank outputs.

The partition is the one `saf_eval_host_multi` uses inside one process (csrc/api.cu): rank g of G owns
[g*N/G, (g+1)*N/G).  The reference partitions the same independent points into 256-point Rayon chunks
(src/evaluator/logic/bytecode/execute/drivers/batch.rs:55-74).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np


def shard_bounds(n_points: int, world: int, rank: int) -> Tuple[int, int]:
    """[begin, end) of `rank`'s contiguous shard; shards cover [0, n_points) without overlap."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("rank must be in [0, world)")
    return (n_points * rank) // world, (n_points * (rank + 1)) // world


def shard_columns(cols: Sequence[np.ndarray], n_points: int, world: int, rank: int) -> List[np.ndarray]:
    """This rank's slice of full-length columns (views, no copy).  A column shorter than n_points is a broadcast
    column (eval_batch semantics, scalar.rs:203-207): the shard keeps what overlaps it, or the last element."""
    b, e = shard_bounds(n_points, world, rank)
    out = []
    for c in cols:
        c = np.asarray(c)
        if c.size >= e:
            out.append(c[b:e])
        elif c.size > b:
            out.append(c[b:])       # runs out inside the shard: the evaluator broadcasts its last element
        else:
            out.append(c[-1:])      # exhausted before the shard: a one-element broadcast column
    return out


def max_over_ranks(values: Sequence[float], device=None) -> List[float]:
    """Element-wise MAX over all ranks of a few scalars (timings); identity without a process group."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.cpu()]


def gather_outputs(local: Sequence[np.ndarray], n_points: int) -> Optional[List[np.ndarray]]:
    """Host gather: rank 0 receives every rank's output shard and returns full-length columns (other ranks get
    None).  This is the `host gather` the north star allows; it is not on the timed data path."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [np.asarray(a) for a in local]
    world, rank = dist.get_world_size(), dist.get_rank()
    gathered: List = [None] * world if rank == 0 else None
    dist.gather_object([np.asarray(a) for a in local], gathered, dst=0)
    if rank != 0:
        return None
    full = []
    for k in range(len(local)):
        col = np.empty(n_points, dtype=np.float64)
        for r in range(world):
            b, e = shard_bounds(n_points, world, r)
            col[b:e] = gathered[r][k]
        full.append(col)
    return full


def eval_sharded(evaluate: Callable[[List[np.ndarray], int], List[np.ndarray]], cols: Sequence[np.ndarray],
                 n_points: int, world: int, rank: int) -> List[np.ndarray]:
    """Runs `evaluate(shard_columns, shard_points)` on this rank's shard and returns its local outputs."""
    b, e = shard_bounds(n_points, world, rank)
    return evaluate(shard_columns(cols, n_points, world, rank), e - b)
