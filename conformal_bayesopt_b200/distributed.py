"""Candidate sharding across the GPUs of one box.

The path has no data-path collective: every candidate's acquisition value is independent given the
(replicated) model state and the shared random inputs.  Ranks score contiguous candidate blocks and
exchange either the scores (`all_gather`, b/P doubles per rank) or just (max, argmax) pairs (16
bytes per rank).  Works with any callable `acq(X) -> (b,)`, so the host logic is testable on CPU
with the gloo backend.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(b: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of rank `rank`; blocks differ by at most one candidate."""
    base, rem = divmod(b, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_shared_inputs(tensors, src: int = 0, group=None):
    """Make the injected random inputs (eps_grid, z, eps_mc) identical on every rank: rank `src` draws, all receive."""
    if dist.is_available() and dist.is_initialized():
        for t in tensors:
            dist.broadcast(t, src=src, group=group)
    return tensors


NEUTRAL_STATS = [float("inf"), float("-inf"), 0.0, 0.0]     # {min, max, sum, count} of an empty shard


def _collective_device(group, device):
    """NCCL process groups reduce CUDA tensors only: default to the current CUDA device for them, CPU otherwise."""
    if device is not None:
        return device
    if dist.get_backend(group) == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def refine_reducer(group=None, device=None):
    """Hook for `ops.conformal_acq(..., refine_reduce=...)` when the candidate batch is sharded and
    max_grid_refinements > 0: one tiny all-reduce of {min, max, sum, count} per refinement pass makes the rescale /
    best-pass decisions batch-global, as in the reference's single batch (helpers.py:151-160,177).
    The returned callable counts its calls (`reduce.calls`), so that a rank with an EMPTY shard can take part in exactly
    the reductions the other ranks make (`sharded_acquisition`)."""
    def reduce(stats):
        reduce.calls += 1
        if not (dist.is_available() and dist.is_initialized()):
            return stats
        t = torch.tensor(stats, dtype=torch.float64, device=_collective_device(group, device))
        mn, mx, sm = t[0:1].clone(), t[1:2].clone(), t[2:4].clone()
        dist.all_reduce(mn, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
        dist.all_reduce(sm, op=dist.ReduceOp.SUM, group=group)
        return [float(mn), float(mx), float(sm[0]), float(sm[1])]
    reduce.calls = 0
    return reduce


def refine_idle_rank(reduce, max_grid_refinements: int, grid_res: int):
    """What a rank WITHOUT candidates does while the others run the refinement loop: it contributes neutral statistics
    to every reduction and follows the same break rule (cbo_api.cu / helpers.py:177), so that no rank is left waiting in
    an all-reduce."""
    for _ in range(max_grid_refinements + 1):
        mn, mx, _, _ = reduce(list(NEUTRAL_STATS))
        if not (mn < 2.0 or mx > grid_res - 2):
            break


def sharded_acquisition(acq: Callable[[torch.Tensor], torch.Tensor], X: torch.Tensor, group=None) -> torch.Tensor:
    """Every rank holds the full candidate batch X (b, q, d); returns the full (b,) score vector on every rank.
    A rank whose shard is empty (b < world size) still joins the refinement all-reduces of an acquisition object that
    has `refine_reduce` installed."""
    if not (dist.is_available() and dist.is_initialized()):
        return acq(X)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    b = X.shape[0]
    lo, hi = shard_bounds(b, rank, world)
    if hi > lo:
        local = acq(X[lo:hi])
    else:
        local = X.new_empty(0)
        reducer, R = getattr(acq, "refine_reduce", None), int(getattr(acq, "max_grid_refinements", 0) or 0)
        if reducer is not None and R > 0:
            refine_idle_rank(reducer, R, int(acq.grid_res))
    sizes = [shard_bounds(b, r, world) for r in range(world)]
    width = max(h - l for l, h in sizes)
    buf = local.new_zeros(width)
    buf[: hi - lo] = local
    gathered = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(gathered, buf, group=group)
    return torch.cat([g[: h - l] for g, (l, h) in zip(gathered, sizes)])


def sharded_argmax(acq: Callable[[torch.Tensor], torch.Tensor], X_local: torch.Tensor, offset: int, group=None
                   ) -> Tuple[float, int]:
    """Each rank scores ITS OWN candidates `X_local` (global index of the first one = `offset`); returns the global
    (best value, best index) with lowest-index tie-break, exchanging 16 bytes per rank
    (matches `all_acq_vals.argmax()`, conformalbo/utils.py:283)."""
    vals = acq(X_local)
    best = torch.max(vals, dim=0)
    pair = torch.stack([best.values.double(), (best.indices + offset).double()])
    if not (dist.is_available() and dist.is_initialized()):
        return float(pair[0]), int(pair[1])
    world = dist.get_world_size(group)
    gathered = [torch.empty_like(pair) for _ in range(world)]
    dist.all_gather(gathered, pair, group=group)
    allp = torch.stack(gathered)                                   # (world, 2)
    vmax = allp[:, 0].max()
    idx = allp[allp[:, 0] == vmax, 1].min()                        # lowest global index among ties
    return float(vmax), int(idx)
