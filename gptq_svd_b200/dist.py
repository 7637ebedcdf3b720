"""
Multi-GPU plumbing for the hot path (one process per GPU, torch.distributed over
NCCL / NVLink; gloo on CPU for the host-logic tests).  New design - the reference is
single-process, single-GPU.

Two levels of parallelism, both exact for the hot path:

1. Token sharding of H.  Rows of X are independent summands of X^T X, so each rank
   accumulates its own tokens with the tcgen05 SYRK and ONE all-reduce (sum, fp64 payload:
   134 MB at n=4096, 1.2 GB at n=12288) produces the full Hessian on every rank.  This is
   the only collective on the path - it is a real exchange step.
2. Independent units.  The spectral solves of different groups and the GPTQ loops of
   different Linears do not depend on each other once H is known (and whole decoder layers
   do not depend on each other under synthetic activations), so they are sharded across
   ranks longest-first (LPT) with no data-path collective; a solve's (R, R_x, perm) is
   broadcast to the ranks that own its sibling Linears (q/k/v share one H, gate/up share one,
   reference quantize.py:110-219).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from .gptq_utils import HessianAccumulator, Quantizer, SpectralFactors, gptq_quantize, spectral_solve


# ----------------------------------------------------------------------------- host logic
def shard_range(total: int, world: int, rank: int, multiple: int = 1) -> Tuple[int, int]:
    """[begin, end) of `total` items for `rank`: contiguous, balanced, boundaries on a
    multiple of `multiple` (e.g. whole 2048-token sequences)."""
    units = (total + multiple - 1) // multiple
    base, rem = divmod(units, world)
    b = rank * base + min(rank, rem)
    e = b + base + (1 if rank < rem else 0)
    return min(b * multiple, total), min(e * multiple, total)


def lpt_assign(costs: Sequence[float], world: int) -> List[int]:
    """Longest-processing-time-first assignment of jobs to `world` ranks; returns owner per job.
    Deterministic: ties go to the lower job index, then to the lower rank."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0.0] * world
    owner = [0] * len(costs)
    for i in order:
        r = min(range(world), key=lambda q: (load[q], q))
        owner[i] = r
        load[r] += costs[i]
    return owner


def solve_cost(n: int) -> float:
    return float(n) ** 3            # eigh + QRCP + QR are all O(n^3)


def loop_cost(m: int, n: int) -> float:
    return float(m) * float(n) ** 2


@dataclass
class BlockPlan:
    solve_owner: List[int]                  # per group
    loop_owner: List[List[int]]             # per group, per Linear


def plan_block(groups: Sequence[Tuple[int, Sequence[int]]], world: int) -> BlockPlan:
    """groups = [(in_features, [out_features, ...]), ...].  Solves are placed by LPT on n^3;
    loops by LPT on m n^2 starting from the load the solves already put on each rank."""
    solve_owner = lpt_assign([solve_cost(n) for n, _ in groups], world)
    load = [0.0] * world
    for (n, _), r in zip(groups, solve_owner):
        load[r] += solve_cost(n) * 30.0      # measured: a solve costs ~30x the flops-equivalent of a loop unit
    jobs = [(gi, li, loop_cost(m, n)) for gi, (n, outs) in enumerate(groups) for li, m in enumerate(outs)]
    loop_owner = [[0] * len(outs) for _, outs in groups]
    for gi, li, c in sorted(jobs, key=lambda j: (-j[2], j[0], j[1])):
        r = min(range(world), key=lambda q: (load[q], q))
        loop_owner[gi][li] = r
        load[r] += c
    return BlockPlan(solve_owner, loop_owner)


# ----------------------------------------------------------------------------- collectives
def allreduce_hessian(H: torch.Tensor, n_samples: int, group=None) -> int:
    """In-place sum of the un-normalised H over ranks and of the token counts.  fp64 payload:
    the result is identical on every rank, so k / perm / codes agree across ranks."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return n_samples
    dist.all_reduce(H, op=dist.ReduceOp.SUM, group=group)
    cnt = torch.tensor([n_samples], dtype=torch.int64, device=H.device)
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM, group=group)
    return int(cnt.item())


class ShardedHessianAccumulator(HessianAccumulator):
    """HessianAccumulator whose add_batch sees only this rank's tokens; get_hessian()
    all-reduces once and returns the global H / n_samples."""

    def __init__(self, in_features, device, dtype=torch.float64, group=None):
        super().__init__(in_features, device, dtype)
        self.group = group
        self._reduced = False

    def get_hessian(self):
        if not self._reduced:
            self.n_samples = allreduce_hessian(self.H, self.n_samples, self.group)
            self._reduced = True
        return super().get_hessian()


def broadcast_factors(f: Optional[SpectralFactors], n: int, src: int, device, group=None) -> SpectralFactors:
    """Send (k, R, R_x, perm, eigvals) from `src` to every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return f
    rank = dist.get_rank(group)
    k_t = torch.tensor([f.k if rank == src else 0], dtype=torch.int64, device=device)
    dist.broadcast(k_t, src, group=group)
    k = int(k_t.item())
    if rank == src:
        R, Rx, perm, eig = f.R.contiguous(), f.R_x.contiguous(), f.perm, f.eigvals
    else:
        R = torch.empty((k, n), dtype=torch.float64, device=device)
        Rx = torch.empty((k, n), dtype=torch.float64, device=device)
        perm = torch.empty(n, dtype=torch.int64, device=device)
        eig = torch.empty(n, dtype=torch.float64, device=device)
    for t in (R, Rx, perm, eig):
        dist.broadcast(t, src, group=group)
    return SpectralFactors(R=R, R_x=Rx, perm=perm, eigvals=eig, k=k)


# ----------------------------------------------------------------------------- one decoder block on N GPUs
def quantize_block_parallel(x_shards: Sequence[Sequence[torch.Tensor]], weights: Sequence[Sequence[torch.Tensor]],
                            groups: Sequence[Tuple[int, Sequence[int]]], bits: int = 4, group_size: int = 128,
                            sym: bool = False, eps: float = 1e-4, block_size: int = 1024,
                            group=None) -> Dict[Tuple[int, int], object]:
    """Token-sharded Hessians + LPT-scheduled solves and loops for one decoder block.

    x_shards[g] : this rank's calibration batches for group g (each (rows, n) or (B, S, n))
    weights[g][l]: the Linear's weight (every rank holds it; only the owner quantises it)
    Returns {(g, l): QuantizedLinear} for the Linears this rank owns."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    dev = weights[0][0].device
    plan = plan_block(groups, world)
    hessians = []
    for gi, (n, _) in enumerate(groups):
        acc = ShardedHessianAccumulator(n, dev, group=group)
        for xb in x_shards[gi]:
            acc.add_batch(xb)
        hessians.append(acc.get_hessian())          # the one collective
    out = {}
    for gi, (n, outs) in enumerate(groups):
        owner = plan.solve_owner[gi]
        f = spectral_solve(hessians[gi], eps, "energy") if rank == owner else None
        needed = any(o != owner for o in plan.loop_owner[gi])
        if needed:
            f = broadcast_factors(f, n, owner, dev, group)
        for li, _m in enumerate(outs):
            if plan.loop_owner[gi][li] == rank:
                q = Quantizer(bits, group_size, sym)
                out[(gi, li)] = gptq_quantize(weights[gi][li], f.R, q, f.perm, block_size, True, f.R_x)
    return out
