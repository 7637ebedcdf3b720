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
    """Seconds of one spectral solve on a B200 (measured, round 2: 0.17 s at n = 4096, ~0.9 s at 12288 - the small
    ones are latency-bound, so the cost grows like n^2 rather than n^3 over the sizes of interest)."""
    return 8.0e-9 * float(n) ** 2 * max(1.0, float(n) / 16384.0)


def loop_cost(m: int, n: int) -> float:
    """Seconds of one gptq_fwrd (measured: 30 ms for 4096 x 12288, 12.7 ms for 12288 x 4096)."""
    return 5.5e-14 * float(m) * float(n) ** 2


def transfer_cost(n: int) -> float:
    """Seconds to hand (R, R_x) - 16 n^2 bytes at k = n - to another GPU over NVLink (~400 GB/s sustained)."""
    return 16.0 * float(n) ** 2 / 400e9


@dataclass
class BlockPlan:
    solve_owner: List[int]                  # per group
    loop_owner: List[List[int]]             # per group, per Linear


def plan_block(groups: Sequence[Tuple[int, Sequence[int]]], world: int) -> BlockPlan:
    """groups = [(in_features, [out_features, ...]), ...].  Solves are placed longest-first (LPT); every rank
    runs its solves widest first, then its loops.  A loop goes to the rank where it FINISHES first: its group's
    solve owner (factors are local) or another rank (factors arrive `transfer_cost` after the solve) - so
    sibling Linears (q/k/v, gate/up: one H, reference quantize.py:110-112) spread over idle GPUs only when
    that actually shortens the block."""
    costs = [solve_cost(n) for n, _ in groups]
    solve_owner = lpt_assign(costs, world)
    avail = [0.0] * world
    ready = [0.0] * len(groups)
    for gi in sorted(range(len(groups)), key=lambda g: (-costs[g], g)):      # execution order on each rank
        r = solve_owner[gi]
        avail[r] += costs[gi]
        ready[gi] = avail[r]
    jobs = [(gi, li, loop_cost(m, n)) for gi, (n, outs) in enumerate(groups) for li, m in enumerate(outs)]
    loop_owner = [[0] * len(outs) for _, outs in groups]
    for gi, li, c in sorted(jobs, key=lambda j: (ready[j[0]], -j[2], j[0], j[1])):
        n = groups[gi][0]

        def finish(q):
            arrive = ready[gi] + (0.0 if q == solve_owner[gi] else transfer_cost(n))
            return max(avail[q], arrive) + c
        r = min(range(world), key=lambda q: (finish(q), q != solve_owner[gi], q))
        loop_owner[gi][li] = r
        avail[r] = finish(r)
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
            self.check()                      # each rank verifies its own partial sum against its own probe
            self.verify = False               # the probe describes the local tokens only
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


def send_factors(f: SpectralFactors, dst: int, group=None) -> None:
    """Point-to-point hand-off of (k, R, R_x, perm, eigvals) to ONE rank (only the two ranks take part, so a
    rank that is still busy with the wide solve never holds up the exchange of the narrow ones)."""
    dev = f.R.device
    dist.send(torch.tensor([f.k], dtype=torch.int64, device=dev), dst, group=group)
    for t in (f.R.contiguous(), f.R_x.contiguous(), f.perm, f.eigvals):
        dist.send(t, dst, group=group)


def recv_factors(n: int, src: int, device, group=None) -> SpectralFactors:
    k_t = torch.zeros(1, dtype=torch.int64, device=device)
    dist.recv(k_t, src, group=group)
    k = int(k_t.item())
    R = torch.empty((k, n), dtype=torch.float64, device=device)
    Rx = torch.empty((k, n), dtype=torch.float64, device=device)
    perm = torch.empty(n, dtype=torch.int64, device=device)
    eig = torch.empty(n, dtype=torch.float64, device=device)
    for t in (R, Rx, perm, eig):
        dist.recv(t, src, group=group)
    return SpectralFactors(R=R, R_x=Rx, perm=perm, eigvals=eig, k=k)


def exchange_order(groups: Sequence[Tuple[int, Sequence[int]]], plan: BlockPlan) -> List[Tuple[int, int, int]]:
    """(group, src, dst) hand-offs in ONE global order, narrow groups first (their solves finish first); every
    rank walks this list and takes part in the entries that name it, so any two ranks meet in the same order."""
    out = []
    for gi in sorted(range(len(groups)), key=lambda g: (solve_cost(groups[g][0]), g)):
        src = plan.solve_owner[gi]
        for dst in sorted(set(plan.loop_owner[gi])):
            if dst != src:
                out.append((gi, src, dst))
    return out


# ----------------------------------------------------------------------------- one decoder block on N GPUs
def quantize_block_parallel(x_shards: Sequence[Sequence[torch.Tensor]], weights: Sequence[Sequence[torch.Tensor]],
                            groups: Sequence[Tuple[int, Sequence[int]]], bits: int = 4, group_size: int = 128,
                            sym: bool = False, eps: float = 1e-4, block_size: int = 1024,
                            group=None, timers: Optional[dict] = None,
                            plan: Optional[BlockPlan] = None) -> Dict[Tuple[int, int], object]:
    """Token-sharded Hessians + LPT-scheduled solves and loops for one decoder block.

    x_shards[g] : this rank's calibration batches for group g (each (rows, n) or (B, S, n))
    weights[g][l]: the Linear's weight (every rank holds it; only the owner quantises it)
    Returns {(g, l): QuantizedLinear} for the Linears this rank owns.

    Order on every rank: (1) accumulate the local tokens of every group and all-reduce H (fp64, the one
    collective; widest group first so that its owner can start the long solve early); (2) solve the groups this
    rank owns, widest first; (3) loops whose factors are local; (4) point-to-point hand-offs of factors in one
    global order, narrow groups first; (5) the remaining loops.  `timers` (optional dict) receives CUDA events
    around the all-reduces: timers["allreduce"] = [(bytes, start_event, end_event), ...]."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    dev = weights[0][0].device
    if plan is None:
        plan = plan_block(groups, world)
    order = sorted(range(len(groups)), key=lambda g: (-solve_cost(groups[g][0]), g))
    hessians: Dict[int, torch.Tensor] = {}
    for gi in order:
        n = groups[gi][0]
        acc = ShardedHessianAccumulator(n, dev, group=group)
        for xb in x_shards[gi]:
            acc.add_batch(xb)
        if timers is not None and world > 1:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            acc.check()
            acc.verify = False
            e0.record()
            acc.n_samples = allreduce_hessian(acc.H, acc.n_samples, group)
            acc._reduced = True
            e1.record()
            timers.setdefault("allreduce", []).append((acc.H.numel() * 8, e0, e1))
        hessians[gi] = acc.get_hessian()            # the one collective
        del acc
    factors: Dict[int, SpectralFactors] = {}
    for gi in order:
        if plan.solve_owner[gi] == rank:
            factors[gi] = spectral_solve(hessians[gi], eps, "energy")
    hessians.clear()
    out: Dict[Tuple[int, int], object] = {}

    def run_loops(gi):
        f = factors[gi]
        for li, _m in enumerate(groups[gi][1]):
            if plan.loop_owner[gi][li] == rank and (gi, li) not in out:
                q = Quantizer(bits, group_size, sym)
                out[(gi, li)] = gptq_quantize(weights[gi][li], f.R, q, f.perm, block_size, True, f.R_x)

    for gi in sorted(factors):
        run_loops(gi)
    for gi, src, dst in exchange_order(groups, plan):
        if rank == src:
            send_factors(factors[gi], dst, group)
        elif rank == dst:
            factors[gi] = recv_factors(groups[gi][0], src, dev, group)
            run_loops(gi)
    return out
