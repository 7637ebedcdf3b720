"""
Several spectral solves in flight on ONE GPU.

At Qwen3-8B's n = 4096 a solve is latency-bound (grid barriers and short per-column phases; the
lower triangle even fits L2), so three of them - the q/k/v, o_proj and gate/up Hessians of a
decoder block, independent once their H is accumulated - finish sooner side by side than one after
another.  Each solve gets its own host thread (the C library keeps per-thread handles, pinned
staging and error state), its own CUDA stream and an SM budget (`tq_set_sm_budget`) that sizes
its cooperative panel kernels so that all of them are co-resident; cooperative launches are
gang-scheduled, so there is no deadlock when the budgets do not add up - the kernels just serialise.

    pool = SolverPool(workers=3)
    factors = pool.process_hessian_alt_many([H_qkv, H_o, H_gateup], 1e-4, "energy")

Results agree with the one-at-a-time call to rounding (same k and pivot order, R within 2e-12 relative at
n = 4096): the panel kernels add their per-CTA partial sums in a fixed order that depends on the grid size,
so a run is reproducible for a given budget, not across budgets.
"""
from __future__ import annotations

import queue
import threading
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from .gptq_utils import SpectralFactors, spectral_solve


class SolverPool:
    def __init__(self, workers: int = 3, device=None):
        dev = torch.device("cuda") if device is None else torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError(f"SolverPool: {dev} is not a CUDA device (there is no CPU path)")
        self.device = torch.device("cuda", torch.cuda.current_device() if dev.index is None else dev.index)
        self.workers = max(1, int(workers))
        self._jobs: "queue.Queue" = queue.Queue()
        self._threads = [threading.Thread(target=self._run, args=(i,), daemon=True) for i in range(self.workers)]
        self.launches = [0] * self.workers          # libtruncgptq kernel launches per worker thread
        self._broken: Optional[BaseException] = None
        for t in self._threads:
            t.start()

    def _run(self, idx: int):
        try:
            torch.cuda.set_device(self.device)
            streams = {False: torch.cuda.Stream(device=self.device),
                       True: torch.cuda.Stream(device=self.device, priority=-1)}
            lib = _lib.load()
        except BaseException as ex:               # a worker that cannot start must not leave callers waiting
            self._broken = ex
            return
        while True:
            job = self._jobs.get()
            if job is None:
                return
            fn, budget, ready, done, out, slot, urgent = job
            stream = streams[bool(urgent)]
            try:
                lib.tq_set_sm_budget(int(budget))
                with torch.cuda.stream(stream):
                    stream.wait_event(ready)                  # inputs were produced on the caller's stream
                    out[slot] = fn()
                    ev = torch.cuda.Event()
                    ev.record(stream)
                out[slot] = (out[slot], ev, stream)
            except BaseException as ex:                       # surfaced in the caller
                out[slot] = ex
            finally:
                self.launches[idx] = int(lib.tq_launch_count())
                done.release()

    def submit(self, fn, sm_budget: Optional[int] = None, urgent: bool = False):
        """Start `fn()` on a worker (its stream first waits for everything enqueued so far on the caller's
        current stream) and return a handle for `result()`.  Tensors `fn` reads must stay alive until
        `result()` returns, or be `record_stream`-ed inside `fn` (it runs under the worker's stream).
        `urgent`: run on the worker's high-priority stream - for the job on the critical path, whose chains of
        short dependent kernels (divide and conquer merges, panel factorisations) would otherwise queue behind
        the full-grid kernels of the other jobs at every link."""
        if sm_budget is None:
            sms = torch.cuda.get_device_properties(self.device).multi_processor_count
            sm_budget = max(8, sms // self.workers)
        cur = torch.cuda.current_stream(self.device)
        ready = torch.cuda.Event()
        ready.record(cur)
        out: List = [None]
        done = threading.Semaphore(0)
        self._jobs.put((fn, sm_budget, ready, done, out, 0, urgent))
        return (out, done)

    def result(self, handle):
        """Wait for a submitted job; the caller's current stream waits for the worker's stream."""
        out, done = handle
        while not done.acquire(timeout=0.5):                  # never wait on a pool whose workers are gone
            if self._broken is not None or not any(t.is_alive() for t in self._threads):
                raise RuntimeError(f"SolverPool: worker threads are not running ({self._broken!r})")
        done.release()                                        # result() may be called again
        r = out[0]
        if isinstance(r, BaseException):
            raise r
        val, ev, stream = r
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        for t in _tensors(val):
            t.record_stream(cur)                              # allocated on the worker's stream, used on ours
        return val

    def map(self, fns, sm_budget: Optional[int] = None):
        """Run the callables concurrently (at most `workers` at a time); returns their results.  The
        caller's current stream waits for every result before it can use it.  `sm_budget`: one value for
        all jobs or one per job (default: the SMs split evenly)."""
        n = len(fns)
        if n == 0:
            return []
        if sm_budget is None:
            sms = torch.cuda.get_device_properties(self.device).multi_processor_count
            sm_budget = max(8, sms // min(n, self.workers))
        budgets = list(sm_budget) if isinstance(sm_budget, (list, tuple)) else [sm_budget] * n
        handles = [self.submit(fn, b) for fn, b in zip(fns, budgets)]
        return [self.result(h) for h in handles]

    def spectral_solve_many(self, Hs: Sequence[torch.Tensor], threshold: float = 0.0005,
                            threshold_method: str = "mean_trimmed",
                            sm_budget: Optional[int] = None) -> List[SpectralFactors]:
        return self.map([lambda H=H: spectral_solve(H, threshold, threshold_method) for H in Hs], sm_budget)

    def process_hessian_alt_many(self, Hs: Sequence[torch.Tensor], threshold: float = 0.0005,
                                 threshold_method: str = "mean_trimmed", sm_budget: Optional[int] = None
                                 ) -> List[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]]:
        """[process_hessian_alt(H, ...) for H in Hs], solved side by side."""
        return [(f.R, f.R_x, f.perm) for f in self.spectral_solve_many(Hs, threshold, threshold_method, sm_budget)]

    def launch_count(self) -> int:
        return sum(self.launches)

    def close(self):
        for _ in self._threads:
            self._jobs.put(None)
        for t in self._threads:
            t.join(timeout=5)


def _tensors(val):
    if isinstance(val, torch.Tensor):
        yield val
    elif isinstance(val, SpectralFactors):
        yield from (val.R, val.R_x, val.perm, val.eigvals)
    elif isinstance(val, (list, tuple)):
        for v in val:
            yield from _tensors(v)
