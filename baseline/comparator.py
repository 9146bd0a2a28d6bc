"""LABELLED reference-EQUIVALENT comparators (NOT the reference itself).

The reference cannot be installed or even imported here (no ``setup.py``; ``mpi4py`` / ``blosc`` /
``codings`` missing; ``/root/reference/mpi_comms.py:50`` is a SyntaxError on Python 3.12), so the
live baseline is a faithful re-creation of *what it does per step* on the same box:

``ComparatorSGD(kind='host')``
    the reference's wired algorithm (``/root/reference/ps.py:92-190``): per-parameter backward hook →
    thread pool → device→host copy → pickle-style framing → variable-size all-gather of host bytes
    between processes → unpickle → host→device copies → Python ``sum`` → eager per-parameter SGD ops.
    Implemented by this repo's *host engine* (``engine='host', mode='allgather'``) over the shm /
    gloo transport (the stand-in for mpi4py's shared-memory BTL).
``ComparatorSGD(kind='nccl')``
    the obvious library baseline for a rank-0 PS: NCCL ``reduce`` of every gradient to rank 0, fused
    ``torch.optim.SGD`` step there, NCCL ``broadcast`` of every parameter back.

Both are what "a path that only calls NCCL/MPI for the named ops" looks like — the baseline the
fused kernels are measured against.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

import pytorch_ps_mpi_b200 as ps


class _NcclPS:
    def __init__(self, named_params, lr, momentum, weight_decay):
        self.params = [p for _, p in named_params]
        self.w = ps.runtime.world()
        self.inner = torch.optim.SGD(self.params, lr=lr, momentum=momentum, weight_decay=weight_decay)
        if self.w.size > 1 and self.w.backend != "nccl":
            raise RuntimeError("the NCCL comparator needs the default process group on NCCL")

    def zero_grad(self, set_to_none=True):
        self.inner.zero_grad(set_to_none=set_to_none)

    def step(self):
        if self.w.size > 1:
            works = [dist.reduce(p.grad, dst=0, op=dist.ReduceOp.SUM, async_op=True) for p in self.params
                     if p.grad is not None]
            for wk in works:
                wk.wait()
        if self.w.rank == 0:
            self.inner.step()
        if self.w.size > 1:
            works = [dist.broadcast(p.data, src=0, async_op=True) for p in self.params]
            for wk in works:
                wk.wait()
        return None, {}

    def close(self):
        pass


def ComparatorSGD(named_params, lr=0.05, momentum=0.9, weight_decay=1e-4, kind: str = "host"):
    named = list(named_params)
    if kind == "nccl":
        return _NcclPS(named, lr, momentum, weight_decay)
    return ps.SGD(named, [p for _, p in named], lr=lr, momentum=momentum, weight_decay=weight_decay,
                  code=ps.Identity(), engine="host", mode="allgather", cuda=True)
