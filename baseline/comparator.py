"""LABELLED reference-EQUIVALENT comparators (NOT the reference itself).

The reference cannot be installed or even imported here (no ``setup.py``; ``mpi4py`` / ``blosc`` /
``codings`` missing; ``/root/reference/mpi_comms.py:50`` is a SyntaxError on Python 3.12), so the
live baselines are re-creations of *what it does per step*, written against stock tools only
(``pickle``, ``numpy``, ``torch.distributed`` gloo / NCCL) — none of this repo's engines, kernels or
transports are on these paths:

``ComparatorSGD(kind='host')``  → :class:`RefEquivalentSGD`
    the reference's wired algorithm, step by step (``/root/reference/ps.py:92-190``,
    ``mpi_comms.py:32-58,144-193``): per-parameter backward hook → thread pool → synchronous
    device→host copy (``to_np``) → ``pickle.dumps`` → level-0 framing → exchange of message lengths
    (``Iallgather`` of one int per parameter) → variable-size all-gather of the host bytes
    (``Iallgatherv``; gloo here, mpi4py there) → ``pickle.loads`` → host→device copies (``to_torch``) →
    Python ``sum(grads)`` → eager per-parameter SGD ops (``ps.py:197-214``).
    Where the reference would be slower than this re-creation it is noted inline (every choice is generous to it).
``ComparatorSGD(kind='nccl')``  → :class:`NcclPS`
    the obvious library baseline for a rank-0 PS: NCCL ``reduce`` of every gradient to rank 0, fused
    ``torch.optim.SGD`` step there, NCCL ``broadcast`` of every parameter back.
``ComparatorSGD(kind='engine-host')``
    this repo's own *host engine* (C pickler, shm rings): faster than the reference would be; for our tables only.

The first two are what "a path that only calls NCCL/MPI for the named ops" looks like — the baseline the
fused kernels are measured against.  ``bench.py`` runs them in the SAME invocation as the product arm and prints
``vs_comparator`` so every driver record carries a same-box, same-run ratio.
"""
from __future__ import annotations

import pickle
import time
from concurrent.futures import ThreadPoolExecutor
from functools import partial

import numpy as np
import torch
import torch.distributed as dist

import pytorch_ps_mpi_b200 as ps

_HDR = 16          # blosc's level-0 frame: a 16-byte header in front of the raw bytes (mpi_comms.py:18-26)


class NcclPS:
    def __init__(self, named_params, lr, momentum, weight_decay):
        self.params = [p for _, p in named_params]
        self.w = ps.runtime.world()
        self.inner = torch.optim.SGD(self.params, lr=lr, momentum=momentum, weight_decay=weight_decay)
        self.group = None
        if self.w.size > 1:
            # the product's default group may be gloo (PSB200_PG_BACKEND=gloo): the comparator brings its own NCCL group
            self.group = dist.group.WORLD if self.w.backend == "nccl" else dist.new_group(backend="nccl")

    def zero_grad(self, set_to_none=True):
        self.inner.zero_grad(set_to_none=set_to_none)

    def step(self):
        if self.w.size > 1:
            works = [dist.reduce(p.grad, dst=0, op=dist.ReduceOp.SUM, async_op=True, group=self.group)
                     for p in self.params if p.grad is not None]
            for wk in works:
                wk.wait()
        if self.w.rank == 0:
            self.inner.step()
        if self.w.size > 1:
            works = [dist.broadcast(p.data, src=0, async_op=True, group=self.group) for p in self.params]
            for wk in works:
                wk.wait()
        return None, {}

    def close(self):
        pass


class RefEquivalentSGD:
    """The reference's per-step algorithm on stock tools (see the module docstring for the line-by-line map)."""

    def __init__(self, named_params, lr, momentum, weight_decay):
        self.named = list(named_params)
        self.lr, self.momentum, self.wd = lr, momentum, weight_decay
        self.w = ps.runtime.world()
        self.group = self.w.cpu_group                      # host bytes between processes (mpi4py → gloo)
        self.pool = ThreadPoolExecutor(max_workers=200)    # ps.py:85
        self.futures, self.names = [], []
        self.buf = {}
        self.by_name = dict(self.named)
        self._hooks = [p.register_hook(partial(self._hook, name=n)) for n, p in self.named if p.requires_grad]

    # ps.py:92-101 — hook → pool thread: encode (identity) → to_np → pickle → frame
    def _format(self, grad):
        g = grad.detach()
        # to_np (mpi_comms.py:32-43) is a synchronous .cpu().numpy(); numpy has no bfloat16, so the 2-byte payload
        # travels as int16 (the reference itself only handles cuda.FloatTensor: it would ship 4 bytes per element)
        host = g.contiguous().cpu()
        arr = host.view(torch.int16).numpy() if host.dtype == torch.bfloat16 else host.numpy()
        msg = pickle.dumps({"grad": arr, "bf16": host.dtype == torch.bfloat16}, protocol=pickle.HIGHEST_PROTOCOL)
        return b"\0" * _HDR + msg          # level-0 "compression" = header + memcpy

    def _hook(self, grad, name):
        self.futures.append(self.pool.submit(self._format, grad))
        self.names.append(name)

    def zero_grad(self, set_to_none=True):
        for _, p in self.named:
            p.grad = None

    def step(self):
        t0 = time.time()
        msgs = [f.result() for f in self.futures]          # ps.py:129-138
        names, self.futures, self.names = self.names, [], []
        n = self.w.size
        dev = self.named[0][1].device
        if n > 1:
            # Iallgather.prepare (mpi_comms.py:150-158): one int per parameter — here a single collective for all of them
            mine = torch.tensor([len(m) for m in msgs], dtype=torch.int64)
            sizes = [torch.empty_like(mine) for _ in range(n)]
            dist.all_gather(sizes, mine, group=self.group)
        for i, (name, msg) in enumerate(zip(names, msgs)):
            if n > 1:
                # Iallgatherv (mpi_comms.py:160-163); gloo needs equal counts, so pad to the longest message
                cap = int(max(s[i] for s in sizes))
                send = torch.zeros(cap, dtype=torch.uint8)
                send[: len(msg)] = torch.frombuffer(bytearray(msg), dtype=torch.uint8)
                recv = [torch.empty(cap, dtype=torch.uint8) for _ in range(n)]
                dist.all_gather(recv, send, group=self.group)
                blobs = [bytes(r[: int(sizes[k][i])].numpy()) for k, r in enumerate(recv)]
            else:
                blobs = [msg]
            grads = []
            for b in blobs:                                  # recv: decompress → pickle.loads → to_torch (H2D)
                obj = pickle.loads(b[_HDR:])
                t = torch.from_numpy(obj["grad"])
                if obj["bf16"]:
                    t = t.view(torch.bfloat16)
                grads.append(t.to(dev, non_blocking=True))
            p = self.by_name[name]
            d_p = sum(grads)                                 # ps.py:176 (N-1 eager adds)
            # ps.py:197-214 — eager SGD ops
            if self.wd != 0:
                d_p = d_p.add(p.data, alpha=self.wd)
            if self.momentum != 0:
                buf = self.buf.get(name)
                if buf is None:
                    buf = self.buf[name] = d_p.clone()
                else:
                    buf.mul_(self.momentum).add_(d_p)
                d_p = buf
            p.data.add_(d_p.to(p.dtype), alpha=-self.lr)
        return None, {"step_time": time.time() - t0}

    def close(self):
        for h in self._hooks:
            h.remove()
        self.pool.shutdown(wait=False)


def ComparatorSGD(named_params, lr=0.05, momentum=0.9, weight_decay=1e-4, kind: str = "host"):
    named = list(named_params)
    if kind == "nccl":
        return NcclPS(named, lr, momentum, weight_decay)
    if kind == "host":
        return RefEquivalentSGD(named, lr, momentum, weight_decay)
    if kind == "engine-host":
        return ps.SGD(named, [p for _, p in named], lr=lr, momentum=momentum, weight_decay=weight_decay,
                      code=ps.Identity(), engine="host", mode="allgather", cuda=True)
    raise ValueError(kind)
