#!/usr/bin/env python
"""BASELINE config 1 measured: 2-layer MLP on MNIST-shaped synthetic data, synchronous PS, world size 2, CPU — the reference's
`mpirun -n 2` plumbing scenario (`/root/reference/Makefile:2`).  Arms, all in one invocation per rank:

* ``host-ps`` / ``host-allgather``: this package's host engine (C pickler, native shm rings or gloo), per-parameter messages;
* ``host-ps-coalesced``: the same with one framed message per step;
* ``ref-equivalent``: the stock-tools re-creation of the reference's per-step algorithm (``baseline/comparator.py::RefEquivalentSGD``:
  hook → D2H → pickle → gloo all-gather of lengths and bytes → unpickle → sum → eager SGD).

    python -m pytorch_ps_mpi_b200.launch -n 2 bench/host_mlp_config1.py --out profiles/bench_cpu_mlp_config1.json
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pytorch_ps_mpi_b200 as ps          # noqa: E402
from pytorch_ps_mpi_b200.models import mnist_mlp   # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=60)
ap.add_argument("--warmup", type=int, default=10)
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--hidden", type=int, default=128)
ap.add_argument("--rounds", type=int, default=5)
ap.add_argument("--out", default="")
a = ap.parse_args()

w = ps.runtime.init()
torch.set_num_threads(max(1, (os.cpu_count() or 2) // max(w.size, 1) // 2))
rows = []


def run(name, make_opt):
    torch.manual_seed(0)
    model = mnist_mlp(hidden=a.hidden)
    opt = make_opt(model)
    gen = torch.Generator().manual_seed(100 + w.rank)
    batches = [(torch.randn(a.batch, 1, 28, 28, generator=gen), torch.randint(0, 10, (a.batch,), generator=gen)) for _ in range(8)]
    t0 = 0.0
    for s in range(a.warmup + a.steps):
        if s == a.warmup:
            w.barrier()
            t0 = time.perf_counter()
        x, y = batches[s % 8]
        opt.zero_grad()
        torch.nn.functional.cross_entropy(model(x), y).backward()
        opt.step()
    w.barrier()
    dt = time.perf_counter() - t0
    every = w.all_gather_object(dt)
    if hasattr(opt, "close"):
        opt.close()
    if w.rank == 0:
        ms = max(every) / a.steps * 1e3
        rows.append({"arm": name, "ms_per_step": ms, "samples_per_s": a.batch * w.size / (ms / 1e3)})


hyper = dict(lr=0.05, momentum=0.9)
from baseline.comparator import ComparatorSGD   # noqa: E402
ARMS = [("host-ps", lambda m: ps.SGD(m.named_parameters(), m.parameters(), mode="ps", engine="host", **hyper)),
        ("host-allgather", lambda m: ps.SGD(m.named_parameters(), m.parameters(), mode="allgather", engine="host", **hyper)),
        ("host-ps-coalesced", lambda m: ps.SGD(m.named_parameters(), m.parameters(), mode="ps", engine="host", coalesce=True, **hyper)),
        ("ref-equivalent", lambda m: ComparatorSGD(m.named_parameters(), lr=0.05, momentum=0.9, weight_decay=0.0, kind="host"))]
for rnd in range(a.rounds):              # interleaved rounds: a shared sandbox drifts by 2-4x within minutes
    for name, mk in ARMS:
        run(name, mk)
if w.rank == 0:
    import statistics
    merged = []
    for name, _ in ARMS:
        ms = sorted(r["ms_per_step"] for r in rows if r["arm"] == name)
        merged.append({"arm": name, "ms_per_step": ms[0], "ms_per_step_median": statistics.median(ms), "rounds_ms": ms,
                       "samples_per_s": a.batch * w.size / (ms[0] / 1e3)})
    rows = merged
if w.rank == 0:
    ref = next(r for r in rows if r["arm"] == "ref-equivalent")
    out = {"config": f"BASELINE config 1: mnist_mlp(hidden={a.hidden}), batch {a.batch}/rank, world {w.size}, CPU, "
                     f"transport={os.environ.get('PSB200_TRANSPORT', 'shm')}, best of {a.rounds} interleaved rounds of {a.steps} steps after {a.warmup} warm-up, "
                     "max over ranks",
           "host": f"{os.cpu_count()} logical CPUs (sandbox)", "rows": rows,
           "vs_ref_equivalent": {r["arm"]: r["samples_per_s"] / ref["samples_per_s"] for r in rows if r is not ref}}
    print(json.dumps(out))
    if a.out:
        with open(a.out, "w") as f:
            json.dump(out, f, indent=1)
ps.runtime.shutdown()
