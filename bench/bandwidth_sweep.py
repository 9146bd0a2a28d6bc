#!/usr/bin/env python
"""BASELINE.json config 5: PS gather + broadcast over a parameter vector of 1 KB … 1 GB at N GPUs.

For every size the SAME public objects are exercised three ways and device-timed (CUDA events,
max over ranks, warm-up, distinct gradients every iteration):

``fused``   this framework's device engine: encode → epoch flag → ``psb_update_kernel`` (pull every
            rank's wire tile over NVLink, sum, SGD on fp32 masters, multimem.st / P2P publish) → wait.
``nccl``    library baseline: ``dist.reduce`` to rank 0 + one fused axpy + ``dist.broadcast``.
``host``    reference-equivalent host path (pickle framing over shm/gloo) — small sizes only.

Reported per size: µs per round trip, gather "bus" GB/s = (N-1)·B_wire / t, broadcast GB/s = B_param / t,
and the achieved fraction of the roofline  max((N-1)·B_wire, B_bcast) / 770 GB/s (measured peer
copy per direction, B200_PROFILING.md).  One JSON line per (size, impl) on rank 0.

    python -m torch.distributed.run --nproc-per-node 8 bench/bandwidth_sweep.py --max-mb 1024
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pytorch_ps_mpi_b200 as ps   # noqa: E402

NVLINK_GBS = 770.0


def timed(w, device, fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize(device)
    w.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize(device)
    t = torch.tensor([s.elapsed_time(e) / iters], dtype=torch.float64)
    if w.size > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=w.cpu_group)
    return float(t.item()) * 1e3   # µs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--min-kb", type=float, default=1)
    ap.add_argument("--max-mb", type=float, default=256)
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--code", default="identity")
    ap.add_argument("--impls", default="fused,nccl")
    ap.add_argument("--reduce", default="auto")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    w = ps.runtime.init()
    dev = w.device
    dtype = {"bf16": torch.bfloat16, "fp32": torch.float32}[a.dtype]
    esz = 2 if dtype == torch.bfloat16 else 4
    sizes = []
    b = a.min_kb * 1024
    while b <= a.max_mb * (1 << 20) + 1:
        sizes.append(int(b))
        b *= 4
    rows = []
    for nbytes in sizes:
        n = max(8, nbytes // esz)
        iters = 200 if nbytes < (1 << 20) else (50 if nbytes < (64 << 20) else 10)
        for impl in a.impls.split(","):
            if impl == "host" and nbytes > (16 << 20):
                continue
            torch.manual_seed(0)
            p = torch.nn.Parameter(torch.zeros(n, device=dev, dtype=dtype))
            grads = [torch.randn(n, device=dev).to(dtype) for _ in range(2)]
            k = [0]
            if impl in ("fused", "host"):
                code = ps.Identity() if a.code == "identity" else (
                    ps.TopK(ratio=float(a.code.split(":")[1]), values="bf16") if a.code.startswith("topk") else
                    ps.Cast(a.code.split(":")[1]))
                opt = ps.SGD([("v", p)], [p], lr=1e-3, code=code, mode="ps",
                             engine="device" if impl == "fused" else "host", reduce=a.reduce, cuda=True)
                eng = opt._engine

                def fn():
                    k[0] ^= 1
                    if eng is not None:
                        eng.on_grad(grads[k[0]], "v", p)
                    else:
                        p.grad = grads[k[0]]
                        opt.async_code(grads[k[0]], name="v", encode=code.encode)
                    opt.step()
                wire = eng.bpt * eng.layout.ntiles if eng is not None else n * esz
            else:
                buf = torch.zeros(n, device=dev, dtype=dtype)

                def fn():
                    k[0] ^= 1
                    buf.copy_(grads[k[0]])
                    if w.size > 1:
                        dist.reduce(buf, dst=0)
                    if w.rank == 0:
                        p.data.add_(buf, alpha=-1e-3)
                    if w.size > 1:
                        dist.broadcast(p.data, src=0)
                wire = n * esz
                opt = None
            us = timed(w, dev, fn, iters)
            gather_b = (w.size - 1) * wire
            bcast_b = n * esz
            roof_us = max(gather_b, bcast_b if w.size > 1 else 0) / (NVLINK_GBS * 1e3) if w.size > 1 else 0.0
            row = {"bytes": n * esz, "impl": impl, "n_gpus": w.size, "us": us,
                   "gather_GBs": gather_b / us / 1e3 if w.size > 1 else None,
                   "bcast_GBs": bcast_b / us / 1e3,
                   "roofline_us": roof_us, "roofline_frac": (roof_us / us) if roof_us else None,
                   "wire_bytes": wire, "dtype": a.dtype, "code": a.code,
                   "multicast": bool(getattr(getattr(opt, "_engine", None), "arena", None) and opt._engine.arena.has_multicast),
                   "reduce": getattr(getattr(opt, "_engine", None), "reduce", None)}
            if opt is not None:
                if opt._engine is not None:
                    opt._engine.check()
                opt.close()
            rows.append(row)
            if w.rank == 0:
                print(json.dumps(row), flush=True)
            w.barrier()
    if w.rank == 0 and a.out:
        with open(a.out, "w") as f:
            json.dump(rows, f, indent=1)
    ps.runtime.shutdown()


if __name__ == "__main__":
    main()
