#!/usr/bin/env python
"""BASELINE.json config 1: 2-layer MLP on MNIST-shaped synthetic data, synchronous rank-0 parameter
server, world size 2, CPU plumbing (the reference's `mpirun -n 2` scenario).

    python -m pytorch_ps_mpi_b200.launch -n 2 examples/train_mlp_cpu.py --mode ps --code topk
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pytorch_ps_mpi_b200 as ps          # noqa: E402
from pytorch_ps_mpi_b200.models import mnist_mlp   # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="ps", choices=["ps", "allgather", "async"])
ap.add_argument("--code", default="identity", choices=["identity", "cast", "scale", "topk"])
ap.add_argument("--steps", type=int, default=20)
a = ap.parse_args()

w = ps.runtime.init()
torch.manual_seed(0)
model = mnist_mlp(hidden=128)
code = {"identity": ps.Identity(), "cast": ps.Cast("bf16"), "scale": ps.Scale("int8"), "topk": ps.TopK(ratio=0.1)}[a.code]
opt = ps.SGD(model.named_parameters(), model.parameters(), lr=0.05, momentum=0.9, code=code, mode=a.mode)

gen = torch.Generator().manual_seed(100 + w.rank)            # every rank sees different data
if a.mode == "async" and w.rank == 0 and w.size > 1:
    n = opt.serve()                                           # AsySG-InCon: rank 0 only serves
    print(f"[server] applied {n} updates")
else:
    for step in range(a.steps):
        x, y = torch.randn(32, 1, 28, 28, generator=gen), torch.randint(0, 10, (32,), generator=gen)
        opt.zero_grad()
        loss = torch.nn.functional.cross_entropy(model(x), y)
        loss.backward()                                       # hooks encode each gradient as it appears
        _, data = opt.step()                                  # (loss, data) like the reference
        if step % 5 == 0:
            print(f"[rank {w.rank}] step {step:3d} loss {loss.item():.4f} comm_wait {data['comm_wait'] * 1e3:.2f} ms "
                  f"msg_bytes {data['msg_bytes']:.0f} packaged_bytes {data['packaged_bytes']:.0f}")
opt.close()
ps.runtime.shutdown()
