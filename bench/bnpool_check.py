#!/usr/bin/env python
"""Validate and time the EXPERIMENTAL fused BatchNorm + ReLU + max-pool kernels (``bn_kernels.cu``) — needs a GPU.

* forward: fused vs ``FusedMaxPool2d(FusedBatchNormAct2d(relu)(x))`` — expected BIT-IDENTICAL (same bf16 rounding point);
* backward: dx / dgamma / dbeta vs the unfused pair, and running statistics;
* timing at the ResNet-18 stem-tail shape (256 x 64 x 112 x 112), forward and forward+backward.

Exit code 1 on a failed check; JSON lines on stdout and appended to ``gpurun_out/bnpool_check.jsonl``.
"""
import copy
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pytorch_ps_mpi_b200.ops.batchnorm import FusedBatchNormAct2d, fused_bn_relu_maxpool   # noqa: E402
from pytorch_ps_mpi_b200.ops.pooling import FusedMaxPool2d                                   # noqa: E402

BAD = 0


def emit(**rec):
    global BAD
    if rec.get("ok") is False:
        BAD += 1
    line = json.dumps(rec)
    print(line, flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "bnpool_check.jsonl"), "a") as f:
        f.write(line + "\n")


def bench(fn, dev, iters=10):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def rel(a, b):
    return (a.float() - b.float()).abs().max().item() / max(b.float().abs().max().item(), 1e-6)


def main():
    dev = torch.device("cuda", 0)
    pool = FusedMaxPool2d(3, 2, 1)
    for (n, c, h, w) in [(4, 64, 112, 112), (2, 64, 17, 23), (3, 128, 8, 8), (1, 8, 5, 4), (2, 256, 9, 14)]:
        torch.manual_seed(0)
        bn_a = FusedBatchNormAct2d(c, relu=True).to(dev).bfloat16().train()
        with torch.no_grad():
            bn_a.weight.copy_(torch.randn(c).abs() + 0.5)
            bn_a.bias.copy_(torch.randn(c) * 0.3)
        bn_b = copy.deepcopy(bn_a)
        x = (torch.randn(n, c, h, w, device=dev) * 2 + 0.3).bfloat16().contiguous(memory_format=torch.channels_last)
        xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
        ya = pool(bn_a(xa))
        yb = fused_bn_relu_maxpool(xb, bn_b)
        g = torch.randn_like(ya)
        ya.backward(g)
        yb.backward(g)
        same = bool(torch.equal(ya, yb))
        # each path sums its own batch statistics with float atomics (order-dependent in the last bit): a rare 1-ulp flip of a
        # bf16 output is legitimate
        near = float((ya.float() - yb.float()).abs().max()) <= 2.0 ** -7 * float(ya.float().abs().max()) and \
            float((ya != yb).float().mean()) < 1e-3
        e_dx, e_dg, e_db = rel(xb.grad, xa.grad), rel(bn_b.weight.grad, bn_a.weight.grad), rel(bn_b.bias.grad, bn_a.bias.grad)
        e_rm, e_rv = rel(bn_b.running_mean, bn_a.running_mean), rel(bn_b.running_var, bn_a.running_var)
        emit(check="bnpool", shape=[n, c, h, w], forward_bit_identical=same, dx=e_dx, dgamma=e_dg, dbeta=e_db, running_mean=e_rm,
             running_var=e_rv, ok=bool(near and e_dx < 2e-2 and e_dg < 2e-2 and e_db < 2e-2 and e_rm < 1e-5 and e_rv < 1e-5))

    bn = FusedBatchNormAct2d(64, relu=True).to(dev).bfloat16().train()
    x = torch.randn(256, 64, 112, 112, device=dev).bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    g = torch.randn(256, 64, 56, 56, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        f_unfused = bench(lambda: pool(bn(x)), dev)
        f_fused = bench(lambda: fused_bn_relu_maxpool(x, bn), dev)

    def fb(fn):
        x.grad = None
        fn().backward(g)

    fb_unfused = bench(lambda: fb(lambda: pool(bn(x))), dev)
    fb_fused = bench(lambda: fb(lambda: fused_bn_relu_maxpool(x, bn)), dev)
    emit(check="bnpool_timing", batch=256, fwd_unfused_ms=f_unfused, fwd_fused_ms=f_fused, fwdbwd_unfused_ms=fb_unfused,
         fwdbwd_fused_ms=fb_fused, ok=True)
    if BAD:
        print(f"{BAD} check(s) FAILED", file=sys.stderr)
        sys.exit(1)


if __name__ == "__main__":
    main()
