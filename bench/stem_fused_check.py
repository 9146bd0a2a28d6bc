#!/usr/bin/env python
"""Validate and time the EXPERIMENTAL fused stem kernel (``csrc/kernels/stem_kernels.cu``) — needs a GPU.

1. numerics: ``stem_fwd`` vs ``F.conv2d`` in fp32 on several image sizes (borders, odd heights, partial tiles), and its
   BatchNorm sums vs sums of its own bf16 output;
2. gradients: weight gradient of the fused path vs the im2col + GEMM path;
3. model: one ResNet-18 forward/backward with ``PSB200_STEM=fused`` semantics vs the default path;
4. timing (batch 256, 224x224): stem + BN1 forward, default vs fused, CUDA events, L2 flushed.

Exit code 1 if any check fails.  One JSON line per check on stdout and in ``gpurun_out/stem_fused_check.jsonl``.
"""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pytorch_ps_mpi_b200.models import resnet as resnet_mod          # noqa: E402
from pytorch_ps_mpi_b200.ops import ext                              # noqa: E402
from pytorch_ps_mpi_b200.ops.batchnorm import FusedBatchNormAct2d    # noqa: E402
from pytorch_ps_mpi_b200.ops.stem import _w2d, stem_conv, stem_conv_fused   # noqa: E402

OUT = []
BAD = 0


def emit(**rec):
    global BAD
    if rec.get("ok") is False:
        BAD += 1
    line = json.dumps(rec)
    print(line, flush=True)
    OUT.append(line)


def bench(fn, flush, iters=10):
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


def main():
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    m = ext.cuda()
    w = (torch.randn(64, 3, 7, 7, device=dev) * 0.05).bfloat16()

    # 1. numerics
    for (n, h, wd) in [(2, 224, 224), (3, 64, 64), (1, 30, 40), (5, 17, 8), (2, 225, 256), (300, 32, 32)]:
        x = torch.randn(n, 3, h, wd, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
        y, sums = m.stem_fwd(x, _w2d(w), True)
        ref = F.conv2d(x.float(), w.float(), stride=2, padding=3)
        err = (y.float() - ref).abs().max().item() / max(ref.abs().max().item(), 1e-6)
        yf = y.float()
        s_ref = torch.cat([yf.sum((0, 2, 3)), (yf * yf).sum((0, 2, 3))])
        serr = ((sums - s_ref).abs() / (s_ref.abs() + 1.0)).max().item()
        emit(check="numerics", shape=[n, h, wd], max_rel_err=err, sums_rel_err=serr,
             channels_last=bool(y.is_contiguous(memory_format=torch.channels_last)), ok=bool(err < 2e-2 and serr < 1e-3))

    # 2. weight gradient, fused vs im2col + GEMM
    x = torch.randn(8, 3, 224, 224, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
    gy = torch.randn(8, 64, 112, 112, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
    wa = w.clone().requires_grad_(True)
    wb = w.clone().requires_grad_(True)
    stem_conv(x, wa).backward(gy)
    stem_conv_fused(x, wb)[0].backward(gy)
    gerr = (wa.grad.float() - wb.grad.float()).abs().max().item() / max(wa.grad.float().abs().max().item(), 1e-6)
    emit(check="wgrad", max_rel_diff=gerr, ok=bool(gerr < 1e-2))

    # 2b. implicit weight-gradient kernel vs the same reference, several shapes (partial K steps, small grids)
    from pytorch_ps_mpi_b200.ops.stem import stem_wgrad_implicit
    for (n, h, wd) in [(8, 224, 224), (2, 64, 64), (1, 30, 40), (3, 17, 8), (2, 33, 256)]:
        xs = torch.randn(n, 3, h, wd, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
        oh, ow = (h - 1) // 2 + 1, (wd - 1) // 2 + 1
        g = torch.randn(n, 64, oh, ow, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
        a = m.im2col_stem(xs).float()
        ref = g.permute(0, 2, 3, 1).reshape(-1, 64).float().t() @ a               # [64,176]
        got = stem_wgrad_implicit(xs, g).float()
        e = (got - ref).abs().max().item() / max(ref.abs().max().item(), 1e-6)
        emit(check="wgrad_implicit", shape=[n, h, wd], max_rel_err=e, ok=bool(e < 1e-2))
    flush0 = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    xs = torch.randn(256, 3, 224, 224, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
    g = torch.randn(256, 64, 112, 112, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
    t_impl = bench(lambda: stem_wgrad_implicit(xs, g), flush0)
    t_gemm = bench(lambda: g.permute(0, 2, 3, 1).reshape(-1, 64).t() @ m.im2col_stem(xs), flush0)
    emit(check="wgrad_timing", batch=256, implicit_ms=t_impl, im2col_plus_gemm_ms=t_gemm, ok=True)
    del xs, g, flush0

    # 3. whole model, default vs fused stem
    losses = {}
    for fused in (False, True):
        torch.manual_seed(1)
        net = resnet_mod.resnet18(num_classes=100).to(dev).bfloat16().to(memory_format=torch.channels_last).train()
        resnet_mod._FUSED_STEM = fused
        xi = torch.randn(16, 3, 224, 224, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
        tgt = torch.randint(0, 100, (16,), device=dev)
        loss = F.cross_entropy(net(xi).float(), tgt)
        loss.backward()
        losses[fused] = (loss.item(), net.conv1.weight.grad.float().norm().item(), net.bn1.running_var.float().mean().item())
    resnet_mod._FUSED_STEM = False
    d = abs(losses[True][0] - losses[False][0])
    emit(check="model", default=losses[False], fused=losses[True], ok=bool(d < 5e-2 and abs(losses[True][2] - losses[False][2]) < 1e-2))

    # 4. timing at the benchmark shape
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    x = torch.randn(256, 3, 224, 224, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
    bn = FusedBatchNormAct2d(64, relu=True).to(dev).bfloat16().train()
    w2 = _w2d(w)
    with torch.no_grad():
        t_default = bench(lambda: bn(stem_conv(x, w)), flush)
        t_fused = bench(lambda: (lambda ys: bn(ys[0], sums=ys[1]))(stem_conv_fused(x, w)), flush)
        t_kernel = bench(lambda: m.stem_fwd(x, w2, True), flush)
        t_conv_only = bench(lambda: stem_conv(x, w), flush)
    emit(check="timing", batch=256, default_stem_bn_ms=t_default, fused_stem_bn_ms=t_fused, fused_kernel_ms=t_kernel,
         default_conv_only_ms=t_conv_only, ok=True)

    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "stem_fused_check.jsonl"), "w") as f:
        f.write("\n".join(OUT) + "\n")
    if BAD:
        print(f"{BAD} check(s) FAILED", file=sys.stderr)
        sys.exit(1)


if __name__ == "__main__":
    main()
