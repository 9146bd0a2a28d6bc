#!/usr/bin/env python
"""Where does the 2-CTA tcgen05 GEMM lose time on short-K shapes?  (round-2 diagnosis, needs a GPU)

Runs the production choice of ``psb_bcast_gemm2_kernel`` (``bcast_gemm2.cu``) next to every epilogue it can be built with
(the round-2 run of this script is ``profiles/gemm_variants_r2.jsonl``; it made ``epi3`` / ``epi1`` the defaults):

* ``epi0`` the round-1 epilogue (lane == row, row-strided 16-byte stores),
* ``epi1`` staged epilogue (padded smem transpose → full 128-byte lines),
* ``epi3`` TMA-store epilogue (swizzled staging, double-buffered ``cp.async.bulk.tensor`` stores; needs N % 8 == 0),

(the round-2 sweep also had an eight-warp epilogue and ``nostore`` / ``nomma`` diagnostic builds — rows ``epi2*`` / ``*.nostore`` /
``*.nomma`` of ``profiles/gemm_variants_r2.jsonl``; they were deleted after the decision).
Numerics of every storing variant are checked against an fp32 torch matmul first.  CUDA-event timing, L2 flushed
between iterations, one JSON line per (shape, variant) on stdout and in ``gpurun_out/gemm_variants.jsonl``.

Reading the table: if ``nostore`` closes the gap to cuBLAS the output stores are the problem (pick the better of
epi1 / a TMA store); if ``nomma`` is not much faster than the full kernel the epilogue path alone is too slow (epi2);
if neither moves, the loss is in the tile hand-off (accumulator stages / barriers), not in the epilogue.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pytorch_ps_mpi_b200.ops.linear import bcast_linear   # noqa: E402

TWO_CTA = 2
VARIANTS = [("prod", TWO_CTA)]            # epilogue selector 0 = the production choice (TMA store if N % 8 == 0, else staged)
for epi, sel in ((0, 4), (1, 1), (3, 3)):   # kernel template EPI → selector bits of `variant`
    VARIANTS.append((f"epi{epi}", TWO_CTA | sel << 4))

SHAPES = [("bert.ffn_in", 16384, 3072, 768), ("bert.qkv", 16384, 2304, 768), ("bert.ffn_out", 16384, 768, 3072),
          ("mlp.fc1", 8192, 4096, 784), ("stem", 256 * 112 * 112 // 8, 64, 176), ("square4096", 4096, 4096, 4096),
          ("ragged", 1000, 328, 264), ("ragged.unaligned_n", 515, 330, 72)]


def bench(fn, flush, iters=20):
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
    return ts[len(ts) // 2] * 1e-3


def main():
    dev = torch.device("cuda", 0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    out = open(os.path.join(ROOT, "gpurun_out", "gemm_variants.jsonl"), "w")
    bad = 0
    for name, M, N, K in SHAPES:
        torch.manual_seed(0)
        x = (torch.randn(M, K, device=dev) / K ** 0.5).bfloat16()
        w = torch.randn(N, K, device=dev).bfloat16()
        b = torch.randn(N, device=dev)
        ref = torch.relu(x.float() @ w.float().t() + b)
        t_lib = bench(lambda: torch.nn.functional.linear(x, w), flush)
        fl = 2.0 * M * N * K
        for tag, v in VARIANTS:
            if (v >> 4) & 15 == 3 and N % 8:
                continue
            rec = {"shape": name, "M": M, "N": N, "K": K, "variant": tag, "code": v}
            if (v >> 8) == 0:             # a storing variant: numerics first (bias + ReLU exercised too)
                y = bcast_linear(x, w, b, relu=True, variant=v).float()
                err = (y - ref).abs().max().item() / max(ref.abs().max().item(), 1e-6)
                rec["max_rel_err"] = err
                if not err < 2e-2:
                    bad += 1
                    rec["FAILED"] = True
            t = bench(lambda: bcast_linear(x, w, variant=v), flush)
            rec.update(us=t * 1e6, tflops=fl / t / 1e12, cublas_us=t_lib * 1e6)
            line = json.dumps(rec)
            print(line, flush=True)
            out.write(line + "\n")
            out.flush()
    out.close()
    if bad:
        print(f"{bad} variant(s) FAILED the numerics check", file=sys.stderr)
        sys.exit(1)


if __name__ == "__main__":
    main()
