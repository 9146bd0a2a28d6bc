"""One bcast_gemm launch for an ncu capture: `python bench/gemm_one.py [M N K]` (default BERT FFN-in)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_ps_mpi_b200.ops.linear import bcast_linear   # noqa: E402

M, N, K = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (16384, 3072, 768)
dev = torch.device("cuda", 0)
x = (torch.randn(M, K, device=dev) / K ** 0.5).bfloat16()
w = torch.randn(N, K, device=dev).bfloat16()
b = torch.randn(N, device=dev).bfloat16()
for _ in range(3):
    y = bcast_linear(x, w, b, True)
torch.cuda.synchronize()
print("ok", M, N, K, float(y.float().abs().mean()))
