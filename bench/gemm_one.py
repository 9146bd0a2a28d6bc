"""One representative bcast_gemm launch (BERT FFN-in, 16384x3072x768) for an ncu capture."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_ps_mpi_b200.ops.linear import bcast_linear
dev = torch.device("cuda", 0)
x = (torch.randn(16384, 768, device=dev) / 28).bfloat16()
w = torch.randn(3072, 768, device=dev).bfloat16()
b = torch.randn(3072, device=dev).bfloat16()
for _ in range(3):
    y = bcast_linear(x, w, b, True)
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()))
