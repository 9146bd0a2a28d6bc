import torch, sys
sys.path.insert(0, '.')
from pytorch_ps_mpi_b200.ops import ext
from pytorch_ps_mpi_b200.ops.linear import bcast_linear
m = ext.cuda()
dev = torch.device("cuda", 0)
x = torch.randn(256, 3, 224, 224, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
w2d = torch.randn(64, 160, device=dev).bfloat16()
def t(fn, n=10):
    for _ in range(3): r = fn()
    torch.cuda.synchronize(); s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True); s.record()
    for _ in range(n): r = fn()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / n
a = m.im2col_stem(x); w2d = torch.randn(64, a.shape[1], device=dev).bfloat16()
print("im2col", t(lambda: m.im2col_stem(x)))
print("fwd gemm 2cta", t(lambda: bcast_linear(a, w2d, variant=2)))
print("fwd gemm 1cta", t(lambda: bcast_linear(a, w2d, variant=1)))
print("fwd cublas", t(lambda: a @ w2d.t()))
gy = torch.randn(a.shape[0], 64, device=dev).bfloat16()
print("wgrad gy.t().contiguous() @ a", t(lambda: gy.t().contiguous() @ a))
print("wgrad gy.t() @ a", t(lambda: gy.t() @ a))
print("wgrad (a.t() @ gy)", t(lambda: a.t() @ gy))
gyf = gy.float()
print("wgrad fp32 out via einsum", t(lambda: torch.einsum('mk,mn->nk', a, gy)))
# split-K by hand: reshape into chunks and bmm then sum
def splitk(S=64):
    M = a.shape[0]; c = M // S
    return torch.bmm(gy[:c*S].view(S, c, 64).transpose(1, 2), a[:c*S].view(S, c, a.shape[1])).float().sum(0)
print("wgrad manual split-K bmm 64", t(lambda: splitk(64)))
print("wgrad manual split-K bmm 256", t(lambda: splitk(256)))
conv = torch.nn.Conv2d(3, 64, 7, 2, 3, bias=False).to(dev).to(memory_format=torch.channels_last).bfloat16()
torch.backends.cudnn.benchmark = True
print("cudnn fwd", t(lambda: conv(x)))
