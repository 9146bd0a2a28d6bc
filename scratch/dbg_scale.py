import torch, sys
sys.path.insert(0, '.')
import pytorch_ps_mpi_b200 as ps
from tests.test_gpu_kernels import Virtual, SHAPES, sgd_h
torch.manual_seed(0)
dtype = torch.bfloat16
code = ps.Scale("fp8_e4m3")
V = Virtual(SHAPES, dtype, code, 8, master=True)
grads = [[(torch.randn(s, device=V.dev) * (1 + r)).to(dtype) for s in SHAPES] for r in range(8)]
for r in range(8):
    V.encode(r, grads[r])
torch.cuda.synchronize()
L = V.L
bad = 0
for r in range(8):
    for i, p in enumerate(V.params):
        s = L.by_id[id(p)]
        enc = code.encode(grads[r][i])
        q = V.wires[r][s.first_tile * V.bpt: s.first_tile * V.bpt + s.numel].view(torch.float8_e4m3fn)
        inv_k = V.scales[r][s.index].item()
        eq = (q.view(torch.uint8) == enc["q"].reshape(-1).view(torch.uint8))
        if not eq.all() or abs(inv_k - enc["inv"].item()) > 0:
            idx = (~eq).nonzero().flatten()[:5]
            print("rank", r, "param", i, "mismatch n=", (~eq).sum().item(), "inv k/o", inv_k, enc["inv"].item())
            g = grads[r][i].reshape(-1).float()
            for j in idx.tolist():
                print("   j", j, "g", g[j].item(), "g/inv", (g[j] / enc["inv"]).item(), "kernel", q[j].float().item(), "oracle", enc["q"].reshape(-1)[j].float().item())
            bad += 1
print("bad", bad)
