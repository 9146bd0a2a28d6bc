import os, sys, torch
sys.path.insert(0, '.')
sys.argv = ['bench.py']
import importlib.util
spec = importlib.util.spec_from_file_location('bench_main', 'bench.py'); bench = importlib.util.module_from_spec(spec); spec.loader.exec_module(bench)
import pytorch_ps_mpi_b200 as ps
args = bench.parse()
w = ps.runtime.init(); device = w.device
torch.backends.cudnn.benchmark = True
model, make_batch, loss_fn, cfg = bench.build(args, device, ps)
gen0 = torch.Generator().manual_seed(7)
xb, yb = (t.to(device) for t in make_batch(gen0))
for _ in range(2):
    loss_fn(xb, yb).backward(); model.zero_grad(set_to_none=True)
named = list(model.named_parameters())
opt = ps.SGD(named, [p for _, p in named], lr=0.05, momentum=0.9, weight_decay=1e-4, code=ps.Identity(), engine="device", average=True)
gen = torch.Generator().manual_seed(1234)
dev = [tuple(t.to(device) for t in make_batch(gen)) for _ in range(4)]
evs = [torch.cuda.Event(enable_timing=True) for _ in range(41)]
torch.cuda.synchronize()
evs[0].record()
for i in range(40):
    x, y = dev[i % 4]
    opt.zero_grad(set_to_none=True)
    loss_fn(x, y).backward()
    opt.step()
    evs[i + 1].record()
torch.cuda.synchronize()
print("per-step ms:", " ".join(f"{evs[i].elapsed_time(evs[i+1]):.2f}" for i in range(40)))
print("mem reserved MB", torch.cuda.memory_reserved() >> 20)
opt.close()
