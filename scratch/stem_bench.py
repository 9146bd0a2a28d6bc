import torch, time
torch.backends.cudnn.benchmark = True
dev = torch.device("cuda", 0)
def bench(C, fmt, steps=20):
    conv = torch.nn.Conv2d(C, 64, 7, 2, 3, bias=False).to(dev).bfloat16()
    x = torch.randn(256, C, 224, 224, device=dev).bfloat16()
    if fmt == "cl":
        conv = conv.to(memory_format=torch.channels_last); x = x.contiguous(memory_format=torch.channels_last)
    g = None
    for i in range(steps + 5):
        if i == 5:
            torch.cuda.synchronize(); s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True); s.record()
        y = conv(x)
        if g is None: g = torch.randn_like(y)
        conv.weight.grad = None
        y.backward(g)
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / steps
for C in (3, 4, 8):
    for fmt in ("cl", "nchw"):
        try:
            print(C, fmt, f"{bench(C, fmt):.3f} ms fwd+wgrad", flush=True)
        except Exception as ex:
            print(C, fmt, "err", ex)
