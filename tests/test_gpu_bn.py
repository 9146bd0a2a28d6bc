"""Fused channels-last bf16 BatchNorm(+residual)(+ReLU) kernels vs a plain PyTorch fp32 reference."""
import pytest
import torch
import torch.nn.functional as F

from pytorch_ps_mpi_b200.ops.batchnorm import FusedBatchNormAct2d

pytestmark = pytest.mark.gpu


def _ref(x, res, w, b, eps, relu):
    xf = x.float()
    mean = xf.mean((0, 2, 3), keepdim=True)
    var = xf.var((0, 2, 3), unbiased=False, keepdim=True)
    y = (xf - mean) * torch.rsqrt(var + eps) * w.float().view(1, -1, 1, 1) + b.float().view(1, -1, 1, 1)
    if res is not None:
        y = y + res.float()
    return F.relu(y) if relu else y


@pytest.mark.parametrize("shape", [(8, 64, 28, 28), (4, 512, 7, 7), (3, 24, 5, 9), (2, 192, 14, 14), (16, 128, 16, 16)])
@pytest.mark.parametrize("relu,has_res", [(False, False), (True, False), (True, True), (False, True)])
def test_bn_forward_backward(shape, relu, has_res):
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    N, C, H, W = shape
    bn = FusedBatchNormAct2d(C, relu=relu).to(dev).bfloat16()
    with torch.no_grad():
        bn.weight.copy_(torch.rand(C) + 0.5)
        bn.bias.copy_(torch.randn(C) * 0.2)
    assert bn.running_mean.dtype == torch.float32
    x = (torch.randn(shape, device=dev) * 2 + 0.5).bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    res = torch.randn(shape, device=dev).bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True) if has_res else None
    y = bn(x, res)
    assert y.is_contiguous(memory_format=torch.channels_last) and y.dtype == torch.bfloat16
    # fp32 reference through autograd
    xr = x.detach().float().requires_grad_(True)
    rr = res.detach().float().requires_grad_(True) if has_res else None
    wr = bn.weight.detach().float().requires_grad_(True)
    br = bn.bias.detach().float().requires_grad_(True)
    z = _ref(xr, rr, wr, br, bn.eps, False)
    yr = F.relu(z) if relu else z
    assert torch.allclose(y.float(), yr, rtol=2e-2, atol=3e-2), (y.float() - yr).abs().max()
    if relu:   # backward with the kernel's own mask: bf16 rounding may put an output on the other side of 0
        yr = z * (y.detach().float() > 0).float()
    g = torch.randn_like(yr)
    yr.backward(g)
    y.backward(g.bfloat16().contiguous(memory_format=torch.channels_last))
    scale = max(1.0, xr.grad.abs().max().item())
    assert torch.allclose(x.grad.float(), xr.grad, rtol=5e-2, atol=5e-2 * scale), (x.grad.float() - xr.grad).abs().max()
    if has_res:
        assert torch.allclose(res.grad.float(), rr.grad, rtol=2e-2, atol=2e-2)
    gs = max(1.0, wr.grad.abs().max().item())
    assert torch.allclose(bn.weight.grad.float(), wr.grad, rtol=3e-2, atol=3e-2 * gs)
    assert torch.allclose(bn.bias.grad.float(), br.grad, rtol=3e-2, atol=3e-2 * gs)
    # running statistics follow nn.BatchNorm2d's rule
    ref_bn = torch.nn.BatchNorm2d(C).to(dev)
    ref_bn(x.detach().float())
    assert torch.allclose(bn.running_mean, ref_bn.running_mean, rtol=1e-2, atol=1e-2)
    assert torch.allclose(bn.running_var, ref_bn.running_var, rtol=1e-2, atol=1e-2)


def test_bn_eval_mode_and_fallback():
    dev = torch.device("cuda", 0)
    bn = FusedBatchNormAct2d(32, relu=True).to(dev).bfloat16()
    x = torch.randn(4, 32, 6, 6, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
    bn.train()
    bn(x)
    bn.eval()
    y = bn(x)
    ref = F.relu(F.batch_norm(x.float(), bn.running_mean, bn.running_var, bn.weight.float(), bn.bias.float(), False, 0.1, bn.eps))
    assert torch.allclose(y.float(), ref, rtol=2e-2, atol=2e-2)
    # NCHW-contiguous input takes the stock path with the same semantics
    y2 = bn(x.contiguous())
    assert torch.allclose(y2.float(), ref, rtol=2e-2, atol=2e-2)


def test_resnet18_fused_matches_stock_math():
    from pytorch_ps_mpi_b200 import models
    import torchvision
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    m = models.resnet18(num_classes=10).to(dev).to(memory_format=torch.channels_last).bfloat16()
    tv = torchvision.models.resnet18(num_classes=10).to(dev).to(memory_format=torch.channels_last)
    tv.load_state_dict({k: v.float() for k, v in m.state_dict().items()})
    x = torch.randn(8, 3, 64, 64, device=dev)
    y = m(x.bfloat16().contiguous(memory_format=torch.channels_last))
    yr = tv(x.contiguous(memory_format=torch.channels_last))
    assert torch.allclose(y.float(), yr, rtol=0.1, atol=0.15), (y.float() - yr).abs().max()


@pytest.mark.parametrize("shape", [(4, 64, 112, 112), (3, 16, 7, 9), (2, 8, 1, 1)])
def test_maxpool_matches_torch(shape):
    from pytorch_ps_mpi_b200.ops.pooling import FusedMaxPool2d
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    x = torch.randn(shape, device=dev).bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    xr = x.detach().float().requires_grad_(True)
    y = FusedMaxPool2d()(x)
    yr = F.max_pool2d(xr, 3, 2, 1)
    assert y.shape == yr.shape and torch.equal(y.float(), yr)
    g = torch.randn_like(yr)
    y.backward(g.bfloat16().contiguous(memory_format=torch.channels_last))
    yr.backward(g.bfloat16().float())
    assert torch.allclose(x.grad.float(), xr.grad, rtol=1e-2, atol=1e-2)


def test_normalize_pad8_and_padded_stem():
    from pytorch_ps_mpi_b200.ops.preprocess import normalize_pad8, IMAGENET_MEAN, IMAGENET_STD
    from pytorch_ps_mpi_b200 import models
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    x = torch.randint(0, 256, (4, 3, 32, 40), dtype=torch.uint8, device=dev)
    y = normalize_pad8(x)
    assert y.shape == (4, 8, 32, 40) and y.dtype == torch.bfloat16 and y.is_contiguous(memory_format=torch.channels_last)
    m = torch.tensor(IMAGENET_MEAN, device=dev).view(1, 3, 1, 1)
    s = torch.tensor(IMAGENET_STD, device=dev).view(1, 3, 1, 1)
    ref = (x.float() - m) / s
    assert torch.allclose(y[:, :3].float(), ref, rtol=1e-2, atol=1e-2) and float(y[:, 3:].abs().max()) == 0.0
    net = models.resnet18(num_classes=10).to(dev).to(memory_format=torch.channels_last).bfloat16()
    a = net.stem(y)                                             # 8-channel path, zero-padded weight
    b = net.stem(ref.to(torch.bfloat16).contiguous(memory_format=torch.channels_last))   # stock 3-channel path
    assert torch.allclose(a.float(), b.float(), rtol=3e-2, atol=3e-2)
    a.float().square().mean().backward()
    assert net.conv1.weight.grad is not None and net.conv1.weight.grad.shape == net.conv1.weight.shape


def test_gemm_stem_matches_conv2d():
    from pytorch_ps_mpi_b200.ops.stem import stem_conv, stem_supported
    from pytorch_ps_mpi_b200.ops.preprocess import normalize_nhwc, IMAGENET_MEAN, IMAGENET_STD
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    raw = torch.randint(0, 256, (6, 3, 64, 96), dtype=torch.uint8, device=dev)
    x = normalize_nhwc(raw)
    m = torch.tensor(IMAGENET_MEAN, device=dev).view(1, 3, 1, 1)
    s = torch.tensor(IMAGENET_STD, device=dev).view(1, 3, 1, 1)
    assert torch.allclose(x.float(), (raw.float() - m) / s, rtol=1e-2, atol=1e-2)
    conv = torch.nn.Conv2d(3, 64, 7, 2, 3, bias=False).to(dev).to(memory_format=torch.channels_last).bfloat16()
    assert stem_supported(x, conv)
    y = stem_conv(x, conv.weight)
    ref = F.conv2d(x.float(), conv.weight.float(), None, 2, 3)
    assert y.shape == ref.shape and y.is_contiguous(memory_format=torch.channels_last)
    assert torch.allclose(y.float(), ref, rtol=3e-2, atol=3e-2), (y.float() - ref).abs().max()
    g = torch.randn_like(ref)
    (gw,) = torch.autograd.grad(y, [conv.weight], g.bfloat16())
    wref = conv.weight.detach().float().requires_grad_(True)
    (gr,) = torch.autograd.grad(F.conv2d(x.float(), wref, None, 2, 3), [wref], g.bfloat16().float())
    assert gw.shape == conv.weight.shape
    assert torch.allclose(gw.float(), gr, rtol=5e-2, atol=5e-2 * gr.abs().max().item())


def test_num_batches_tracked_is_counted_lazily():
    """The kernel path counts training batches on the host and writes the buffer when it is looked at (state_dict)."""
    import torch
    from pytorch_ps_mpi_b200.ops.batchnorm import FusedBatchNormAct2d
    dev = torch.device("cuda", 0)
    bn = FusedBatchNormAct2d(16, relu=True).to(dev).bfloat16().train()
    x = torch.randn(4, 16, 8, 8, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
    for _ in range(3):
        bn(x)
    assert int(bn.state_dict()["num_batches_tracked"]) == 3
    bn(x)
    assert int(bn.state_dict()["num_batches_tracked"]) == 4
