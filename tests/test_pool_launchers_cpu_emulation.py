"""``pool_kernels.cu`` — max-pool forward / backward (row / quad kernels AND the general fallbacks), the uint8 → bf16 input
normalisers and the stem's im2col — executed on the CPU **through their real launchers** (grid sizing, kernel selection by
shape): the whole translation unit is compiled with g++ against the fiber emulator (``tests/_cuda_emu.py``) with its
``<<< >>>`` launches rewritten, and compared with ``F.max_pool2d`` + autograd / plain torch expressions."""
import ctypes
import shutil

import pytest
import torch
import torch.nn.functional as F

from tests import _cuda_emu


@pytest.fixture(autouse=True)
def _single_threaded_torch():
    n = torch.get_num_threads()
    torch.set_num_threads(1)
    yield
    torch.set_num_threads(n)


DRIVER = r'''
extern "C" void emu_pool_fwd(const void* x, void* y, void* arg, int N, int H, int W, int C) { psb_maxpool3x3s2_forward(0, x, y, arg, N, H, W, C); }
extern "C" void emu_pool_bwd(const void* dy, const void* arg, void* dx, int N, int H, int W, int C) { psb_maxpool3x3s2_backward(0, dy, arg, dx, N, H, W, C); }
extern "C" void emu_norm8(const void* x, void* y, const float* m, const float* s, int N, long long HW) { psb_normalize_pad8_launch(0, x, y, m, s, N, HW); }
extern "C" void emu_norm3(const void* x, void* y, const float* m, const float* s, int N, long long HW) { psb_normalize_nhwc3_launch(0, x, y, m, s, N, HW); }
extern "C" void emu_im2col(const void* x, void* a, int N, int H, int W) { psb_im2col_stem_launch(0, x, a, N, H, W); }
extern "C" void emu_stem_finalize(const float* partial, int grid, void* out) { psb_stem_wgrad_finalize_launch(0, partial, grid, out); }
'''


@pytest.fixture(scope="module")
def lib():
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    return _cuda_emu.compile_shared(_cuda_emu.pool_source() + DRIVER, "psb_emu_pool_")


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


@pytest.fixture(params=[148, 1])
def sms(lib, request):
    """148 SMs, or ONE: then the launchers' grid caps (16 CTAs per SM) bind and every kernel walks its grid-stride loop several
    times per CTA — what happens on the real ResNet-18 tensors."""
    lib.emu_set_sm_count(request.param)
    yield request.param
    lib.emu_set_sm_count(148)


@pytest.mark.parametrize("shape", [(2, 64, 12, 16), (1, 16, 6, 4), (2, 24, 8, 8), (1, 64, 7, 9), (2, 8, 5, 6), (1, 512, 2, 2),
                                   (5, 64, 16, 8), (3, 8, 22, 40)])
def test_maxpool_launchers_match_torch(lib, sms, shape):
    """Even H, W with 256 % (C/8) == 0 → the row / quad kernels; anything else → the general ones: same answer as ATen, ties
    included (the first maximum wins)."""
    N, C, H, W = shape
    g = torch.Generator().manual_seed(H * 100 + W)
    x = torch.randn(N, C, H, W, generator=g).bfloat16()
    x[:, :, ::2, ::3] = x[:, :, 1:2, 1:2]                                  # plant ties inside windows
    xn = x.permute(0, 2, 3, 1).contiguous()                                # NHWC as the op sees it
    OH, OW = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    y = torch.empty(N, OH, OW, C, dtype=torch.bfloat16)
    arg = torch.empty(N, OH, OW, C, dtype=torch.uint8)
    lib.emu_pool_fwd(_p(xn), _p(y), _p(arg), N, H, W, C)
    xt = x.float().requires_grad_(True)
    yt = F.max_pool2d(xt, 3, 2, 1)
    assert torch.equal(y.permute(0, 3, 1, 2).float(), yt.detach())
    dy = torch.randn(N, OH, OW, C, generator=g).bfloat16()
    dx = torch.full((N, H, W, C), float("nan"), dtype=torch.bfloat16)
    lib.emu_pool_bwd(_p(dy), _p(arg), _p(dx), N, H, W, C)
    yt.backward(dy.permute(0, 3, 1, 2).float())
    # ATen accumulates overlapping windows in fp32 and so does the kernel (then one rounding to bf16)
    assert torch.equal(dx.permute(0, 3, 1, 2).float(), xt.grad.bfloat16().float())


def test_input_normalisers_match_torch(lib, sms):
    N, H, W = 3, 37, 41
    g = torch.Generator().manual_seed(0)
    x = torch.randint(0, 256, (N, 3, H, W), dtype=torch.uint8, generator=g)
    mean = torch.tensor([123.675, 116.28, 103.53])
    inv = 1.0 / torch.tensor([58.395, 57.12, 57.375])
    want = ((x.float() - mean.view(1, 3, 1, 1)) * inv.view(1, 3, 1, 1)).bfloat16().permute(0, 2, 3, 1)      # NHWC
    y8 = torch.full((N, H, W, 8), 7.0, dtype=torch.bfloat16)
    lib.emu_norm8(_p(x), _p(y8), _p(mean), _p(inv), N, ctypes.c_longlong(H * W))
    assert torch.equal(y8[..., :3], want) and float(y8[..., 3:].abs().max()) == 0.0
    y3 = torch.empty(N, H, W, 3, dtype=torch.bfloat16)
    lib.emu_norm3(_p(x), _p(y3), _p(mean), _p(inv), N, ctypes.c_longlong(H * W))
    assert torch.equal(y3, want)


@pytest.mark.parametrize("shape", [(2, 16, 24), (1, 9, 10), (1, 32, 8)])
def test_stem_im2col_matches_unfold(lib, sms, shape):
    """The [N*OH*OW, 176] patch matrix of the 7x7 / stride-2 / pad-3 stem: per kernel row 7 x 3 = 21 values padded to 24, then 8
    zero columns — interior pixels take the aligned 4-byte-load path, border pixels the element-wise one."""
    N, H, W = shape
    g = torch.Generator().manual_seed(1)
    x = torch.randn(N, 3, H, W, generator=g).bfloat16()
    xn = x.permute(0, 2, 3, 1).contiguous()
    OH, OW = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    a = torch.full((N * OH * OW, 176), 9.0, dtype=torch.bfloat16)
    lib.emu_im2col(_p(xn), _p(a), N, H, W)
    cols = F.unfold(x.float(), 7, padding=3, stride=2)                     # [N, 3*49, L], rows ordered (c, kh, kw)
    cols = cols.view(N, 3, 7, 7, OH * OW).permute(0, 4, 2, 3, 1).reshape(N * OH * OW, 7, 21)
    want = F.pad(F.pad(cols, (0, 3)).reshape(-1, 168), (0, 8)).bfloat16()
    assert torch.equal(a, want)


@pytest.mark.parametrize("grid", [1, 5, 148])
def test_stem_wgrad_finalize_sums_partials_in_order(lib, grid):
    """``psb_stem_wgrad_finalize_kernel``: dW2d[co][k] = Σ_cta partial[cta][k][co] in CTA order (deterministic), cast to bf16 — the
    [64,176] GEMM-layout gradient the engine receives straight in the wire arena."""
    g = torch.Generator().manual_seed(grid)
    partial = torch.randn(grid, 176, 64, generator=g)
    out = torch.full((64, 176), float("nan"), dtype=torch.bfloat16)
    lib.emu_stem_finalize(_p(partial), grid, _p(out))
    acc = torch.zeros(176, 64)
    for c in range(grid):                                                   # the kernel's summation order
        acc += partial[c]
    assert torch.equal(out, acc.t().bfloat16())
