"""CPU model of the row-based max-pool forward and the 2x2-QUAD max-pool backward of ``csrc/kernels/pool_kernels.cu``
(``psb_maxpool_fwd_rows`` / ``psb_maxpool_bwd_quads``): the same index arithmetic and the same tap partition, transcribed
line by line, checked against ``F.max_pool2d`` + autograd.  (The GPU tests compare the real kernels; this one pins the
decomposition itself — which window taps fall into which pixel of a quad — on every CPU run.)"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F


def model_forward(x):
    """x [N,H,W,C] float → (y [N,OH,OW,C], arg [N,OH,OW,C] uint8); 3x3 / stride 2 / pad 1, first maximum wins."""
    N, H, W, C = x.shape
    OH, OW = (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1
    y = np.full((N, OH, OW, C), -np.inf, dtype=np.float64)
    arg = np.full((N, OH, OW, C), 255, dtype=np.uint8)
    for row in range(N * OH):                         # one CTA walks whole output rows
        n, oh = divmod(row, OH)
        h0 = oh * 2 - 1
        for ow in range(OW):
            w0 = ow * 2 - 1
            for kh in range(3):
                h = h0 + kh
                if h < 0 or h >= H:
                    continue
                for kw in range(3):
                    w = w0 + kw
                    if w < 0 or w >= W:
                        continue
                    v = x[n, h, w]
                    better = v > y[n, oh, ow]
                    y[n, oh, ow][better] = v[better]
                    arg[n, oh, ow][better] = kh * 3 + kw
    return y, arg


def model_backward_quads(dy, arg, H, W):
    """The quad (h0..h0+1, w0..w0+1), h0 / w0 even, is covered by the windows (oh0+a, ow0+b), a, b in {0,1}; the taps that fall
    into the quad: window (0,0): 4,5,7,8 → pixels 0,1,2,3;  (0,1): 3,6 → 1,3;  (1,0): 1,2 → 2,3;  (1,1): 0 → 3."""
    N, OH, OW, C = dy.shape
    assert H % 2 == 0 and W % 2 == 0
    dx = np.zeros((N, H, W, C), dtype=np.float64)
    hp, wq = H // 2, W // 2
    for row in range(N * hp):
        n, oh0 = divmod(row, hp)
        for qd in range(wq):
            t = np.full((4, C), 255, dtype=np.int64)
            f = np.zeros((4, C))
            for a in range(2):
                for b in range(2):
                    if oh0 + a < OH and qd + b < OW:
                        t[a * 2 + b] = arg[n, oh0 + a, qd + b]
                        f[a * 2 + b] = dy[n, oh0 + a, qd + b]
            d = [np.where(t[0] == 4, f[0], 0.0),
                 np.where(t[0] == 5, f[0], 0.0) + np.where(t[1] == 3, f[1], 0.0),
                 np.where(t[0] == 7, f[0], 0.0) + np.where(t[2] == 1, f[2], 0.0),
                 np.where(t[0] == 8, f[0], 0.0) + np.where(t[1] == 6, f[1], 0.0) + np.where(t[2] == 2, f[2], 0.0)
                 + np.where(t[3] == 0, f[3], 0.0)]
            for p in range(4):
                dx[n, 2 * oh0 + (p >> 1), 2 * qd + (p & 1)] = d[p]
    return dx


@pytest.mark.parametrize("shape", [(2, 8, 8, 3), (1, 6, 10, 2), (3, 2, 2, 1), (1, 12, 4, 5)])
def test_quad_backward_and_row_forward_match_torch(shape):
    N, H, W, C = shape
    rng = np.random.default_rng(0)
    x = rng.standard_normal((N, H, W, C))
    y, arg = model_forward(x)
    xt = torch.tensor(x).permute(0, 3, 1, 2).clone().requires_grad_(True)
    yt = F.max_pool2d(xt, 3, 2, 1)
    assert np.allclose(y, yt.detach().permute(0, 2, 3, 1).numpy())
    dy = rng.standard_normal(y.shape)
    yt.backward(torch.tensor(dy).permute(0, 3, 1, 2))
    dx = model_backward_quads(dy, arg, H, W)
    assert np.allclose(dx, xt.grad.permute(0, 2, 3, 1).numpy())
    # every output's gradient lands on exactly one input pixel
    assert np.isclose(dx.sum(), dy.sum())
