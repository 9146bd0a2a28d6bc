"""Input pre-processing fused into one kernel: uint8 NCHW image batch → normalised bf16 NHWC with
the channel dimension zero-padded 3 → 8 (``csrc/kernels/pool_kernels.cu::psb_normalize_pad8``)."""
from __future__ import annotations

from typing import Sequence

import torch

from . import ext

IMAGENET_MEAN = (0.485 * 255, 0.456 * 255, 0.406 * 255)
IMAGENET_STD = (0.229 * 255, 0.224 * 255, 0.225 * 255)


def normalize_pad8(x: torch.Tensor, mean: Sequence[float] = IMAGENET_MEAN, std: Sequence[float] = IMAGENET_STD) -> torch.Tensor:
    """``[N,3,H,W] uint8`` → ``[N,8,H,W] bf16`` channels-last, ``(x - mean) / std`` in channels 0..2."""
    if x.is_cuda and x.dtype == torch.uint8 and x.dim() == 4 and x.shape[1] == 3 and x.is_contiguous():
        return ext.cuda().normalize_pad8(x, list(mean), list(std))
    m = torch.tensor(mean, device=x.device, dtype=torch.float32).view(1, 3, 1, 1)
    s = torch.tensor(std, device=x.device, dtype=torch.float32).view(1, 3, 1, 1)
    y = ((x.float() - m) / s).to(torch.bfloat16)
    y = torch.nn.functional.pad(y, (0, 0, 0, 0, 0, 5))
    return y.contiguous(memory_format=torch.channels_last)


def normalize_nhwc(x: torch.Tensor, mean: Sequence[float] = IMAGENET_MEAN, std: Sequence[float] = IMAGENET_STD) -> torch.Tensor:
    """``[N,3,H,W] uint8`` → ``[N,3,H,W] bf16`` channels-last, ``(x - mean) / std`` — one kernel."""
    if x.is_cuda and x.dtype == torch.uint8 and x.dim() == 4 and x.shape[1] == 3 and x.is_contiguous():
        return ext.cuda().normalize_nhwc3(x, list(mean), list(std))
    m = torch.tensor(mean, device=x.device, dtype=torch.float32).view(1, 3, 1, 1)
    s = torch.tensor(std, device=x.device, dtype=torch.float32).view(1, 3, 1, 1)
    return ((x.float() - m) / s).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
