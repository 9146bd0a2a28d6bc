"""``FusedMaxPool2d``: 3x3 / stride 2 / pad 1 max pooling on channels-last bf16 in our own kernels
(``csrc/kernels/pool_kernels.cu``); anything else falls through to ``F.max_pool2d``."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ext


class _MaxPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        y, arg = ext.cuda().maxpool_forward(x)
        ctx.save_for_backward(arg)
        ctx.hw = (x.shape[2], x.shape[3])
        return y

    @staticmethod
    def backward(ctx, dy):
        (arg,) = ctx.saved_tensors
        if not dy.is_contiguous(memory_format=torch.channels_last):
            dy = dy.contiguous(memory_format=torch.channels_last)
        return ext.cuda().maxpool_backward(dy, arg, ctx.hw[0], ctx.hw[1])


class FusedMaxPool2d(nn.Module):
    def __init__(self, kernel_size=3, stride=2, padding=1):
        super().__init__()
        self.kernel_size, self.stride, self.padding = kernel_size, stride, padding

    def forward(self, x):
        if ((self.kernel_size, self.stride, self.padding) == (3, 2, 1) and x.is_cuda and x.dtype == torch.bfloat16
                and x.dim() == 4 and x.shape[1] % 8 == 0 and x.is_contiguous(memory_format=torch.channels_last)):
            return _MaxPool.apply(x)
        return F.max_pool2d(x, self.kernel_size, self.stride, self.padding)
