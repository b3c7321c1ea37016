"""Drop-in for the reference's Losses.Loss (Losses.py:12-38): the contrastive pair loss
mean(flag*d + (1-flag)*relu(margin-d)), d = sum_k (a-b)^2, forward AND backward in one CUDA kernel
(dm_contrastive_fwd_bwd), exposed as a torch.autograd.Function so it trains like the original."""
from __future__ import annotations

import torch
from torch import nn

from ._lib import lib
from .raster import _p, _stream


class _Contrastive(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, flag, margin):
        if not (a.is_cuda and b.is_cuda and flag.is_cuda):
            raise ValueError("Loss needs CUDA tensors (there is no CPU path)")
        a32, b32 = a.contiguous().float(), b.contiguous().float()
        f = flag.contiguous().to(torch.int64)
        B, D = a32.shape
        loss = torch.empty(1, dtype=torch.float32, device=a.device)
        ga, gb = torch.empty_like(a32), torch.empty_like(b32)
        L = lib()
        with torch.cuda.device(a.device):
            L.check(L.dm_contrastive_fwd_bwd(_p(a32), _p(b32), _p(f), B, D, float(margin), _p(loss), _p(ga), _p(gb),
                                             _stream()), "dm_contrastive_fwd_bwd")
        ctx.save_for_backward(ga, gb)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        ga, gb = ctx.saved_tensors
        return g * ga, g * gb, None, None


class Loss(nn.Module):
    def __init__(self, margin, lamda, belta):
        super().__init__()
        self.margin, self.lamda, self.belta = margin, lamda, belta

    def forward(self, positive, negative, flag, size_average=True):
        return _Contrastive.apply(positive, negative, flag, self.margin)
