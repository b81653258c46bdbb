"""Latent-stack spatial attention (API of the reference's dgmr/layers/Attention.py:23-85)."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from .core import PlainConv


class AttentionLayer(nn.Module):
    """1x1 q/k/v projections (no bias), un-scaled softmax attention, 1x1 output conv, gamma-scaled residual.

    Reference quirk kept for parity: `forward` passes [C,H,W] tensors to `attention_einsum`, whose
    einsums are written for "h w c" (dgmr/layers/Attention.py:9-20, 79-82), so attention positions are the
    (channel,row) pairs and the contracted axis is the image column.  The kernel implements that."""

    def __init__(self, input_channels: int, output_channels: int, ratio_kq: int = 8, ratio_v: int = 8):
        super().__init__()
        self.ratio_kq, self.ratio_v = ratio_kq, ratio_v
        self.output_channels, self.input_channels = output_channels, input_channels
        self.query = PlainConv(input_channels, output_channels // ratio_kq, (1, 1), bias=False)
        self.key = PlainConv(input_channels, output_channels // ratio_kq, (1, 1), bias=False)
        self.value = PlainConv(input_channels, output_channels // ratio_v, (1, 1), bias=False)
        self.last_conv = PlainConv(output_channels // 8, output_channels, (1, 1), bias=False)
        self.gamma = nn.Parameter(torch.zeros(1))

    def run(self, x):
        q, k, v = self.query.run(x), self.key.run(x), self.value.run(x)
        o = ops.attention(q, k, v)
        scale = self.gamma.view(1, 1).expand(1, self.output_channels).contiguous()
        return self.last_conv.run(o, res=x, scale=scale, exact_dscale=True)  # gamma * conv(o) + x, fused in the conv epilogue

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return ops.cl_to_nchw(self.run(ops.nchw_to_cl(x)))
