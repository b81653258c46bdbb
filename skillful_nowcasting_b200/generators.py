"""Sampler and Generator on the B200 path (API of the reference's dgmr/generators.py:20-212).

The reference runs every block once per forecast step through Python list comprehensions
(generators.py:152-178).  Here the T steps of each non-recurrent block are ONE launch over a
timestep-major batch [T*B, ...] with T groups: BatchNorm batch statistics, running-stat updates and
spectral-norm power iterations stay per timestep, so results are those of
the reference's per-step calls.  Only the ConvGRU recurrence is sequential.
"""
from __future__ import annotations

from typing import List

import torch
import torch.nn as nn
from .hub import HubMixin as PyTorchModelHubMixin   # same API; saves compact copies (see hub.py)

from . import ops
from .common import GBlock, UpsampleGBlock
from .layers.ConvGRU import ConvGRU
from .layers.core import BatchNorm, SNConv, prefetch_sigmas


class Sampler(nn.Module, PyTorchModelHubMixin):
    """ref: dgmr/generators.py:20-182."""

    def __init__(self, forecast_steps: int = 18, latent_channels: int = 768, context_channels: int = 384,
                 output_channels: int = 1):
        super().__init__()
        self.forecast_steps = forecast_steps
        lc, cc = latent_channels, context_channels
        self.convGRU1 = ConvGRU(lc + cc, cc, 3)
        self.gru_conv_1x1 = SNConv(cc, lc, (1, 1))
        self.g1 = GBlock(lc, lc)
        self.up_g1 = UpsampleGBlock(lc, lc // 2)
        self.convGRU2 = ConvGRU(lc // 2 + cc // 2, cc // 2, 3)
        self.gru_conv_1x1_2 = SNConv(cc // 2, lc // 2, (1, 1))
        self.g2 = GBlock(lc // 2, lc // 2)
        self.up_g2 = UpsampleGBlock(lc // 2, lc // 4)
        self.convGRU3 = ConvGRU(lc // 4 + cc // 4, cc // 4, 3)
        self.gru_conv_1x1_3 = SNConv(cc // 4, lc // 4, (1, 1))
        self.g3 = GBlock(lc // 4, lc // 4)
        self.up_g3 = UpsampleGBlock(lc // 4, lc // 8)
        self.convGRU4 = ConvGRU(lc // 8 + cc // 8, cc // 8, 3)
        self.gru_conv_1x1_4 = SNConv(cc // 8, lc // 8, (1, 1))
        self.g4 = GBlock(lc // 8, lc // 8)
        self.up_g4 = UpsampleGBlock(lc // 8, lc // 16)
        self.bn = BatchNorm(lc // 16)
        self.conv_1x1 = SNConv(lc // 16, 4 * output_channels, (1, 1))
        self.output_channels = output_channels

    def run(self, init_states: List[torch.Tensor], latent: torch.Tensor) -> torch.Tensor:
        """init_states: channels-last [B,1,h,w,c], largest first; latent: channels-last [1,1,h,w,c].
        Returns forecasts [B,T,C_out,H,W] (reference layout)."""
        T = self.forecast_steps
        B = init_states[0].shape[0]
        levels = ((self.convGRU1, self.gru_conv_1x1, self.g1, self.up_g1),
                  (self.convGRU2, self.gru_conv_1x1_2, self.g2, self.up_g2),
                  (self.convGRU3, self.gru_conv_1x1_3, self.g3, self.up_g3),
                  (self.convGRU4, self.gru_conv_1x1_4, self.g4, self.up_g4))
        # all 37 spectral norms of the sampler (T power iterations each) in one launch
        calls = [(self.conv_1x1, T)]
        for gru, c11, g, ug in levels:
            cell = gru.cell
            calls += [(cell.read_gate_conv, T), (cell.update_gate_conv, T), (cell.output_conv, T), (c11, T)] + g.sn_calls(T) + ug.sn_calls(T)
        prefetch_sigmas(calls)
        hs = latent
        for lvl, (gru, c11, g, ug) in enumerate(levels):
            # level 0: identical latent input at every step and for every sample (generators.py:146-149)
            # The level input is read by convolutions only: above level 0 it is the previous up-block's output, written tf32-rounded by that
            # block's last epilogue (conv_operand passes it through; unrounded tensors get one private rounded copy that both gate convs share).
            # The recurrence hands back the tf32-rounded copy of its outputs, which the 1x1 conv consumes as is.
            xin = ops.mark_conv_only(hs) if lvl == 0 else ops.conv_operand(hs)
            hs = gru.cell.run_sequence(xin, init_states[3 - lvl], T, shared_input=(lvl == 0), rounded_out=True)
            hs = c11.run(hs if getattr(hs, "_dgmr_tf32", False) else ops.mark_conv_only(hs), T)
            hs = g.run(hs, T)
            hs = ug.run(hs, T, round_out=(lvl < 3))      # levels 0-2: read by the next level's gate convolutions only
        hs = ops.mark_conv_only(self.bn.run(hs, T, relu=True, conv_only=True))
        hs = self.conv_1x1.run(hs, T)  # [T*B,1,h,w,4*Co]
        _, _, h, w, c4 = hs.shape
        co = c4 // 4
        # PixelShuffle(2) + stack on dim 1 (:178,181) in one permute:
        # out[b, t, co, 2h+i, 2w+j] = hs[t*B+b, h, w, co*4 + i*2 + j]
        return ops.permute(hs, (B, T, co, 2 * h, 2 * w), (T, B, h, w, co, 2, 2),
                           (B * h * w * c4, h * w * c4, w * c4, c4, 4, 2, 1),
                           (co * 4 * h * w, T * co * 4 * h * w, 4 * w, 2, 4 * h * w, 2 * w, 1))

    def forward(self, conditioning_states: List[torch.Tensor], latent_dim: torch.Tensor) -> torch.Tensor:
        """NCHW conditioning states (largest first) + latent [1,C,h,w] -> [B,T,C_out,H,W]."""
        return self.run([ops.nchw_to_cl(s) for s in conditioning_states], ops.nchw_to_cl(latent_dim))


class Generator(nn.Module, PyTorchModelHubMixin):
    """ref: dgmr/generators.py:185-212."""

    def __init__(self, conditioning_stack: nn.Module, latent_stack: nn.Module, sampler: nn.Module):
        super().__init__()
        self.conditioning_stack = conditioning_stack
        self.latent_stack = latent_stack
        self.sampler = sampler

    def forward(self, x: torch.Tensor):
        """x: [B,T_in,C,H,W] -> [B,T,C_out,H,W]; context stack, then latent stack, then sampler (RNG order)."""
        if all(hasattr(m, "run") for m in (self.conditioning_stack, self.latent_stack, self.sampler)):
            cond = self.conditioning_stack.run(x)   # stays channels-last between the stacks
            lat = self.latent_stack.run(x)
            return self.sampler.run(list(cond), lat)
        cond = self.conditioning_stack(x)
        lat = self.latent_stack(x)
        return self.sampler(cond, lat)
