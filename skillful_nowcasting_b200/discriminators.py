"""Spatial / temporal discriminators on the B200 path (API of the reference's dgmr/discriminators.py:12-232).

The reference's Python loops over the 8 sampled frames (:201) and the 5 remaining timesteps (:119) become
groups of one launch; BatchNorm1d statistics and spectral-norm iterations stay per frame.
"""
from __future__ import annotations

import torch
import torch.nn as nn
from .hub import HubMixin as PyTorchModelHubMixin   # same API; saves compact copies (see hub.py)

from . import ops
from .common import DBlock
from .layers.core import BatchNorm, SNConv, prefetch_sigmas


def _head(rep, bn: BatchNorm, fc: SNConv, G: int):
    """sum(relu) over H,W -> BatchNorm1d -> SN linear -> sum over the G frames (:209-231, :129-137)."""
    gn = rep.shape[0]
    n = gn // G
    r = ops.sumpool_relu(rep)                                   # [G*N, C]
    r = bn.run(r.reshape(gn, 1, 1, 1, r.shape[1]), G)           # per-frame batch statistics
    r = fc.run(r, G)                                            # [G*N,1,1,1,1]
    return ops.reduce_mid(r.reshape(1, G, n)).reshape(n, 1, 1)  # sum over frames -> [N,1,1]


class SpatialDiscriminator(nn.Module, PyTorchModelHubMixin):
    """ref: dgmr/discriminators.py:141-232."""

    def __init__(self, input_channels: int = 12, num_timesteps: int = 8, num_layers: int = 4, conv_type: str = "standard"):
        super().__init__()
        self.num_timesteps = num_timesteps
        self.input_channels = input_channels
        internal_chn = 24
        self.d1 = DBlock(4 * input_channels, 2 * internal_chn * input_channels, first_relu=False, conv_type=conv_type)
        self.intermediate_dblocks = nn.ModuleList()
        for _ in range(num_layers):
            internal_chn *= 2
            self.intermediate_dblocks.append(
                DBlock(internal_chn * input_channels, 2 * internal_chn * input_channels, conv_type=conv_type))
        self.d6 = DBlock(2 * internal_chn * input_channels, 2 * internal_chn * input_channels, keep_same_output=True,
                         conv_type=conv_type)
        self.fc = SNConv(2 * internal_chn * input_channels, 1, ())
        self.bn = BatchNorm(2 * internal_chn * input_channels)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x: [N,T,C,H,W] -> [N,1,1]."""
        n, t, c, h, w = x.shape
        # same draw as the reference: CPU global RNG, with replacement, over all T frames (:199)
        idxs = torch.randint(low=0, high=t, size=(self.num_timesteps,)).tolist()
        G = len(idxs)
        prefetch_sigmas(self.d1.sn_calls(G) + [c_ for d in self.intermediate_dblocks for c_ in d.sn_calls(G)] + self.d6.sn_calls(G)
                        + [(self.fc, G)])
        rep = ops.gather_frames(x.reshape(n, t, c * h * w), idxs)          # [G*N, C*H*W] frame-major
        rep = ops.nchw_to_cl(rep.reshape(G * n, c, h, w))                  # [G*N,1,H,W,C]
        rep = ops.avg_pool(rep, 1, 2, 2)
        rep = ops.space_to_depth(rep)
        rep = self.d1.run(rep, G)
        for d in self.intermediate_dblocks:
            rep = d.run(rep, G)
        rep = self.d6.run(rep, G)
        return _head(rep, self.bn, self.fc, G)


class TemporalDiscriminator(nn.Module, PyTorchModelHubMixin):
    """ref: dgmr/discriminators.py:47-138."""

    def __init__(self, input_channels: int = 12, num_layers: int = 3, conv_type: str = "standard"):
        super().__init__()
        self.input_channels = input_channels
        internal_chn = 48
        self.d1 = DBlock(4 * input_channels, internal_chn * input_channels, conv_type="3d", first_relu=False)
        self.d2 = DBlock(internal_chn * input_channels, 2 * internal_chn * input_channels, conv_type="3d")
        self.intermediate_dblocks = nn.ModuleList()
        for _ in range(num_layers):
            internal_chn *= 2
            self.intermediate_dblocks.append(
                DBlock(internal_chn * input_channels, 2 * internal_chn * input_channels, conv_type=conv_type))
        self.d_last = DBlock(2 * internal_chn * input_channels, 2 * internal_chn * input_channels, keep_same_output=True,
                             conv_type=conv_type)
        self.fc = SNConv(2 * internal_chn * input_channels, 1, ())
        self.bn = BatchNorm(2 * internal_chn * input_channels)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x: [N,T,C,H,W] -> [N,1,1]."""
        n, t, c, h, w = x.shape
        # [N,T,C,H,W] viewed channels-last with T as depth: [N, D=T, H, W, C]
        if c == 1:
            v = x.reshape(n, t, h, w, 1)
        else:
            v = ops.permute(x.contiguous(), (n, t, h, w, c), (n * t, c, h * w), (c * h * w, h * w, 1), (h * w * c, 1, c))
        t2_ = (t // 2) // 2                # timesteps left after the two 3-D DBlocks
        prefetch_sigmas(self.d1.sn_calls(1) + self.d2.sn_calls(1) + [c_ for d in self.intermediate_dblocks for c_ in d.sn_calls(t2_)]
                        + self.d_last.sn_calls(t2_) + [(self.fc, t2_)])
        v = ops.avg_pool(v, 1, 2, 2)       # AvgPool3d((1,2,2)) (:106)
        v = ops.space_to_depth(v)          # PixelUnshuffle + permute to N,C,T,H,W (:108-110) == [N,T,h,w,4C] here
        v = self.d1.run(v, 1)              # 3-D DBlocks: T 22 -> 11 -> 5
        v = self.d2.run(v, 1)
        _, t2, hh, ww, cc = v.shape
        # per remaining timestep (:119): regroup to timestep-major [T2*N,1,h,w,C]
        rep = ops.permute(v, (t2 * n, 1, hh, ww, cc), (n, t2, hh * ww * cc), (t2 * hh * ww * cc, hh * ww * cc, 1),
                          (hh * ww * cc, n * hh * ww * cc, 1))
        for d in self.intermediate_dblocks:
            rep = d.run(rep, t2)
        rep = self.d_last.run(rep, t2)
        return _head(rep, self.bn, self.fc, t2)


class Discriminator(nn.Module, PyTorchModelHubMixin):
    """ref: dgmr/discriminators.py:12-44: spatial first (fixes RNG order), then temporal; cat on dim 1 -> [N,2,1]."""

    def __init__(self, input_channels: int = 12, num_spatial_frames: int = 8, conv_type: str = "standard"):
        super().__init__()
        self.spatial_discriminator = SpatialDiscriminator(input_channels=input_channels, num_timesteps=num_spatial_frames,
                                                          conv_type=conv_type)
        self.temporal_discriminator = TemporalDiscriminator(input_channels=input_channels, conv_type=conv_type)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        s = self.spatial_discriminator(x)
        t = self.temporal_discriminator(x)
        return ops.concat_channels(s.reshape(-1, 1), t.reshape(-1, 1)).reshape(-1, 2, 1)
