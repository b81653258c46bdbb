"""Parameter containers whose ``state_dict()`` layout is identical to the reference's
(`torch.nn.Conv2d/Conv3d/Linear` wrapped by `torch.nn.utils.parametrizations.spectral_norm`,
`torch.nn.BatchNorm2d/1d`), but whose arithmetic runs through the C ABI (ops.py).

State-dict contract: spectrally normalised layers expose
``<layer>.bias``, ``<layer>.parametrizations.weight.original`` (OIHW / OIDHW / [out,in]),
``<layer>.parametrizations.weight.0._u`` and ``..._v``; BatchNorm exposes ``weight, bias,
running_mean, running_var, num_batches_tracked``.  Construction consumes the global RNG in the same
order as the reference constructors, so `torch.manual_seed(s); Module(...)` yields the same tensors.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple, Union

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from ..ops import ACT_NONE, ACT_RELU


def _init_conv_like(weight: torch.Tensor, bias: Optional[torch.Tensor]):
    # same draws as torch.nn.modules.conv._ConvNd.reset_parameters / nn.Linear.reset_parameters
    nn.init.kaiming_uniform_(weight, a=math.sqrt(5))
    if bias is not None:
        fan_in = 1
        for d in weight.shape[1:]:
            fan_in *= d
        bound = 1 / math.sqrt(fan_in) if fan_in > 0 else 0
        nn.init.uniform_(bias, -bound, bound)


class _SNBuffers(nn.Module):
    """`parametrizations.weight.0` of the reference: the power-iteration vectors."""

    def __init__(self, u: torch.Tensor, v: torch.Tensor):
        super().__init__()
        self.register_buffer("_u", u)
        self.register_buffer("_v", v)


class _SNWeight(nn.Module):
    """`parametrizations.weight` of the reference: holds `original` and child `0`."""

    def __init__(self, weight: torch.Tensor, eps: float):
        super().__init__()
        self.original = nn.Parameter(weight)
        with torch.no_grad():
            wm = weight.flatten(1)
            h, w = wm.shape
            # torch/nn/utils/parametrizations.py:430-439: u, v ~ N(0,1) normalised, then 15 warm-up iterations
            u = F.normalize(wm.new_empty(h).normal_(0, 1), dim=0, eps=eps)
            v = F.normalize(wm.new_empty(w).normal_(0, 1), dim=0, eps=eps)
            # (+1: register_parametrization evaluates the parametrization once, in training mode, as its
            #  shape check, which runs one more iteration: torch/nn/utils/parametrize.py)
            for _ in range(16):
                u = F.normalize(torch.mv(wm, v), dim=0, eps=eps)
                v = F.normalize(torch.mv(wm.t(), u), dim=0, eps=eps)
        self.add_module("0", _SNBuffers(u, v))


class SNConv(nn.Module):
    """Spectrally normalised convolution / linear layer (kernel extents 1 or 3, stride 1, same padding)."""

    def __init__(self, in_channels: int, out_channels: int, kernel: Sequence[int] = (1, 1), eps: float = 1e-12):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel = tuple(kernel)
        self.eps = eps
        w = torch.empty((out_channels, in_channels) + self.kernel)
        b = torch.empty(out_channels)
        _init_conv_like(w, b)
        self.bias = nn.Parameter(b)
        self.parametrizations = nn.ModuleDict({"weight": _SNWeight(w, eps)})

    # -- accessors
    @property
    def weight_orig(self) -> torch.Tensor:
        return self.parametrizations["weight"].original

    @property
    def _uv(self) -> Tuple[torch.Tensor, torch.Tensor]:
        b = getattr(self.parametrizations["weight"], "0")
        return b._u, b._v

    def inv_sigma(self, G: int) -> torch.Tensor:
        """1/sigma for the next G reference calls (power iteration advances G times in training).  If the enclosing module
        prefetched it (all of its spectral norms in one launch, see `prefetch_sigmas`), that result is consumed here."""
        pending = self.__dict__.pop("_pending_sigma", None)
        # a prefetched value is only valid for the weights and mode it was computed for (a forward that raised between the
        # prefetch and this call may have left a stale one behind)
        if pending is not None and pending[0] == G and pending[2] == (self.weight_orig._version, self.training, torch.is_grad_enabled()):
            return pending[1]
        u, v = self._uv
        return ops.spectral_inv_sigma(self.weight_orig, u, v, G, self.eps, self.training)

    def scale_of(self, inv_sigma: torch.Tensor) -> torch.Tensor:
        g = inv_sigma.shape[0]
        return inv_sigma.view(g, 1).expand(g, self.out_channels).contiguous()

    def run(self, x, G: int = 1, act: int = ACT_NONE, res=None, res_up2: bool = False, round_out: bool = False, pool=None):
        """x channels-last [N,D,H,W,Cin] -> [N,D,H,W,Cout], one spectral-norm call per group.  res_up2 / round_out / pool: see ops._Conv."""
        scale = self.scale_of(self.inv_sigma(G))
        return ops.conv(x, self.weight_orig, self.bias, scale, res, 0, self.in_channels, G, act, res_up2=res_up2, round_out=round_out,
                        pool=pool)

    def forward(self, x: torch.Tensor) -> torch.Tensor:  # NCHW in / out (drop-in use)
        if len(self.kernel) == 0:  # linear: [N, Cin]
            y = self.run(x.reshape(x.shape[0], 1, 1, 1, x.shape[1]))
            return y.reshape(x.shape[0], self.out_channels)
        keep = len(self.kernel) == 3
        return ops.cl_to_nchw(self.run(ops.nchw_to_cl(x)), keep_depth=keep)


def prefetch_sigmas(calls):
    """calls: [(SNConv, G), ...] that the caller is about to run, each exactly once.  Runs all their power iterations
    (u, v updates included) in ONE kernel launch and parks each 1/sigma on its layer for the upcoming `inv_sigma(G)`."""
    calls = [(l, g) for l, g in calls]
    if len(calls) <= 1:
        return
    training = calls[0][0].training
    assert all(l.training == training for l, _ in calls)
    entries = []
    for l, g in calls:
        l.__dict__.pop("_pending_sigma", None)   # left behind by a forward that did not finish: drop it
        u, v = l._uv
        entries.append((l.weight_orig, u, v, g, l.eps))
    for (l, g), s in zip(calls, ops.spectral_inv_sigma_multi(entries, training)):
        l.__dict__["_pending_sigma"] = (g, s, (l.weight_orig._version, training, torch.is_grad_enabled()))


class PlainConv(nn.Module):
    """Un-normalised Conv2d (latent stack L-blocks and attention 1x1s; ref: dgmr/common.py:263-286)."""

    def __init__(self, in_channels: int, out_channels: int, kernel: Sequence[int] = (1, 1), bias: bool = True):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        w = torch.empty((out_channels, in_channels) + tuple(kernel))
        b = torch.empty(out_channels) if bias else None
        _init_conv_like(w, b)
        self.weight = nn.Parameter(w)
        if bias:
            self.bias = nn.Parameter(b)
        else:
            self.register_parameter("bias", None)

    def run(self, x, act: int = ACT_NONE, res=None, scale=None, exact_dscale=False, round_out: bool = False):
        return ops.conv(x, self.weight, self.bias, scale, res, 0, self.in_channels, 1, act, exact_dscale, round_out=round_out)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return ops.cl_to_nchw(self.run(ops.nchw_to_cl(x)))


class BatchNorm(nn.Module):
    """BatchNorm2d / BatchNorm1d state (eps 1e-5, momentum 0.1, affine, tracked running stats)."""

    def __init__(self, num_features: int, eps: float = 1e-5, momentum: float = 0.1):
        super().__init__()
        self.num_features, self.eps, self.momentum = num_features, eps, momentum
        self.weight = nn.Parameter(torch.ones(num_features))
        self.bias = nn.Parameter(torch.zeros(num_features))
        self.register_buffer("running_mean", torch.zeros(num_features))
        self.register_buffer("running_var", torch.ones(num_features))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))

    def run(self, x, G: int = 1, relu: bool = False, up2: bool = False, conv_only: bool = False, branch: bool = False):
        """branch: returns (y, x_skip), x_skip being the input for its other consumer (residual shortcut): see ops._BatchNorm."""
        if self.training:
            self.num_batches_tracked += G
        return ops.batch_norm(x, self.weight, self.bias, self.running_mean, self.running_var, G, self.training,
                              relu, up2, self.eps, self.momentum, conv_only, branch)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() == 2:
            n, c = x.shape
            return self.run(x.reshape(n, 1, 1, 1, c)).reshape(n, c)
        return ops.cl_to_nchw(self.run(ops.nchw_to_cl(x)))
