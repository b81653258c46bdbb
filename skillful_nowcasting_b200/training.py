"""Native step driver: the GAN step in the parity-preserving minimal schedule, the fused Adam
optimiser on flat parameter / gradient buffers, and batch-sharded data parallelism (one process per GPU,
NCCL all-reduce of the flat G / D gradient buffers over NVLink; the only collective the path needs, 8e).
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist

from . import _lib, ops
from .losses import loss_hinge_disc_both, loss_hinge_gen


class Adam(torch.optim.Optimizer):
    """torch.optim.Adam semantics (no weight decay / amsgrad; eps 1e-8) on ONE flat fp32 buffer.

    All parameters are re-pointed to views of a single flat tensor and their `.grad`s to views of a flat
    gradient tensor, so `step()` is one fused kernel (`dgmr_adam`) and data parallelism is one
    all-reduce per optimiser step.  Parameters that never receive gradients keep an all-zero gradient, for which
    the Adam update is exactly zero, matching torch's skip of `grad is None`.
    """

    def __init__(self, params: Iterable[torch.Tensor], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 process_group=None):
        params = [p for p in params]
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self._params: List[torch.Tensor] = [p for g in self.param_groups for p in g["params"]]
        dev = self._params[0].device
        al = lambda k: (k + 63) // 64 * 64  # every parameter starts on a 256-byte boundary (vector loads, TMA); pads stay zero
        n = sum(al(p.numel()) for p in self._params)
        self.flat_p = torch.zeros(n, device=dev, dtype=torch.float32)
        self.flat_g = torch.zeros(n, device=dev, dtype=torch.float32)
        self.m = torch.zeros(n, device=dev, dtype=torch.float32)
        self.v = torch.zeros(n, device=dev, dtype=torch.float32)
        off = 0
        with torch.no_grad():
            for p in self._params:
                k = p.numel()
                self.flat_p[off:off + k].copy_(p.detach().reshape(-1))
                p.data = self.flat_p[off:off + k].view(p.shape)
                p.grad = self.flat_g[off:off + k].view(p.shape)
                off += al(k)
        self._al = al
        self.steps = 0
        self.process_group = process_group

    def zero_grad(self, set_to_none: bool = False):  # grads stay views of the flat buffer
        self.flat_g.zero_()
        off = 0
        for p in self._params:
            k = p.numel()
            if p.grad is None or p.grad.data_ptr() != self.flat_g.data_ptr() + 4 * off:
                p.grad = self.flat_g[off:off + k].view(p.shape)
            off += self._al(k)

    @torch.no_grad()
    def step(self, closure=None):
        world = 1
        if dist.is_available() and dist.is_initialized():
            world = dist.get_world_size(self.process_group)
            if world > 1:
                dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM, group=self.process_group)
        self.steps += 1
        g = self.param_groups[0]
        _lib.backend().adam(self.flat_p, self.flat_g, self.m, self.v, g["lr"], g["betas"][0], g["betas"][1], g["eps"],
                            self.steps, 1.0 / world)
        torch.autograd.graph.increment_version(self._params)  # invalidates the packed-weight cache
        return None


def gan_step(generator, discriminator, g_opt, d_opt, images: torch.Tensor, future: torch.Tensor,
             generation_steps: int = 1, grid_lambda: float = 20.0, precip_weight_cap: float = 24.0):
    """One GAN step: 2 discriminator updates + 1 generator update with the losses, gradients and parameter
    updates of the reference's `DGMR.training_step` (dgmr/dgmr.py:137-218), minus its wasted work:
    the generator is not back-propagated in the D phase (its grads are discarded there, :199), no checkpoint
    recompute, no trailing forward, and the D weights take no gradient in the G phase.  In the G phase D still
    sees real||generated in one batch because BatchNorm1d statistics couple the halves."""
    b = images.shape[0]
    real_seq = torch.cat([images, future], dim=1)
    d_loss = None
    for _ in range(2):
        d_opt.zero_grad()
        with torch.no_grad():
            pred = generator(images)
        scores = discriminator(torch.cat([real_seq, torch.cat([images, pred], dim=1)], dim=0))
        d_loss = loss_hinge_disc_both(scores)
        d_loss.backward()
        d_opt.step()
    g_opt.zero_grad()
    d_params = [p for p in discriminator.parameters() if p.requires_grad]
    for p in d_params:
        p.requires_grad_(False)
    try:
        preds = [generator(images) for _ in range(generation_steps)]
        grid = ops.grid_cell(ops.mean_k(preds), future, float(precip_weight_cap))
        gen_scores = []
        for pr in preds:
            s = discriminator(torch.cat([real_seq, torch.cat([images, pr], dim=1)], dim=0))
            gen_scores.append(s[b:])
        g_loss = loss_hinge_gen(torch.cat(gen_scores, dim=0)) + grid_lambda * grid
        g_loss.backward()
    finally:
        for p in d_params:
            p.requires_grad_(True)
    g_opt.step()
    return {"d_loss": d_loss.detach(), "g_loss": g_loss.detach(), "grid_loss": grid.detach()}
