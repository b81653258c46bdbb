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

    def _offsets(self):
        off = 0
        for p in self._params:
            yield p, off, p.numel()
            off += self._al(p.numel())

    def zero_grad(self, set_to_none: bool = False):  # grads stay views of the flat buffer
        self.flat_g.zero_()
        self._reattach(copy=False)

    def _reattach(self, copy: bool):
        """Parameters and their gradients must alias the flat buffers, or the fused update silently works on stale data:
        `module.zero_grad()` (set_to_none) makes autograd write the next gradient into a fresh tensor, `module.to()/.cuda()`
        re-points `p.data`.  Re-point them; with `copy` the detached values are copied in first (they are the live ones)."""
        gbase, pbase = self.flat_g.data_ptr(), self.flat_p.data_ptr()
        for p, off, k in self._offsets():
            if p.data_ptr() != pbase + 4 * off:
                if p.device != self.flat_p.device:
                    raise RuntimeError("dgmr Adam: a parameter was moved to another device after the optimiser was built; "
                                       "build the optimiser after .to()/.cuda()")
                if copy:
                    self.flat_p[off:off + k].copy_(p.detach().reshape(-1))
                p.data = self.flat_p[off:off + k].view(p.shape)
            if p.grad is None:
                if copy:
                    self.flat_g[off:off + k].zero_()
                p.grad = self.flat_g[off:off + k].view(p.shape)
            elif p.grad.data_ptr() != gbase + 4 * off:
                if copy:
                    self.flat_g[off:off + k].copy_(p.grad.detach().reshape(-1))
                p.grad = self.flat_g[off:off + k].view(p.shape)

    @torch.no_grad()
    def step(self, closure=None):
        if len(self.param_groups) != 1:
            raise RuntimeError("dgmr Adam: exactly one parameter group is supported (one flat buffer, one fused launch)")
        self._reattach(copy=True)
        world = 1
        if dist.is_available() and dist.is_initialized():
            world = dist.get_world_size(self.process_group)
            if world > 1:
                dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM, group=self.process_group)
        self.steps += 1
        g = self.param_groups[0]
        _lib.backend().adam(self.flat_p, self.flat_g, self.m, self.v, g["lr"], g["betas"][0], g["betas"][1], g["eps"],
                            self.steps, 1.0 / world)
        torch.autograd.graph.increment_version(self._params)  # invalidates the packed-weight cache ...
        ops.refresh_packs(self._params)                       # ... which is refreshed in place, all weights in one launch
        return None

    # ---- checkpointing: the layout of torch.optim.Adam's state dict (exp_avg / exp_avg_sq / step per parameter index),
    # so Lightning checkpoints carry the moments and reference (torch.optim.Adam) checkpoints load here and vice versa
    def state_dict(self):
        sd = super().state_dict()
        state = {}
        if self.steps > 0:
            for i, (p, off, k) in enumerate(self._offsets()):
                state[i] = dict(step=torch.tensor(float(self.steps)), exp_avg=self.m[off:off + k].view(p.shape).clone(),
                                exp_avg_sq=self.v[off:off + k].view(p.shape).clone())
        sd["state"] = state
        return sd

    @torch.no_grad()
    def load_state_dict(self, state_dict):
        state = state_dict.get("state", {})
        groups = state_dict.get("param_groups")
        if groups:
            if len(groups) != 1:
                raise RuntimeError("dgmr Adam: exactly one parameter group is supported")
            for key in ("lr", "betas", "eps"):
                if key in groups[0]:
                    self.param_groups[0][key] = groups[0][key]
        self.m.zero_(); self.v.zero_()
        steps = 0
        for i, (p, off, k) in enumerate(self._offsets()):
            st = state.get(i, state.get(str(i)))
            if st is None:
                continue
            self.m[off:off + k].copy_(st["exp_avg"].reshape(-1))
            self.v[off:off + k].copy_(st["exp_avg_sq"].reshape(-1))
            steps = max(steps, int(float(st["step"])))
        self.steps = steps


def gan_step(generator, discriminator, g_opt, d_opt, images: torch.Tensor, future: torch.Tensor,
             generation_steps: int = 1, grid_lambda: float = 20.0, precip_weight_cap: float = 24.0, d_phase_generator=None):
    """One GAN step: 2 discriminator updates + 1 generator update with the losses, gradients and parameter
    updates of the reference's `DGMR.training_step` (dgmr/dgmr.py:137-218), minus its wasted work:
    the generator is not back-propagated in the D phase (its grads are discarded there, :199), no checkpoint
    recompute, no trailing forward, and the D weights take no gradient in the G phase.  In the G phase D still
    sees real||generated in one batch because BatchNorm1d statistics couple the halves.
    d_phase_generator: optional `inference.GraphedGenerator(generator, images, train_mode=True)`: the two gradient-free generator forwards of
    the D phase replayed from a CUDA graph (same arithmetic and state updates; its output buffer is consumed before the next replay)."""
    b = images.shape[0]
    real_seq = torch.cat([images, future], dim=1)
    d_loss = None
    for _ in range(2):
        d_opt.zero_grad()
        if d_phase_generator is not None:
            pred = d_phase_generator(images)
        else:
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
