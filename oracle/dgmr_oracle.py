"""CPU oracle for the DGMR generator / discriminator training step.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(`skillful_nowcasting_b200/`) may import this file; only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
leg use it, and only as the checker / baseline.

What it is: a *functional* restatement (plain functions over a flat
``state`` dict, no nn.Module) of the arithmetic the reference performs on the
hot path, written against torch fp32 primitives because that is where the
reference's arithmetic lives (`torch.nn.functional.conv2d/conv3d`,
`batch_norm`, `avg_pool`, the `spectral_norm` parametrization
(torch/nn/utils/parametrizations.py:403-527), einops rearranges).  Each function
cites the reference file:line it follows (paths relative to /root/reference).

Pinning: the reference ships no golden vectors for this path (SURVEY.md §8c:
"parity unpinned" by the reference's own tests).  This oracle is pinned instead
against outputs of the reference itself, run in the build container:
`tests/golden/make_golden.py` imports /root/reference, dumps small fixtures, and
`tests/test_oracle.py` checks oracle == fixture (and, when /root/reference is
present, oracle == live reference modules, forward, buffers and gradients).

The ``state`` dict uses exactly the reference's ``state_dict()`` key names, so the
same tensors can be loaded into the reference modules, the oracle and the B200
modules.  In training mode the oracle mutates the buffers in ``state`` in place
(`_u`, `_v`, BN running stats, `num_batches_tracked`) like the reference does.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

State = Dict[str, torch.Tensor]

SN_EPS_G = 1e-4  # GBlock / UpsampleGBlock / ConvGRU (dgmr/common.py:25,49; dgmr/layers/ConvGRU.py:16)
SN_EPS_DEFAULT = 1e-12  # every other spectral_norm call (torch default)
BN_EPS = 1e-5
BN_MOMENTUM = 0.1


# --------------------------------------------------------------------------- helpers
def _normalize(x: torch.Tensor, eps: float) -> torch.Tensor:
    # F.normalize(dim=0): x / max(||x||_2, eps)
    return x / x.norm().clamp_min(eps)


def sn_weight(state: State, prefix: str, training: bool, eps: float) -> torch.Tensor:
    """Spectrally normalised weight of layer ``prefix``.

    Follows torch/nn/utils/parametrizations.py:495-527 (the parametrization the
    reference applies at e.g. dgmr/layers/ConvGRU.py:29-55): in training one power
    iteration updates ``_u``/``_v`` in place under no_grad, then
    ``sigma = u . (W_mat v)`` with u, v treated as constants, and ``W / sigma``.
    """
    w = state[prefix + ".parametrizations.weight.original"]
    u = state[prefix + ".parametrizations.weight.0._u"]
    v = state[prefix + ".parametrizations.weight.0._v"]
    wm = w.flatten(1)
    if training:
        with torch.no_grad():
            wd = wm.detach()
            u.copy_(_normalize(torch.mv(wd, v), eps))
            v.copy_(_normalize(torch.mv(wd.t(), u), eps))
    uc, vc = u.clone(), v.clone()
    sigma = torch.dot(uc, torch.mv(wm, vc))
    return w / sigma


def sn_conv(state: State, prefix: str, x: torch.Tensor, training: bool, eps: float,
            padding: int = 0) -> torch.Tensor:
    w = sn_weight(state, prefix, training, eps)
    b = state[prefix + ".bias"]
    if w.dim() == 5:
        return F.conv3d(x, w, b, padding=padding)
    return F.conv2d(x, w, b, padding=padding)


def plain_conv(state: State, prefix: str, x: torch.Tensor, padding: int = 0) -> torch.Tensor:
    return F.conv2d(x, state[prefix + ".weight"], state.get(prefix + ".bias"), padding=padding)


def batch_norm(state: State, prefix: str, x: torch.Tensor, training: bool) -> torch.Tensor:
    """BatchNorm2d/1d as the reference uses it (dgmr/common.py:38-39; generators.py:113;
    discriminators.py:102,194): batch statistics + running-stat update (momentum 0.1,
    unbiased running var) in training, running statistics in eval."""
    w, b = state[prefix + ".weight"], state[prefix + ".bias"]
    rm, rv = state[prefix + ".running_mean"], state[prefix + ".running_var"]
    dims = [0] + list(range(2, x.dim()))
    shape = [1, -1] + [1] * (x.dim() - 2)
    if training:
        n = x.numel() // x.shape[1]
        if n <= 1:
            raise ValueError("Expected more than 1 value per channel when training")
        mean = x.mean(dims)
        var = x.var(dims, unbiased=False)
        with torch.no_grad():
            rm.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * mean.detach())
            rv.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * var.detach() * (n / (n - 1)))
            state[prefix + ".num_batches_tracked"] += 1
    else:
        mean, var = rm, rv
    xhat = (x - mean.view(shape)) * torch.rsqrt(var.view(shape) + BN_EPS)
    return xhat * w.view(shape) + b.view(shape)


def pixel_unshuffle(x: torch.Tensor, r: int = 2) -> torch.Tensor:
    """out[..., c*r*r + i*r + j, h, w] = in[..., c, h*r+i, w*r+j]  (torch PixelUnshuffle,
    used at dgmr/common.py:326,393 and discriminators.py:69,166).  Pure index permutation."""
    *lead, c, h, w = x.shape
    x = x.reshape(*lead, c, h // r, r, w // r, r)
    nd = len(lead)
    perm = list(range(nd)) + [nd, nd + 2, nd + 4, nd + 1, nd + 3]
    return x.permute(perm).reshape(*lead, c * r * r, h // r, w // r)


def pixel_shuffle(x: torch.Tensor, r: int = 2) -> torch.Tensor:
    """out[..., c, h*r+i, w*r+j] = in[..., c*r*r + i*r + j, h, w] (generators.py:123,178)."""
    *lead, c, h, w = x.shape
    co = c // (r * r)
    x = x.reshape(*lead, co, r, r, h, w)
    nd = len(lead)
    perm = list(range(nd)) + [nd, nd + 3, nd + 1, nd + 4, nd + 2]
    return x.permute(perm).reshape(*lead, co, h * r, w * r)


def upsample2(x: torch.Tensor) -> torch.Tensor:
    # torch.nn.Upsample(scale_factor=2, mode="nearest") (dgmr/common.py:121)
    return x.repeat_interleave(2, dim=-2).repeat_interleave(2, dim=-1)


# --------------------------------------------------------------------------- blocks
def conv_gru_cell(state: State, p: str, x: torch.Tensor, h: torch.Tensor, training: bool):
    """dgmr/layers/ConvGRU.py:57-85."""
    xh = torch.cat([x, h], dim=1)
    r = torch.sigmoid(sn_conv(state, p + ".read_gate_conv", xh, training, SN_EPS_G, 1))
    u = torch.sigmoid(sn_conv(state, p + ".update_gate_conv", xh, training, SN_EPS_G, 1))
    gated = torch.cat([x, r * h], dim=1)
    c = torch.relu(sn_conv(state, p + ".output_conv", gated, training, SN_EPS_G, 1))
    out = u * h + (1.0 - u) * c
    return out, out


def conv_gru(state: State, p: str, xs: Sequence[torch.Tensor], h: torch.Tensor, training: bool):
    """dgmr/layers/ConvGRU.py:102-111 (time loop, outputs stacked on dim 0)."""
    outs = []
    for x in xs:
        o, h = conv_gru_cell(state, p + ".cell", x, h, training)
        outs.append(o)
    return torch.stack(outs, dim=0)


def g_block(state: State, p: str, x: torch.Tensor, training: bool) -> torch.Tensor:
    """dgmr/common.py:68-84."""
    cout = state[p + ".last_conv_3x3.bias"].shape[0]
    sc = x if x.shape[1] == cout else sn_conv(state, p + ".conv_1x1", x, training, SN_EPS_G)
    y = torch.relu(batch_norm(state, p + ".bn1", x, training))
    y = sn_conv(state, p + ".first_conv_3x3", y, training, SN_EPS_G, 1)
    y = torch.relu(batch_norm(state, p + ".bn2", y, training))
    y = sn_conv(state, p + ".last_conv_3x3", y, training, SN_EPS_G, 1)
    return y + sc


def upsample_g_block(state: State, p: str, x: torch.Tensor, training: bool) -> torch.Tensor:
    """dgmr/common.py:139-155."""
    sc = sn_conv(state, p + ".conv_1x1", upsample2(x), training, SN_EPS_G)
    y = torch.relu(batch_norm(state, p + ".bn1", x, training))
    y = upsample2(y)
    y = sn_conv(state, p + ".first_conv_3x3", y, training, SN_EPS_G, 1)
    y = torch.relu(batch_norm(state, p + ".bn2", y, training))
    y = sn_conv(state, p + ".last_conv_3x3", y, training, SN_EPS_G, 1)
    return y + sc


def d_block(state: State, p: str, x: torch.Tensor, training: bool, first_relu: bool = True,
            keep_same_output: bool = False) -> torch.Tensor:
    """dgmr/common.py:220-238 (2-D and 3-D variants; the conv rank follows the weight rank)."""
    w1 = state[p + ".first_conv_3x3.parametrizations.weight.original"]
    cin, cout, is3d = w1.shape[1], w1.shape[0], w1.dim() == 5
    pool = (lambda t: F.avg_pool3d(t, 2, 2)) if is3d else (lambda t: F.avg_pool2d(t, 2, 2))
    if cin != cout:
        x1 = sn_conv(state, p + ".conv_1x1", x, training, SN_EPS_DEFAULT)
        if not keep_same_output:
            x1 = pool(x1)
    else:
        x1 = x
    y = torch.relu(x) if first_relu else x
    y = sn_conv(state, p + ".first_conv_3x3", y, training, SN_EPS_DEFAULT, 1)
    y = torch.relu(y)
    y = sn_conv(state, p + ".last_conv_3x3", y, training, SN_EPS_DEFAULT, 1)
    if not keep_same_output:
        y = pool(y)
    return x1 + y


def l_block(state: State, p: str, x: torch.Tensor) -> torch.Tensor:
    """dgmr/common.py:288-300 (plain, un-normalised convolutions)."""
    cin = x.shape[1]
    cout = state[p + ".last_conv_3x3.weight"].shape[0]
    if cin < cout:
        sc = torch.cat([x, plain_conv(state, p + ".conv_1x1", x)], dim=1)
    else:
        sc = x
    y = plain_conv(state, p + ".first_conv_3x3", torch.relu(x), 1)
    y = plain_conv(state, p + ".last_conv_3x3", torch.relu(y), 1)
    return y + sc


def attention(state: State, p: str, x: torch.Tensor) -> torch.Tensor:
    """dgmr/layers/Attention.py:71-85 with attention_einsum :9-20 (unscaled logits)."""
    q = plain_conv(state, p + ".query", x)
    k = plain_conv(state, p + ".key", x)
    v = plain_conv(state, p + ".value", x)
    # Reference quirk (parity depends on it): forward() hands `query[b]`, a [C, H, W] tensor, to
    # attention_einsum(), whose einsum strings label the axes "h w c".  So the "positions" are
    # the C*H (channel, row) pairs and the contracted feature axis is the image column W:
    #   beta = softmax_L( Q K^T ),  Q, K, V = q/k/v[b].reshape(C*H, W);  out = (beta V) as [C,H,W]
    b, c, h, w = q.shape
    qf = q.reshape(b, c * h, w)
    kf = k.reshape(b, c * h, w)
    vf = v.reshape(b, v.shape[1] * h, w)
    beta = torch.softmax(torch.bmm(qf, kf.transpose(1, 2)), dim=-1)
    o = torch.bmm(beta, vf).reshape(b, v.shape[1], h, w)
    o = state[p + ".gamma"] * plain_conv(state, p + ".last_conv", o)
    return o + x


# --------------------------------------------------------------------------- stacks
def context_stack(state: State, p: str, x: torch.Tensor, training: bool):
    """dgmr/common.py:388-424.  x: [B, T_in, C, H, W] -> 4 scales (largest first)."""
    x = pixel_unshuffle(x, 2)
    steps = x.shape[1]
    scales: List[List[torch.Tensor]] = [[], [], [], []]
    for i in range(steps):
        s = x[:, i]
        for lvl, name in enumerate(("d1", "d2", "d3", "d4")):
            s = d_block(state, f"{p}.{name}" if p else name, s, training)
            scales[lvl].append(s)
    outs = []
    for lvl, name in enumerate(("conv1", "conv2", "conv3", "conv4")):
        st = torch.stack(scales[lvl], dim=1)  # b t c h w
        b, t, c, h, w = st.shape
        mixed = st.permute(0, 2, 1, 3, 4).reshape(b, c * t, h, w)  # "(c t)": channel = c*T + t
        outs.append(torch.relu(sn_conv(state, f"{p}.{name}" if p else name, mixed, training,
                                       SN_EPS_DEFAULT, 1)))
    return tuple(outs)


def latent_stack(state: State, p: str, shape: Tuple[int, int, int], training: bool,
                 z: Optional[torch.Tensor] = None, dtype=torch.float32, device=None) -> torch.Tensor:
    """dgmr/common.py:469-497.  The reference draws z = Normal(0,1).sample(shape) on the
    CPU default generator (:481), which is `torch.randn(shape + (1,))` [probe, SURVEY 8a]."""
    q = (lambda n: f"{p}.{n}" if p else n)
    if z is None:
        z = torch.normal(torch.zeros(tuple(shape) + (1,)), torch.ones(tuple(shape) + (1,)))
    z = z.permute(3, 0, 1, 2).to(dtype=dtype, device=device)   # `.type_as(x)` in the reference (:483)
    z = sn_conv(state, q("conv_3x3"), z, training, SN_EPS_DEFAULT, 1)
    z = l_block(state, q("l_block1"), z)
    z = l_block(state, q("l_block2"), z)
    z = l_block(state, q("l_block3"), z)
    if q("att_block.gamma") in state:
        z = attention(state, q("att_block"), z)
    z = l_block(state, q("l_block4"), z)
    return z


def sampler(state: State, p: str, cond: Sequence[torch.Tensor], latent: torch.Tensor,
            forecast_steps: int, training: bool) -> torch.Tensor:
    """dgmr/generators.py:125-182."""
    q = (lambda n: f"{p}.{n}" if p else n)
    b = cond[0].shape[0]
    latent = latent.repeat(b, 1, 1, 1)  # "(repeat b) c h w", latent batch is 1
    hs: List[torch.Tensor] = [latent] * forecast_steps
    names = [("convGRU1", "gru_conv_1x1", "g1", "up_g1"),
             ("convGRU2", "gru_conv_1x1_2", "g2", "up_g2"),
             ("convGRU3", "gru_conv_1x1_3", "g3", "up_g3"),
             ("convGRU4", "gru_conv_1x1_4", "g4", "up_g4")]
    for lvl, (gru, c11, g, ug) in enumerate(names):
        hs = list(conv_gru(state, q(gru), hs, cond[3 - lvl], training))
        hs = [sn_conv(state, q(c11), h, training, SN_EPS_DEFAULT) for h in hs]
        hs = [g_block(state, q(g), h, training) for h in hs]
        hs = [upsample_g_block(state, q(ug), h, training) for h in hs]
    hs = [torch.relu(batch_norm(state, q("bn"), h, training)) for h in hs]
    hs = [sn_conv(state, q("conv_1x1"), h, training, SN_EPS_DEFAULT) for h in hs]
    hs = [pixel_shuffle(h, 2) for h in hs]
    return torch.stack(hs, dim=1)


def generator(state: State, x: torch.Tensor, forecast_steps: int, latent_shape, training: bool,
              z: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dgmr/generators.py:207-212 (context stack, then latent stack, then sampler)."""
    cond = context_stack(state, "conditioning_stack", x, training)
    lat = latent_stack(state, "latent_stack", latent_shape, training, z=z, dtype=x.dtype, device=x.device)
    return sampler(state, "sampler", cond, lat, forecast_steps, training)


def _d_head(state: State, p: str, rep: torch.Tensor, training: bool) -> torch.Tensor:
    rep = torch.relu(rep).sum(dim=[2, 3])
    rep = batch_norm(state, p + ".bn", rep, training)
    w = sn_weight(state, p + ".fc", training, SN_EPS_DEFAULT)
    return F.linear(rep, w, state[p + ".fc.bias"])


def spatial_discriminator(state: State, p: str, x: torch.Tensor, training: bool,
                          num_timesteps: int = 8, idxs: Optional[torch.Tensor] = None):
    """dgmr/discriminators.py:196-232."""
    if idxs is None:
        idxs = torch.randint(low=0, high=x.shape[1], size=(num_timesteps,))
    n_inter = len({k.split(".")[len(p.split(".")) + 1] for k in state
                   if k.startswith(p + ".intermediate_dblocks.")})
    reps = []
    for idx in idxs.tolist():
        rep = F.avg_pool2d(x[:, idx], 2)
        rep = pixel_unshuffle(rep, 2)
        rep = d_block(state, p + ".d1", rep, training, first_relu=False)
        for i in range(n_inter):
            rep = d_block(state, f"{p}.intermediate_dblocks.{i}", rep, training)
        rep = d_block(state, p + ".d6", rep, training, keep_same_output=True)
        reps.append(_d_head(state, p, rep, training))
    return torch.stack(reps, dim=1).sum(dim=1, keepdim=True)


def temporal_discriminator(state: State, p: str, x: torch.Tensor, training: bool):
    """dgmr/discriminators.py:104-138."""
    x = F.avg_pool3d(x, (1, 2, 2), (1, 2, 2))  # T acts as the channel dim (:106)
    x = pixel_unshuffle(x, 2)
    x = x.permute(0, 2, 1, 3, 4)
    x = d_block(state, p + ".d1", x, training, first_relu=False)
    x = d_block(state, p + ".d2", x, training)
    x = x.permute(0, 2, 1, 3, 4)
    n_inter = len({k.split(".")[len(p.split(".")) + 1] for k in state
                   if k.startswith(p + ".intermediate_dblocks.")})
    reps = []
    for t in range(x.shape[1]):
        rep = x[:, t]
        for i in range(n_inter):
            rep = d_block(state, f"{p}.intermediate_dblocks.{i}", rep, training)
        rep = d_block(state, p + ".d_last", rep, training, keep_same_output=True)
        reps.append(_d_head(state, p, rep, training))
    return torch.stack(reps, dim=1).sum(dim=1, keepdim=True)


def discriminator(state: State, x: torch.Tensor, training: bool, num_spatial_frames: int = 8,
                  idxs: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dgmr/discriminators.py:39-44: spatial first (RNG order), then temporal; cat on dim 1."""
    s = spatial_discriminator(state, "spatial_discriminator", x, training, num_spatial_frames, idxs)
    t = temporal_discriminator(state, "temporal_discriminator", x, training)
    return torch.cat([s, t], dim=1)


# --------------------------------------------------------------------------- losses
def loss_hinge_disc(score_generated: torch.Tensor, score_real: torch.Tensor) -> torch.Tensor:
    """dgmr/losses.py:307-313."""
    return torch.relu(1.0 - score_real).mean() + torch.relu(1.0 + score_generated).mean()


def loss_hinge_gen(score_generated: torch.Tensor) -> torch.Tensor:
    """dgmr/losses.py:316-319."""
    return -score_generated.mean()


def grid_cell_loss(generated: torch.Tensor, targets: torch.Tensor, cap: float = 24.0):
    """dgmr/losses.py:172-192 with weight_fn dgmr/dgmr.py:20-33: w(y) = max(y + 1, cap);
    `norm_1 / T * H * W` precedence kept verbatim."""
    w = torch.clamp_min(targets + 1, cap)
    diff = ((generated - targets) * w).abs().sum()
    return diff / targets.size(1) * targets.size(3) * targets.size(4)


# --------------------------------------------------------------------------- state helpers
def split_state(full: State, prefix: str) -> State:
    """Sub-dict of a DGMR state dict with ``prefix`` stripped (tensors shared, not copied)."""
    n = len(prefix)
    return {k[n:]: v for k, v in full.items() if k.startswith(prefix)}


def clone_state(state: State, requires_grad: bool = False) -> State:
    out = {}
    for k, v in state.items():
        t = v.detach().clone()
        if requires_grad and t.is_floating_point() and not (
                k.endswith("._u") or k.endswith("._v") or "running_" in k):
            t.requires_grad_(True)
        out[k] = t
    return out


# --------------------------------------------------------------------------- training step
class AdamState:
    """torch.optim.Adam(lr, betas) restated for a flat list of tensors
    (dgmr/dgmr.py:292-300; eps 1e-8, no weight decay, no amsgrad)."""

    def __init__(self, params: List[torch.Tensor], lr: float, betas=(0.0, 0.999), eps: float = 1e-8):
        self.params, self.lr, self.betas, self.eps = params, lr, betas, eps
        self.m = [torch.zeros_like(p) for p in params]
        self.v = [torch.zeros_like(p) for p in params]
        self.t = 0

    @torch.no_grad()
    def step(self, grads: List[Optional[torch.Tensor]]):
        self.t += 1
        b1, b2 = self.betas
        bc1, bc2 = 1 - b1 ** self.t, 1 - b2 ** self.t
        for p, g, m, v in zip(self.params, grads, self.m, self.v):
            if g is None:
                continue
            m.mul_(b1).add_(g, alpha=1 - b1)
            v.mul_(b2).addcmul_(g, g, value=1 - b2)
            denom = (v.sqrt() / math.sqrt(bc2)).add_(self.eps)
            p.addcdiv_(m, denom, value=-self.lr / bc1)


def _trainable(state: State) -> List[str]:
    return [k for k, v in state.items() if v.is_floating_point() and v.requires_grad]


def gan_step(g_state: State, d_state: State, g_opt: AdamState, d_opt: AdamState,
             images: torch.Tensor, future: torch.Tensor, forecast_steps: int, latent_shape,
             generation_steps: int = 1, grid_lambda: float = 20.0, cap: float = 24.0):
    """One GAN step in the parity-preserving minimal schedule of SURVEY.md 8(d):
    the losses, D/G gradients and parameter updates of dgmr/dgmr.py:137-218, without the
    wrapper's wasted work (G is not back-propagated in the D phase, no checkpoint recompute,
    no trailing forward).  ``g_state``/``d_state`` tensors with requires_grad are the
    parameters; the optimisers were built over ``_trainable`` lists in the same order.
    Returns dict of losses."""
    g_names, d_names = _trainable(g_state), _trainable(d_state)
    real_seq = torch.cat([images, future], dim=1)
    d_loss = None
    for _ in range(2):
        with torch.no_grad():
            pred = generator(g_state, images, forecast_steps, latent_shape, True)
        gen_seq = torch.cat([images, pred], dim=1)
        out = discriminator(d_state, torch.cat([real_seq, gen_seq], dim=0), True)
        b = real_seq.shape[0]
        s_real, s_gen = out[:b], out[b:]
        d_loss = loss_hinge_disc(s_gen[:, 0:1], s_real[:, 0:1]) + loss_hinge_disc(s_gen[:, 1:2], s_real[:, 1:2])
        grads = torch.autograd.grad(d_loss, [d_state[k] for k in d_names], allow_unused=True)
        d_opt.step(list(grads))
    preds = [generator(g_state, images, forecast_steps, latent_shape, True) for _ in range(generation_steps)]
    gen_mean = torch.stack(preds, dim=0).mean(dim=0)
    grid = grid_cell_loss(gen_mean, future, cap)
    scores = []
    for pr in preds:
        out = discriminator(d_state, torch.cat([real_seq, torch.cat([images, pr], dim=1)], dim=0), True)
        scores.append(out[real_seq.shape[0]:])
    g_loss = loss_hinge_gen(torch.cat(scores, dim=0)) + grid_lambda * grid
    grads = torch.autograd.grad(g_loss, [g_state[k] for k in g_names], allow_unused=True)
    g_opt.step(list(grads))
    return {"d_loss": d_loss.detach(), "g_loss": g_loss.detach(), "grid_loss": grid.detach()}
