"""Shared helpers for parity tests: build the B200 modules from a seed, run the CPU oracle on the same
state dict, compare outputs / buffers / gradients."""
import os

import torch

from oracle import dgmr_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c1_gan.pt")
C1 = dict(forecast_steps=4, output_shape=128, latent_channels=384, context_channels=192, batch=2)


def state_checksum(sd):
    tot, atot = 0.0, 0.0
    for k in sorted(sd):
        v = sd[k].double()
        tot += float(v.sum())
        atot += float(v.abs().sum())
    return [tot, atot]


def build_gan(cfg, seed=0, gamma=None):
    """Seeded construction on CPU (consumes the RNG exactly like the reference constructors)."""
    import skillful_nowcasting_b200 as B

    torch.manual_seed(seed)
    s = cfg["output_shape"]
    gen = B.Generator(
        B.ContextConditioningStack(input_channels=1, output_channels=cfg["context_channels"]),
        B.LatentConditioningStack(shape=(8, s // 32, s // 32), output_channels=cfg["latent_channels"]),
        B.Sampler(forecast_steps=cfg["forecast_steps"], latent_channels=cfg["latent_channels"],
                  context_channels=cfg["context_channels"]))
    disc = B.Discriminator(input_channels=1)
    if gamma is not None:
        with torch.no_grad():
            gen.latent_stack.att_block.gamma.fill_(gamma)
    return gen, disc


def c1_inputs(cfg=C1, seed=1):
    torch.manual_seed(seed)
    s, b, t = cfg["output_shape"], cfg["batch"], cfg["forecast_steps"]
    return torch.rand(b, 4, 1, s, s), torch.rand(b, t, 1, s, s)


def rel_err(a, b):
    """max |a-b| / max |b| (the 'rel' used for all tolerances in these tests)."""
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


def oracle_gan_forward(g_state, d_state, x, y, cfg, training, seed):
    """Oracle forward (+ losses, + grads in training) on cloned state.  Returns dict like the golden fixture."""
    gs = O.clone_state(g_state, requires_grad=training)
    ds = O.clone_state(d_state, requires_grad=training)
    s = cfg["output_shape"]
    torch.manual_seed(seed)
    out = O.generator(gs, x, cfg["forecast_steps"], (8, s // 32, s // 32), training)
    scores = O.discriminator(ds, torch.cat([torch.cat([x, y], 1), torch.cat([x, out], 1)], 0), training)
    b = x.shape[0]
    sr, sg = scores[:b], scores[b:]
    d_loss = O.loss_hinge_disc(sg[:, 0:1], sr[:, 0:1]) + O.loss_hinge_disc(sg[:, 1:2], sr[:, 1:2])
    grid = O.grid_cell_loss(out, y)
    g_loss = O.loss_hinge_gen(sg) + 20.0 * grid
    res = dict(out=out.detach(), scores=scores.detach(), d_loss=d_loss.detach(), grid=grid.detach(), g_loss=g_loss.detach(),
               g_state=gs, d_state=ds)
    if training:
        dn, gn = O._trainable(ds), O._trainable(gs)
        dg = torch.autograd.grad(d_loss, [ds[k] for k in dn], retain_graph=True, allow_unused=True)
        gg = torch.autograd.grad(g_loss, [gs[k] for k in gn], allow_unused=True)
        res["d_grads"] = {k: g for k, g in zip(dn, dg) if g is not None}
        res["g_grads"] = {k: g for k, g in zip(gn, gg) if g is not None}
    return res


def module_gan_forward(gen, disc, x, y, training, seed, device):
    """Same computation through the B200 modules (whatever backend is active)."""
    from skillful_nowcasting_b200 import losses, ops

    gen.train(training)
    disc.train(training)
    for p in list(gen.parameters()) + list(disc.parameters()):
        p.grad = None
    x, y = x.to(device), y.to(device)
    torch.manual_seed(seed)
    out = gen(x)
    scores = disc(torch.cat([torch.cat([x, y], 1), torch.cat([x, out], 1)], 0))
    b = x.shape[0]
    d_loss = losses.loss_hinge_disc_both(scores)
    grid = ops.grid_cell(out, y, 24.0)
    g_loss = losses.loss_hinge_gen(scores[b:]) + 20.0 * grid
    res = dict(out=out.detach(), scores=scores.detach(), d_loss=d_loss.detach(), grid=grid.detach(), g_loss=g_loss.detach())
    if training:
        dn, gn = dict(disc.named_parameters()), dict(gen.named_parameters())
        dg = torch.autograd.grad(d_loss, list(dn.values()), retain_graph=True, allow_unused=True)
        gg = torch.autograd.grad(g_loss, list(gn.values()), allow_unused=True)
        res["d_grads"] = {k: g for k, g in zip(dn, dg) if g is not None}
        res["g_grads"] = {k: g for k, g in zip(gn, gg) if g is not None}
    return res


def compare_grads(got, ref, tol_norm, tol_head, zero_floor):
    """ref entries are either tensors (oracle) or fixture summaries {norm, sum, head[, full]}.
    Gradients whose reference norm is below `zero_floor` x the largest norm are mathematically zero
    (conv bias in front of BatchNorm) and only checked to be comparably small."""
    assert set(got) == set(ref), (set(got) ^ set(ref))
    norms = {k: (float(v.double().norm()) if torch.is_tensor(v) else v["norm"]) for k, v in ref.items()}
    big = max(norms.values())
    worst = []
    for k, g in got.items():
        g = g.detach().float().cpu()
        rn = norms[k]
        gn = float(g.double().norm())
        if rn < zero_floor * big:
            assert gn < 10 * zero_floor * big, f"{k}: expected ~0 gradient, got norm {gn:.3e}"
            continue
        e_norm = abs(gn - rn) / rn
        if torch.is_tensor(ref[k]):
            r = ref[k].detach().float()
            e_head = (g - r).abs().max().item() / max(r.abs().max().item(), 1e-30)
        else:
            r = ref[k]["head"]
            e_head = (g.flatten()[:64] - r).abs().max().item() / max(r.abs().max().item(), 1e-30)
        worst.append((max(e_norm / tol_norm, e_head / tol_head), k, e_norm, e_head))
    worst.sort(reverse=True)
    assert not worst or worst[0][0] <= 1.0, f"gradient mismatch (ratio, name, norm err, elem err): {worst[:5]}"
    return worst[:3]


def assert_grads_close(names, got, ref, tol, floor=1e-5, tol_l2=None):
    """Element-wise gradient comparison for block tests.  Gradients that are mathematically zero (a conv bias in front
    of a training-mode BatchNorm) show up as rounding noise on both sides: they are only required to stay tiny."""
    pairs = [(n, a, b) for n, a, b in zip(names, got, ref)]
    for n, a, b in pairs:
        assert (a is None) == (b is None), n
    scale = max([float(b.abs().max()) for _, _, b in pairs if b is not None and b.numel()] + [1e-30])
    for n, a, b in pairs:
        if a is None or not b.numel():
            continue
        if float(b.abs().max()) < floor * scale:
            assert float(a.abs().max()) < 10 * floor * scale, f"{n}: expected ~0 gradient"
        else:
            e = rel_err(a, b)
            assert e < tol, f"{n}: rel err {e:.3e} > {tol}"
            if tol_l2 is not None:  # ||a-b||_2 / ||b||_2: insensitive to the handful of ReLU-mask flips tf32 rounding causes
                e2 = float((a.double().cpu() - b.double()).norm() / b.double().norm())
                assert e2 < tol_l2, f"{n}: L2 rel err {e2:.3e} > {tol_l2}"


def global_grad_error(got, ref):
    """||g - r|| / ||r|| over ALL parameters at once (oracle tensors only): the coarse end-to-end check used where
    per-parameter comparison is dominated by chaotic amplification (train-mode TF32 end to end)."""
    num = den = 0.0
    for k, r in ref.items():
        g = got[k].detach().double().cpu()
        r = r.detach().double()
        num += float(((g - r) ** 2).sum())
        den += float((r ** 2).sum())
    return (num / max(den, 1e-300)) ** 0.5
