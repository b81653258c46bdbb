"""Generate golden fixtures by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Writes tests/golden/c1_gan.pt.  The reference has no golden vectors of its own for the hot
path, so these fixtures — outputs of the reference itself on seeded inputs —
are what pins the oracle (tests/test_oracle.py) and, through it, the CUDA path.

Import bypass: `import dgmr` fails because dgmr/__init__.py pulls in
pytorch_lightning; registering a bare package object lets the hot-path sub-modules import.
"""
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("DGMR_REFERENCE", "/root/reference")


def import_reference():
    if "dgmr" in sys.modules and getattr(sys.modules["dgmr"], "__graft_ref__", False):
        return sys.modules["dgmr"]
    for k in [k for k in sys.modules if k == "dgmr" or k.startswith("dgmr.")]:
        del sys.modules[k]
    m = types.ModuleType("dgmr")
    m.__path__ = [os.path.join(REF, "dgmr")]
    m.__graft_ref__ = True
    sys.modules["dgmr"] = m
    if "pytorch_msssim" not in sys.modules:
        try:
            import pytorch_msssim  # noqa: F401
        except Exception:
            stub = types.ModuleType("pytorch_msssim")
            stub.SSIM = type("SSIM", (torch.nn.Module,), {})
            stub.MS_SSIM = type("MS_SSIM", (torch.nn.Module,), {})
            sys.modules["pytorch_msssim"] = stub
    import dgmr.common  # noqa: F401
    import dgmr.discriminators  # noqa: F401
    import dgmr.generators  # noqa: F401
    import dgmr.layers  # noqa: F401
    import dgmr.losses  # noqa: F401
    return m


# BASELINE.json configs[0]: the reference's own CPU smoke config (tests/test_model.py:285-306).
C1 = dict(forecast_steps=4, output_shape=128, latent_channels=384, context_channels=192, batch=2)


def state_checksum(sd):
    """Order-independent fingerprint of a state dict (fp64 sums), used on the GPU box to check
    that seeded construction reproduced the weights this fixture was generated with."""
    tot, atot = 0.0, 0.0
    for k in sorted(sd):
        v = sd[k].double()
        tot += float(v.sum())
        atot += float(v.abs().sum())
    return [tot, atot]


def summarize_grads(named):
    """Full tensors are too big to commit (G 13 M / D 45 M params): keep exact small ones,
    and (sum, L2 norm, first 64 values) of the rest."""
    out = {}
    for k, g in named.items():
        g = g.detach()
        ent = {"sum": float(g.double().sum()), "norm": float(g.double().norm()),
               "head": g.flatten()[:64].clone()}
        if g.numel() <= 4096:
            ent["full"] = g.clone()
        out[k] = ent
    return out


def build_reference_gan(cfg, seed=0):
    import_reference()
    from dgmr.common import ContextConditioningStack, LatentConditioningStack
    from dgmr.discriminators import Discriminator
    from dgmr.generators import Generator, Sampler

    torch.manual_seed(seed)
    s = cfg["output_shape"]
    gen = Generator(
        ContextConditioningStack(input_channels=1, output_channels=cfg["context_channels"]),
        LatentConditioningStack(shape=(8, s // 32, s // 32), output_channels=cfg["latent_channels"]),
        Sampler(forecast_steps=cfg["forecast_steps"], latent_channels=cfg["latent_channels"],
                context_channels=cfg["context_channels"]),
    )
    disc = Discriminator(input_channels=1)
    return gen, disc


def weight_fn(y, cap=24.0):  # dgmr/dgmr.py:20-33 (cannot import dgmr.dgmr without lightning)
    return torch.max(y + 1, torch.tensor(cap, device=y.device))


def run_case(gen, disc, g0, d0, x, y, training, seed):
    from dgmr.losses import GridCellLoss, loss_hinge_disc, loss_hinge_gen

    gen.load_state_dict(g0)
    disc.load_state_dict(d0)
    gen.train(training)
    disc.train(training)
    # attention gamma is 0 at init, which hides the attention path: give it a value
    with torch.no_grad():
        gen.latent_stack.att_block.gamma.fill_(0.5)
    for p in list(gen.parameters()) + list(disc.parameters()):
        p.grad = None
    torch.manual_seed(seed)
    out = gen(x)
    real = torch.cat([x, y], dim=1)
    fake = torch.cat([x, out], dim=1)
    scores = disc(torch.cat([real, fake], dim=0))
    b = x.shape[0]
    s_real, s_gen = scores[:b], scores[b:]
    d_loss = loss_hinge_disc(s_gen[:, 0:1], s_real[:, 0:1]) + loss_hinge_disc(s_gen[:, 1:2], s_real[:, 1:2])
    grid = GridCellLoss(weight_fn=weight_fn)(out, y)
    g_loss = loss_hinge_gen(s_gen) + 20.0 * grid
    res = {
        "out": out.detach().clone(),
        "scores": scores.detach().clone(),
        "d_loss": d_loss.detach().clone(),
        "grid": grid.detach().clone(),
        "g_loss": g_loss.detach().clone(),
    }
    if training:
        d_params = dict(disc.named_parameters())
        g_params = dict(gen.named_parameters())
        dg = torch.autograd.grad(d_loss, list(d_params.values()), retain_graph=True, allow_unused=True)
        gg = torch.autograd.grad(g_loss, list(g_params.values()), allow_unused=True)
        res["d_grads"] = summarize_grads({k: g for k, g in zip(d_params, dg) if g is not None})
        res["g_grads"] = summarize_grads({k: g for k, g in zip(g_params, gg) if g is not None})
        res["g_state_after"] = {k: v.clone() for k, v in gen.state_dict().items()
                                if k.endswith("._u") or k.endswith("._v") or "running_" in k
                                or "num_batches" in k}
        res["d_state_after"] = {k: v.clone() for k, v in disc.state_dict().items()
                                if k.endswith("._u") or k.endswith("._v") or "running_" in k
                                or "num_batches" in k}
    return res


def main():
    cfg = C1
    gen, disc = build_reference_gan(cfg, seed=0)
    with torch.no_grad():
        gen.latent_stack.att_block.gamma.fill_(0.5)
    g0 = {k: v.clone() for k, v in gen.state_dict().items()}
    d0 = {k: v.clone() for k, v in disc.state_dict().items()}
    torch.manual_seed(1)
    s, b, t = cfg["output_shape"], cfg["batch"], cfg["forecast_steps"]
    x = torch.rand(b, 4, 1, s, s)
    y = torch.rand(b, t, 1, s, s)
    fix = {"cfg": cfg, "g_checksum": state_checksum(g0), "d_checksum": state_checksum(d0),
           "init_seed": 0, "data_seed": 1, "seed": 2, "gamma": 0.5, "torch_version": str(torch.__version__)}
    fix["eval"] = run_case(gen, disc, g0, d0, x, y, False, 2)
    fix["train"] = run_case(gen, disc, g0, d0, x, y, True, 2)
    # RNG draws the reference makes inside that forward, recorded for documentation
    torch.manual_seed(2)
    fix["z"] = torch.normal(torch.zeros(8, s // 32, s // 32, 1), torch.ones(8, s // 32, s // 32, 1))
    fix["idxs"] = torch.randint(low=0, high=4 + t, size=(8,))
    path = os.path.join(HERE, "c1_gan.pt")
    torch.save(fix, path)
    print("wrote", path, os.path.getsize(path) / 1e6, "MB")
    print("eval out sum", float(fix["eval"]["out"].double().sum()), "scores", fix["eval"]["scores"].flatten().tolist())
    print("train out sum", float(fix["train"]["out"].double().sum()), "d_loss", float(fix["train"]["d_loss"]))


if __name__ == "__main__":
    main()
