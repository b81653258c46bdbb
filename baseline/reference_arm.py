"""The reference arm of bench.py: runs the UNMODIFIED reference (openclimatefix/skillful_nowcasting, `dgmr` 1.4.4) through its own
public API -- `dgmr.DGMR(...).training_step(batch, 0)` -- on the host cores or on the GPU (PyTorch eager + cuDNN).

Where the reference comes from: `baseline/_ref/` (a `pip install --no-deps --target baseline/_ref` of /root/reference done by
`__graft_entry__.build()` in the build container; git-ignored, travels to the GPU box with the snapshot) or, in the build
container, /root/reference itself.  Nothing of this repository's kernels, modules or engine is on that path.

Two dependencies of the reference are absent from the image and the wheelhouse: `pytorch_lightning` (the
LightningModule base class of `dgmr.dgmr.DGMR`) and `pytorch_msssim` (only used by loss classes that are off the path).  They are
replaced by the minimal stand-ins below: four Lightning methods that `training_step` calls (manual optimisation), no arithmetic.

`dropin=True` additionally swaps this repository's modules in under the reference's *unmodified* `dgmr/dgmr.py` wrapper
(`dgmr.common`, `dgmr.generators`, `dgmr.discriminators` -> skillful_nowcasting_b200.*): SURVEY.md 8d's mode (i).
"""
from __future__ import annotations

import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_root():
    """Directory that contains the reference's `dgmr/` package, or None."""
    for cand in (os.environ.get("DGMR_REFERENCE"), os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if cand and os.path.isfile(os.path.join(cand, "dgmr", "dgmr.py")):
            return cand
    return None


class _LightningModuleStub(torch.nn.Module):
    """The slice of pytorch_lightning.LightningModule that dgmr/dgmr.py uses (manual optimisation; ref: dgmr/dgmr.py:103,129,144,168)."""

    def __init__(self, *a, **kw):
        super().__init__()
        self._opts = None
        self.logged = {}
        self.global_step = 0

    def save_hyperparameters(self, *a, **kw):
        pass

    def optimizers(self):
        if self._opts is None:
            self._opts = self.configure_optimizers()[0]
        return self._opts

    def manual_backward(self, loss, *a, **kw):
        loss.backward(*a, **kw)

    def log_dict(self, d, **kw):
        self.logged.update({k: v.detach() for k, v in d.items()})

    def log(self, k, v, **kw):
        self.logged[k] = v


def _install_stubs():
    if "pytorch_lightning" not in sys.modules:
        try:
            import pytorch_lightning  # noqa: F401
        except Exception:  # noqa: BLE001
            pl = types.ModuleType("pytorch_lightning")
            pl.LightningModule = _LightningModuleStub
            pl.__graft_stub__ = True
            sys.modules["pytorch_lightning"] = pl
    if "pytorch_msssim" not in sys.modules:
        try:
            import pytorch_msssim  # noqa: F401
        except Exception:  # noqa: BLE001
            ms = types.ModuleType("pytorch_msssim")
            ms.SSIM = type("SSIM", (torch.nn.Module,), {})
            ms.MS_SSIM = type("MS_SSIM", (torch.nn.Module,), {})
            sys.modules["pytorch_msssim"] = ms


def import_reference(dropin: bool = False):
    """Import the reference's `dgmr` package (fresh), returning the `dgmr.dgmr` module.  dropin: the reference's wrapper, this
    repository's modules underneath."""
    root = reference_root()
    if root is None:
        raise RuntimeError("reference not available: neither baseline/_ref nor /root/reference holds dgmr/dgmr.py")
    _install_stubs()
    for k in [k for k in sys.modules if k == "dgmr" or k.startswith("dgmr.")]:
        del sys.modules[k]
    pkg = types.ModuleType("dgmr")           # bare package object: skips dgmr/__init__.py's eager imports
    pkg.__path__ = [os.path.join(root, "dgmr")]
    sys.modules["dgmr"] = pkg
    if dropin:
        if ROOT not in sys.path:
            sys.path.insert(0, ROOT)
        import skillful_nowcasting_b200 as B
        from skillful_nowcasting_b200 import common, discriminators, generators

        sys.modules["dgmr.common"] = common
        sys.modules["dgmr.generators"] = generators
        sys.modules["dgmr.discriminators"] = discriminators
        assert B.Generator is generators.Generator
    import dgmr.dgmr as wrapper   # the reference's own file either way

    assert os.path.realpath(wrapper.__file__).startswith(os.path.realpath(root)), wrapper.__file__
    return wrapper


def build_dgmr(cfg: dict, generation_steps: int = 1, dropin: bool = False, anomaly: bool = True, seed: int = 0):
    """`dgmr.DGMR` of the reference with the benchmark's widths.  The constructor switches autograd anomaly detection on globally
    (ref: dgmr/dgmr.py:130) -- `anomaly=False` switches it back off afterwards (reported as a separate measurement)."""
    wrapper = import_reference(dropin)
    torch.manual_seed(seed)
    model = wrapper.DGMR(forecast_steps=cfg["forecast_steps"], input_channels=1, output_shape=cfg["output_shape"],
                         latent_channels=cfg["latent_channels"], context_channels=cfg["context_channels"],
                         generation_steps=generation_steps)
    torch.autograd.set_detect_anomaly(bool(anomaly))
    return model


def training_step_fn(model, x, y):
    """One call of the reference's public training API on a resident batch."""
    def step():
        model.training_step((x, y), 0)
        return model.logged
    return step


def minimal_step_fn(model, x, y):
    """The parity-preserving minimal schedule on the REFERENCE's modules: what this repository's native step driver runs
    (skillful_nowcasting_b200.training.gan_step), restated with the reference's own generator / discriminator / losses / torch.optim.Adam
    -- so that the native arm can be compared with the reference arithmetic under the same schedule, not only under the wrapper's
    literal one (which back-propagates the generator in the D phase, recomputes it under checkpoint and runs a trailing forward)."""
    import dgmr.dgmr as wrapper

    g_opt, d_opt = model.optimizers()
    b = x.shape[0]

    def step():
        real = torch.cat([x, y], dim=1)
        for _ in range(2):
            d_opt.zero_grad()
            with torch.no_grad():
                pred = model.generator(x)
            out = model.discriminator(torch.cat([real, torch.cat([x, pred], dim=1)], dim=0))
            sr, sg = out[:b], out[b:]
            d_loss = wrapper.loss_hinge_disc(sg[:, 0:1], sr[:, 0:1]) + wrapper.loss_hinge_disc(sg[:, 1:2], sr[:, 1:2])
            d_loss.backward()
            d_opt.step()
        g_opt.zero_grad()
        d_params = [p for p in model.discriminator.parameters()]
        for p in d_params:
            p.requires_grad_(False)
        try:
            preds = [model.generator(x) for _ in range(model.generation_steps)]
            grid = model.grid_regularizer(torch.stack(preds, dim=0).mean(dim=0), y)
            scores = [model.discriminator(torch.cat([real, torch.cat([x, p_], dim=1)], dim=0))[b:] for p_ in preds]
            g_loss = wrapper.loss_hinge_gen(torch.cat(scores, dim=0)) + model.grid_lambda * grid
            g_loss.backward()
        finally:
            for p in d_params:
                p.requires_grad_(True)
        g_opt.step()
        return {"train/d_loss": d_loss.detach(), "train/g_loss": g_loss.detach(), "train/grid_loss": grid.detach()}
    return step
