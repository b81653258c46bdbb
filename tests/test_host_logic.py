"""Host logic of the drop-in modules (wiring, timestep/frame folding into groups, autograd plumbing, state-dict
contract), exercised WITHOUT a GPU by routing the C-ABI calls to the test-only emulator (tests/emu_backend.py),
and compared with the oracle / golden fixture.  The same comparisons run against the real CUDA library in
tests/test_parity_gpu.py."""
import io

import pytest
import torch

from block_cases import block_cases, run_block_case, run_conv_gru_case
from oracle import dgmr_oracle as O
from parity_util import (C1, GOLDEN, assert_grads_close, build_gan, c1_inputs, compare_grads, module_gan_forward,
                         oracle_gan_forward, rel_err)


@pytest.mark.parametrize("mode", ["eval", "train"])
def test_c1_gan_matches_fixture(emu, mode):
    gen, disc = build_gan(C1, seed=0, gamma=0.5)
    fix = torch.load(GOLDEN)
    x, y = c1_inputs()
    training = mode == "train"
    got = module_gan_forward(gen, disc, x, y, training, seed=fix["seed"], device="cpu")
    tol = 5e-4 if training else 1e-5
    assert rel_err(got["out"], fix[mode]["out"]) < tol
    assert rel_err(got["scores"], fix[mode]["scores"]) < tol * 5
    for k in ("d_loss", "grid", "g_loss"):
        assert rel_err(got[k], fix[mode][k]) < tol * 5, k
    if training:
        for sd, after in ((gen.state_dict(), fix[mode]["g_state_after"]), (disc.state_dict(), fix[mode]["d_state_after"])):
            for k, v in after.items():
                if "num_batches" in k:
                    assert int(sd[k]) == int(v), k
                else:
                    assert rel_err(sd[k], v) < 2e-4, k
        compare_grads(got["d_grads"], fix[mode]["d_grads"], 2e-3, 2e-2, zero_floor=1e-6)
        # G gradients through the whole train-mode net are chaotic at the 1e-2 level even reference-vs-reference
        compare_grads(got["g_grads"], fix[mode]["g_grads"], 5e-2, 2e-1, zero_floor=1e-5)


@pytest.mark.parametrize("case", block_cases(False), ids=lambda c: c[0])
@pytest.mark.parametrize("training", [True, False])
def test_block_forward_backward(emu, case, training):
    """Mirrors the reference's block smoke tests (tests/test_model.py:29-48) but checks values, buffers and grads."""
    run_block_case(case, training, "cpu", 1e-5, 2e-4)


@pytest.mark.parametrize("fused", [True, False])
def test_conv_gru_matches_oracle(emu, fused):
    """ref test shape family: tests/test_model.py:51-81 (scaled down).  fused: read|update gates as one conv and the whole
    recurrence as one autograd node (ops.gru_sequence); False: the per-step wiring."""
    from skillful_nowcasting_b200 import ops
    old = ops.config.gru_sequence
    ops.config.gru_sequence = fused
    try:
        gru, xs, h = run_conv_gru_case("cpu", 1e-5, 2e-4)
        out, new = gru.cell(xs[0].detach(), h)
    finally:
        ops.config.gru_sequence = old
    assert out.shape == (2, 8, 8, 8) and torch.equal(out, new)


def test_state_dict_contract_and_hub_roundtrip(emu, tmp_path):
    """State-dict keys of the reference and save_pretrained / from_pretrained round trip
    (ref: tests/test_model.py:341-399)."""
    import skillful_nowcasting_b200 as B

    smp = B.Sampler(forecast_steps=2, latent_channels=64, context_channels=32)
    keys = set(smp.state_dict().keys())
    for k in ("convGRU1.cell.read_gate_conv.bias", "convGRU1.cell.read_gate_conv.parametrizations.weight.original",
              "convGRU1.cell.output_conv.parametrizations.weight.0._u", "gru_conv_1x1_4.parametrizations.weight.0._v",
              "g1.bn1.running_mean", "g1.bn1.num_batches_tracked", "up_g4.conv_1x1.bias", "bn.weight", "conv_1x1.bias"):
        assert k in keys, k
    assert len(B.Sampler().state_dict()) == 249 and len(B.Discriminator(1).state_dict()) == 162
    ctx = B.ContextConditioningStack(input_channels=1, output_channels=96)
    ctx.save_pretrained(tmp_path / "ctx")
    ctx2 = B.ContextConditioningStack.from_pretrained(tmp_path / "ctx")
    assert ctx2.training is False
    for (ka, va), (kb, vb) in zip(ctx.state_dict().items(), ctx2.state_dict().items()):
        assert ka == kb and torch.equal(va, vb)
    lat = B.LatentConditioningStack(shape=(8, 2, 2), output_channels=288)
    lat.save_pretrained(tmp_path / "lat")
    lat2 = B.LatentConditioningStack.from_pretrained(tmp_path / "lat")
    assert tuple(lat2.shape) == (8, 2, 2)
    d = B.Discriminator(input_channels=1)
    buf = io.BytesIO()
    torch.save(d.state_dict(), buf)
    buf.seek(0)
    B.Discriminator(input_channels=1).load_state_dict(torch.load(buf))


def test_shapes_of_reference_smoke_tests(emu):
    """ref: tests/test_model.py:84-140 (output shapes of the stacks and discriminators), at reduced size."""
    import skillful_nowcasting_b200 as B

    lat = B.LatentConditioningStack(shape=(8, 2, 2), output_channels=288)
    assert lat(torch.rand(1, 4, 1, 64, 64)).shape == (1, 288, 2, 2)
    ctx = B.ContextConditioningStack(input_channels=1, output_channels=96)
    outs = ctx(torch.rand(2, 4, 1, 64, 64))
    assert [tuple(o.shape) for o in outs] == [(2, 12, 16, 16), (2, 24, 8, 8), (2, 48, 4, 4), (2, 96, 2, 2)]


def test_dgmr_training_step_runs_reference_schedule(emu):
    """ref: tests/test_model.py:285-306 (config C1 at reduced width) + the literal schedule of dgmr/dgmr.py:137-218."""
    import skillful_nowcasting_b200 as B

    torch.manual_seed(0)
    model = B.DGMR(forecast_steps=2, output_shape=128, latent_channels=288, context_channels=32, generation_steps=1)
    x, y = torch.rand(2, 4, 1, 128, 128), torch.rand(2, 2, 1, 128, 128)
    g_before = [p.detach().clone() for p in model.generator.parameters()]
    d_before = [p.detach().clone() for p in model.discriminator.parameters()]
    model.training_step((x, y), 0)
    assert set(model.logged) == {"train/d_loss", "train/g_loss", "train/grid_loss"}
    assert all(torch.isfinite(v) and v.dim() == 0 for v in model.logged.values())
    g_changed = sum(not torch.equal(a, b) for a, b in zip(g_before, model.generator.parameters()))
    d_changed = sum(not torch.equal(a, b) for a, b in zip(d_before, model.discriminator.parameters()))
    # parameters that never receive gradients stay put (g*.conv_1x1 w+b; attention q/k/v/out while gamma == 0)
    assert g_changed >= len(g_before) - 12 and d_changed >= 30, (g_changed, len(g_before), d_changed, len(d_before))
    assert model(x).shape == (2, 2, 1, 128, 128)
