"""Block-level parity cases shared by the host-logic tests (CPU, emulated ABI) and the GPU tests (real CUDA library).
Each case builds a B200 block, copies its state into the oracle, and compares forward values, mutated buffers and all
gradients on seeded inputs.  Shapes follow the reference's own block tests (tests/test_model.py:29-81) scaled to run in
seconds; the `wide` cases use channel counts that the tcgen05 path serves."""
import torch

from oracle import dgmr_oracle as O
from parity_util import assert_grads_close, rel_err


def block_cases(wide: bool):
    from skillful_nowcasting_b200.common import DBlock, GBlock, LBlock, UpsampleGBlock

    c = 32 if wide else 8   # base channel count
    s = 16 if wide else 8
    return [
        ("g", lambda: GBlock(2 * c, 2 * c), lambda st, x, tr: O.g_block(st, "m", x, tr), (2, 2 * c, s, s)),
        ("g_proj", lambda: GBlock(2 * c, c), lambda st, x, tr: O.g_block(st, "m", x, tr), (2, 2 * c, s, s)),
        ("upg", lambda: UpsampleGBlock(2 * c, c), lambda st, x, tr: O.upsample_g_block(st, "m", x, tr), (2, 2 * c, s, s)),
        ("d", lambda: DBlock(c, 2 * c), lambda st, x, tr: O.d_block(st, "m", x, tr), (2, c, s, s)),
        ("d_first", lambda: DBlock(4, 2 * c, first_relu=False), lambda st, x, tr: O.d_block(st, "m", x, tr, first_relu=False), (2, 4, s, s)),
        ("d3", lambda: DBlock(c, 2 * c, conv_type="3d", first_relu=False), lambda st, x, tr: O.d_block(st, "m", x, tr, first_relu=False),
         (2, c, 5, s, s)),
        ("dkeep", lambda: DBlock(c, c, keep_same_output=True), lambda st, x, tr: O.d_block(st, "m", x, tr, keep_same_output=True), (2, c, 4, 4)),
        ("l", lambda: LBlock(c, 3 * c), lambda st, x, tr: O.l_block(st, "m", x), (1, c, 4, 4)),
    ]


def run_block_case(case, training, device, tol_fwd, tol_grad, tol_buf=None, tol_l2=None):
    _, make, ofn, shape = case
    torch.manual_seed(5)
    mod = make()
    mod.train(training)
    st = O.clone_state({"m." + k: v for k, v in mod.state_dict().items()}, requires_grad=True)
    x = torch.rand(shape)
    xo = x.clone().requires_grad_(True)
    ref = ofn(st, xo, training)
    mod.to(device)
    xm = x.clone().to(device).requires_grad_(True)
    got = mod(xm)
    assert got.shape == ref.shape
    assert rel_err(got, ref) < tol_fwd, rel_err(got, ref)
    w = torch.randn_like(ref)
    names = [k for k in st if st[k].requires_grad]
    rg = torch.autograd.grad((ref * w).sum(), [xo] + [st[k] for k in names], allow_unused=True)
    params = dict(mod.named_parameters())
    mg = torch.autograd.grad((got * w.to(device)).sum(), [xm] + [params[k[2:]] for k in names], allow_unused=True)
    assert_grads_close(["x"] + names, [None if g is None else g.cpu() for g in mg], rg, tol_grad, tol_l2=tol_l2)
    for k, v in mod.state_dict().items():
        if v.numel():
            assert rel_err(v, st["m." + k]) < (tol_buf or tol_fwd), k


def run_conv_gru_case(device, tol_fwd, tol_grad, cx=24, ch=8, s=8, T=4, tol_l2=None):
    from skillful_nowcasting_b200.layers import ConvGRU

    torch.manual_seed(6)
    gru = ConvGRU(cx + ch, ch)
    st = O.clone_state({"g." + k: v for k, v in gru.state_dict().items()}, requires_grad=True)
    xs = [torch.rand(2, cx, s, s, requires_grad=True) for _ in range(T)]
    h = torch.rand(2, ch, s, s)
    ref = O.conv_gru(st, "g", xs, h, True)
    gru.to(device)
    xs2 = [x.detach().clone().to(device).requires_grad_(True) for x in xs]
    got = gru(xs2, h.to(device))
    assert got.shape == (T, 2, ch, s, s)
    assert rel_err(got, ref) < tol_fwd, rel_err(got, ref)
    w = torch.randn_like(ref)
    names = [k for k in st if st[k].requires_grad]
    rg = torch.autograd.grad((ref * w).sum(), xs + [st[k] for k in names])
    params = dict(gru.named_parameters())
    mg = torch.autograd.grad((got * w.to(device)).sum(), xs2 + [params[k[2:]] for k in names])
    assert_grads_close([f"x{i}" for i in range(T)] + names, [g.cpu() for g in mg], rg, tol_grad, tol_l2=tol_l2)
    for k, v in gru.state_dict().items():
        assert rel_err(v, st["g." + k]) < max(tol_fwd, 2e-4) or not (k.endswith("_u") or k.endswith("_v")), k
    return gru, xs2, h
