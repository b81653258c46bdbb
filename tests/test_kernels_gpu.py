"""Per-kernel parity: every C-ABI entry point on the GPU against the host emulator (tests/emu_backend.py,
the plain-torch restatement of each kernel) on identical seeded inputs.  Tolerances are stated per test:
bit-exact for index permutations, 1e-5-ish for fp32 pointwise / reductions, TF32-level for tensor-core convs.
"""
import pytest
import torch

from emu_backend import EmuBackend

pytestmark = pytest.mark.gpu


def _both(method, args, cuda_backend, rtol=1e-5, atol=1e-6, kwargs=None, check=None):
    """Run `method` on the emulator (CPU copies) and on CUDA (device copies); compare every tensor argument."""
    kwargs = kwargs or {}
    emu = EmuBackend()
    a_cpu = [a.clone() if torch.is_tensor(a) else a for a in args]
    a_gpu = [a.cuda() if torch.is_tensor(a) else a for a in args]
    getattr(emu, method)(*a_cpu, **kwargs)
    getattr(cuda_backend, method)(*a_gpu, **{k: (v.cuda() if torch.is_tensor(v) else v) for k, v in kwargs.items()})
    torch.cuda.synchronize()
    for i, (c, g) in enumerate(zip(a_cpu, a_gpu)):
        if torch.is_tensor(c) and (check is None or i in check):
            g = g.cpu()
            if c.dtype == torch.float64:
                c, g = c.float(), g.float()
            err = (c - g).abs().max().item() if c.numel() else 0.0
            ref = c.abs().max().item() if c.numel() else 0.0
            assert err <= atol + rtol * max(ref, 1e-30), f"{method}: arg {i} max err {err:.3e} (ref max {ref:.3e})"


def test_permute_and_reduce(cuda_backend):
    torch.manual_seed(0)
    src = torch.randn(2 * 3 * 4 * 5)
    dst = torch.zeros(2 * 3 * 4 * 5)
    _both("permute", [src, dst, (2, 3, 4, 5), (60, 20, 5, 1), (5, 10, 30, 1)], cuda_backend, rtol=0, atol=0)
    dst = torch.randn(130)
    _both("permute", [src, dst, (4, 5), (5, 1), (10, 1), True, 7, 3], cuda_backend, rtol=0, atol=1e-7)
    x = torch.randn(3 * 37 * 50)
    _both("reduce_mid", [x, torch.zeros(150), 3, 37, 50], cuda_backend)
    # float4 form (contiguous innermost dimension, 16-byte aligned offsets): the depth-fold copy, a zero-stride broadcast, accumulate
    src = torch.randn(2 * 5 * 6 * 8)
    dst = torch.zeros(2 * 5 * 6 * 24)
    _both("permute", [src, dst, (2, 4, 6, 8), (240, 48, 8, 1), (720, 144, 24, 1), False, 48, 8], cuda_backend, rtol=0, atol=0)
    _both("permute", [torch.zeros(8), torch.randn(2 * 5 * 6 * 24), (2, 6, 8), (0, 0, 1), (720, 24, 1), False, 0, 16], cuda_backend, rtol=0, atol=0)
    _both("permute", [src, torch.randn(2 * 5 * 6 * 24), (2, 5, 6, 8), (240, 48, 8, 1), (720, 144, 24, 1), True, 0, 16], cuda_backend, rtol=0, atol=1e-7)
    # misaligned offset: must take the scalar path
    _both("permute", [src, torch.randn(2 * 5 * 6 * 24), (2, 5, 6, 8), (240, 48, 8, 1), (720, 144, 24, 1), False, 0, 3], cuda_backend, rtol=0, atol=0)


def test_rowdot_div(cuda_backend):
    """dL/ds[co] = <dW[co], W[co]> / s[co] over a channel slice of an OIHW weight (dgmr_rowdot_div)."""
    torch.manual_seed(3)
    cout, cintot, taps, ci0, cin = 37, 24, 9, 8, 12
    a, b = torch.randn(cout, cintot, taps), torch.randn(cout, cintot, taps)
    den = torch.rand(cout) + 0.5
    _both("rowdot_div", [a, b, den, torch.empty(1, cout), cout, cin * taps, cintot * taps, ci0 * taps], cuda_backend, rtol=1e-5, atol=1e-5)
    _both("rowdot_div", [a, b, None, torch.empty(cout), cout, cintot * taps, cintot * taps, 0], cuda_backend, rtol=1e-5, atol=1e-5)


def test_pointwise(cuda_backend):
    torch.manual_seed(1)
    n = 10007
    x, y = torch.randn(n), torch.randn(n)
    _both("axpby", [0.5, x, -2.0, y, torch.empty(n)], cuda_backend)
    _both("axpby", [0.25, x, 0.0, None, torch.empty(n)], cuda_backend)
    _both("relu_fwd", [x, torch.empty(n)], cuda_backend, rtol=0, atol=0)
    _both("relu_bwd", [y, x, torch.empty(n)], cuda_backend, rtol=0, atol=0)
    _both("split_tf32", [x, torch.empty(n), torch.empty(n)], cuda_backend, rtol=0, atol=0)
    _both("fill", [torch.empty(n), 3.5], cuda_backend, rtol=0, atol=0)
    _both("round_tf32", [torch.randn(n) * 100], cuda_backend, rtol=0, atol=0)   # cvt.rna.tf32.f32, bit-exact


@pytest.mark.parametrize("dims,win", [((2, 1, 8, 12, 5), (1, 2, 2)), ((2, 5, 6, 6, 3), (2, 2, 2)), ((1, 22, 4, 4, 1), (1, 2, 2)),
                                      ((3, 5, 16, 24, 1), (1, 2, 2)), ((2, 3, 6, 6, 1), (1, 2, 2))])   # single channel: float4 form / W % 4 != 0 fallback
def test_pool_upsample(cuda_backend, dims, win):
    torch.manual_seed(2)
    n, d, h, w, c = dims
    pd, ph, pw = win
    x = torch.randn(dims)
    y = torch.empty(n, d // pd, h // ph, w // pw, c)
    _both("pool_sum", [x, y, n, d, h, w, c, pd, ph, pw, 0.25], cuda_backend)
    g = torch.randn(n, d // pd, h // ph, w // pw, c)
    _both("upsample", [g, torch.empty(dims), n, d // pd, h // ph, w // pw, c, pd, ph, pw, d, h, w, 0.5], cuda_backend)


def test_gru_pointwise(cuda_backend):
    torch.manual_seed(3)
    rows, ch = 300, 24
    pre, h, c, g = (torch.randn(rows, ch) for _ in range(4))
    _both("gru_gate_fwd", [pre, ch, h, torch.empty(rows, ch), rows, ch, 0], cuda_backend, atol=1e-6)
    _both("gru_gate_fwd", [pre, ch, h, torch.empty(rows, ch), rows, ch, 256], cuda_backend, atol=4e-3)   # 1 tf32 ulp: a 1-ulp fp32 difference may round the other way
    for relu_c in (False, True):
        _both("gru_blend_fwd", [pre, ch, h, c, torch.empty(rows, ch), None, rows, ch, relu_c], cuda_backend, atol=1e-6)
        _both("gru_blend_fwd", [pre, ch, h, c, torch.empty(rows, ch), torch.empty(rows, ch), rows, ch, relu_c], cuda_backend, atol=4e-3)
        _both("gru_blend_bwd", [g, pre, ch, h, c, torch.empty(rows, ch), ch, torch.empty(rows, ch), torch.empty(rows, ch), False, rows, ch, relu_c],
              cuda_backend, atol=1e-6)
    _both("gru_gate_bwd", [g, pre, ch, h, torch.empty(rows, ch), ch, torch.empty(rows, ch), False, rows, ch], cuda_backend, atol=1e-6)
    # deferred x part: pre-activation = (h part, in the buffer) + (x part), completed in place; r | u side by side (pitch 2 ch) and an odd width
    for chx in (24, 7):
        pru, xru, hh, cc, xcc = torch.randn(rows, 2 * chx), torch.randn(rows, 2 * chx), torch.randn(rows, chx), torch.randn(rows, chx), torch.randn(rows, chx)
        _both("gru_gate_fwd", [pru, 2 * chx, hh, torch.empty(rows, chx), rows, chx, 0], cuda_backend, atol=1e-6, kwargs=dict(x_r=xru))
        for x_u, x_c in ((xru.view(-1)[chx:], xcc), (None, xcc), (xru.view(-1)[chx:], None)):
            _both("gru_blend_fwd", [pru.view(-1)[chx:].clone(), 2 * chx, hh, cc, torch.empty(rows, chx), torch.empty(rows, chx), rows, chx, True], cuda_backend,
                  atol=2e-6, kwargs=dict(x_u=x_u, x_c=x_c), check=(0, 3, 4))
    # fused backward operands: dz = d_pre * scale[c], fp32 and tf32-rounded
    sc, sc2 = torch.rand(ch) + 0.5, torch.rand(ch) + 0.5
    for rnd in (False, True):
        tol = dict(rtol=6e-4, atol=1e-6) if rnd else dict(atol=1e-6)
        _both("gru_gate_bwd", [g, pre, ch, h, torch.empty(rows, ch), ch, torch.empty(rows, ch), False, rows, ch, sc, torch.empty(rows, ch), rnd], cuda_backend, **tol)
        _both("gru_blend_bwd", [g, pre, ch, h, c, torch.empty(rows, ch), ch, torch.empty(rows, ch), torch.empty(rows, ch), False, rows, ch, True,
                                sc, torch.empty(rows, ch), sc2, torch.empty(rows, ch), rnd], cuda_backend, **tol)


@pytest.mark.parametrize("G,rows,C,relu,up2", [(1, 500, 24, True, False), (3, 2 * 8 * 8, 96, True, True), (4, 6, 768, False, False), (2, 4 * 4, 4, True, True),
                                               (2, 700000, 8, True, False), (1, 2 * 512 * 512, 8, True, True),   # long chunks: the multi-row main loops
                                               (18, 64, 40, True, False)])              # G > the finalize kernel's 8 group lanes, C not a multiple of 32
def test_batchnorm(cuda_backend, G, rows, C, relu, up2):
    torch.manual_seed(4)
    H = W = int((rows // 2) ** 0.5) if up2 else 1
    x = torch.randn(G * rows, C) * 2 + 0.5
    sums = torch.zeros(G, C, 2, dtype=torch.float64)
    _both("bn_stats", [x, sums, rows, G, C], cuda_backend, rtol=1e-5)
    EmuBackend().bn_stats(x, sums, rows, G, C)
    gamma, beta = torch.rand(C) + 0.5, torch.randn(C)
    for training in (True, False):
        rm, rv = torch.randn(C) * 0.1, torch.rand(C) + 0.5
        outs = [torch.empty(G, C) for _ in range(4)]
        _both("bn_finalize", [sums, gamma, beta, rm, rv, rows, G, C, 1e-5, 0.1, training] + outs, cuda_backend, rtol=1e-5, atol=1e-6)
    rm, rv = torch.zeros(C), torch.ones(C)
    mean, invstd, a, b = (torch.empty(G, C) for _ in range(4))
    EmuBackend().bn_finalize(sums, gamma, beta, rm, rv, rows, G, C, 1e-5, 0.1, True, mean, invstd, a, b)
    y = torch.empty(G * rows * (4 if up2 else 1), C)
    _both("bn_apply", [x, a, b, y, rows, G, C, relu, up2, H, W], cuda_backend, atol=1e-6)
    _both("bn_apply", [x, a, b, y, rows, G, C, int(relu) | 256, up2, H, W], cuda_backend, rtol=6e-4, atol=1e-6)  # tf32-rounded output
    if not up2:   # second output: the tf32-rounded copy of x itself (bit-exact: cvt.rna)
        xr = torch.empty_like(x)
        _both("bn_apply", [x, a, b, y, rows, G, C, int(relu) | 256, up2, H, W], cuda_backend, rtol=6e-4, atol=1e-6, kwargs=dict(x_rounded=None))
        emu_y, gpu_y, gpu_xr = torch.empty_like(y), torch.empty_like(y).cuda(), torch.empty_like(x).cuda()
        EmuBackend().bn_apply(x, a, b, emu_y, rows, G, C, int(relu), up2, H, W, x_rounded=xr)
        cuda_backend.bn_apply(x.cuda(), a.cuda(), b.cuda(), gpu_y, rows, G, C, int(relu), up2, H, W, x_rounded=gpu_xr)
        assert torch.equal(gpu_xr.cpu(), xr) and (gpu_y.cpu() - emu_y).abs().max().item() <= 1e-5 * max(emu_y.abs().max().item(), 1)
    dy = torch.randn_like(y)
    red = torch.zeros(G, C, 2, dtype=torch.float64)
    _both("bn_bwd_reduce", [dy, x, a, b, mean, invstd, red, rows, G, C, relu, up2, H, W], cuda_backend, rtol=1e-4, atol=1e-4)
    EmuBackend().bn_bwd_reduce(dy, x, a, b, mean, invstd, red, rows, G, C, relu, up2, H, W)
    for training in (True, False):
        _both("bn_bwd_apply", [dy, x, a, b, mean, invstd, None, red, torch.empty(G * rows, C), torch.empty(C), torch.empty(C), False,
                               rows, G, C, relu, up2, H, W, training], cuda_backend, rtol=1e-4, atol=1e-5)
    # fused forms: dx * out_scale[g, c] (+ dx_add), tf32-rounded (ops._ConvBNRelu / the BatchNorm branch node)
    osc, add = torch.rand(G, C) + 0.5, torch.randn(G * rows, C)
    _both("bn_bwd_apply", [dy, x, a, b, mean, invstd, osc, red, torch.empty(G * rows, C), torch.empty(C), torch.empty(C), False,
                           rows, G, C, relu, up2, H, W, True], cuda_backend, rtol=1e-4, atol=1e-5, kwargs=dict(dx_add=None))
    emu_dx, gpu_dx = torch.empty(G * rows, C), torch.empty(G * rows, C, device="cuda")
    EmuBackend().bn_bwd_apply(dy, x, a, b, mean, invstd, osc, red, emu_dx, None, None, False, rows, G, C, int(relu) | 256, up2, H, W, True, dx_add=add)
    dev = lambda t: t.cuda()
    cuda_backend.bn_bwd_apply(dev(dy), dev(x), dev(a), dev(b), dev(mean), dev(invstd), dev(osc), dev(red), gpu_dx, None, None, False, rows, G, C,
                              int(relu) | 256, up2, H, W, True, dx_add=dev(add))
    assert (gpu_dx.cpu() - emu_dx).abs().max().item() <= 1.1e-3 * emu_dx.abs().max().item()   # one tf32 ulp at rounding boundaries
    assert int((gpu_dx.view(torch.int32) & 0x1fff).abs().max()) == 0


@pytest.mark.parametrize("R,K,G", [(24, 36, 1), (48, 432, 4), (1, 768, 8), (384, 3456, 18), (768, 6912, 3), (96, 2592, 2)])
def test_spectral_norm(cuda_backend, R, K, G):
    torch.manual_seed(5)
    w = torch.randn(R, K) / K ** 0.5
    u = torch.nn.functional.normalize(torch.randn(R), dim=0)
    v = torch.nn.functional.normalize(torch.randn(K), dim=0)
    for training in (True, False):
        args = [w, u.clone(), v.clone(), R, K, G, 1e-4, training, torch.empty(G), torch.empty(G, R), torch.empty(G, K),
                torch.zeros((G + 2) * R + 2 * G + 8)]
        # power iteration is a contraction towards the top singular pair: fp32 reordering stays ~1e-6; u,v sign is fixed by the start
        _both("sn_power_iter", args, cuda_backend, rtol=2e-4, atol=2e-5, check={1, 2, 8, 9, 10})
    dis, isg = torch.randn(G), torch.rand(G) + 0.5
    uh, vh = torch.randn(G, R), torch.randn(G, K)
    _both("sn_bwd", [dis, isg, uh, vh, torch.zeros(R, K), R, K, G, False], cuda_backend, rtol=1e-4, atol=1e-5)


def test_pack_unpack(cuda_backend):
    torch.manual_seed(6)
    cout, cintot, taps = 10, 12, 9
    w = torch.randn(cout, cintot, taps)
    for mode in (0, 1, 256, 257):   # +256: tf32-rounded pack
        _both("pack_weight", [w, torch.empty(taps * cout * 8), cout, cintot, 4, 8, taps, mode], cuda_backend, rtol=0, atol=0)
    p = torch.randn(taps * cout * 8)
    _both("unpack_wgrad", [p, torch.zeros(cout, cintot, taps), cout, cintot, 4, 8, taps, False], cuda_backend, rtol=0, atol=0)
    _both("unpack_wgrad", [p, torch.randn(cout, cintot, taps), cout, cintot, 0, 8, taps, True], cuda_backend, rtol=0, atol=1e-6)


CONV_SHAPES_SIMT = [
    # N, D, H, W, Cin, Cout, kd, kh, kw, G
    (2, 1, 8, 8, 4, 24, 1, 3, 3, 1),
    (4, 1, 6, 10, 12, 20, 1, 3, 3, 2),
    (2, 1, 5, 7, 7, 3, 1, 1, 1, 1),
    (2, 4, 6, 6, 4, 8, 3, 3, 3, 1),
    (3, 1, 1, 1, 768, 1, 1, 1, 1, 3),
    (1, 1, 4, 4, 96, 48, 1, 3, 3, 1),
]


def _conv_args(shape, act, with_res, with_scale, with_bias=True):
    n, d, h, w, cin, cout, kd, kh, kw, g = shape
    taps = kd * kh * kw
    x = torch.randn(n, d, h, w, cin)
    wp = torch.randn(taps * cout * cin) / (taps * cin) ** 0.5
    bias = torch.randn(cout) if with_bias else None
    scale = (torch.rand(g, cout) + 0.5) if with_scale else None
    res = torch.randn(n, d, h, w, cout) if with_res else None
    y = torch.empty(n, d, h, w, cout)
    return [x, wp, bias, scale, res, y, n, d, h, w, cin, cout, kd, kh, kw, g, act]


@pytest.mark.parametrize("shape", CONV_SHAPES_SIMT)
@pytest.mark.parametrize("act,with_res,with_scale", [(0, False, False), (1, True, True)])
def test_conv_simt(cuda_backend, shape, act, with_res, with_scale):
    torch.manual_seed(7)
    args = _conv_args(shape, act, with_res, with_scale)
    _both("conv_fwd", args, cuda_backend, rtol=2e-5, atol=2e-5, kwargs=dict(algo=1))
    n, d, h, w, cin, cout, kd, kh, kw, g = shape
    x, dz = torch.randn(n, d, h, w, cin), torch.randn(n, d, h, w, cout)
    _both("conv_wgrad", [x, dz, torch.zeros(kd * kh * kw * cout * cin), n, d, h, w, cin, cout, kd, kh, kw], cuda_backend,
          rtol=5e-5, atol=5e-5, kwargs=dict(algo=1))


@pytest.mark.parametrize("G,rows,C,act", [(1, 300, 24, 0), (3, 64, 48, 1), (2, 17, 4, 1), (2, 40, 32, 257), (1, 33, 7, 1), (2, 4, 768, 0), (2, 5, 1024, 1), (3, 7, 1536, 0)])
def test_conv_bwd_prep(cuda_backend, G, rows, C, act):
    torch.manual_seed(8)
    dy, y, res = (torch.randn(G * rows, C) for _ in range(3))
    bias, scale = torch.randn(C), torch.rand(G, C) + 0.5
    args = [dy, y, res, bias, scale, torch.empty(G * rows, C), torch.empty(G * rows, C), torch.zeros(C), torch.zeros(G, C), rows, G, C, act]
    _both("conv_bwd_prep", args, cuda_backend, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("G,rows,C", [(1, 1000, 48), (2, 333, 96), (1, 45, 7)])
@pytest.mark.parametrize("variant", ["no_y", "y_no_dscale", "dscale_no_res", "many_rows"])
def test_conv_bwd_prep_variants(cuda_backend, G, rows, C, variant):
    """The template variants of the prologue: no y at all (scale / bias only: the G = 1 form whose scale gradient comes from dgmr_rowdot_div),
    ReLU mask without a scale gradient, scale gradient without a residual; long row chunks (several 4-row groups + tail, fp64 flushes)."""
    torch.manual_seed(28)
    if variant == "many_rows":
        rows = rows * 37
    dy, y, res = (torch.randn(G * rows, C) for _ in range(3))
    bias, scale = torch.randn(C), torch.rand(G, C) + 0.5
    dz, dpre, dbias, dscale = torch.empty(G * rows, C), torch.empty(G * rows, C), torch.zeros(C), torch.zeros(G, C)
    if variant == "no_y":
        args = [dy, None, None, bias, scale, dz, None, dbias, None, rows, G, C, 256]
    elif variant == "y_no_dscale":
        args = [dy, y, None, bias, scale, dz, dpre, dbias, None, rows, G, C, 1]
    elif variant == "dscale_no_res":
        args = [dy, y, None, bias, scale, dz, None, None, dscale, rows, G, C, 257]
    else:
        args = [dy, y, res, bias, scale, dz, dpre, dbias, dscale, rows, G, C, 1]
    _both("conv_bwd_prep", args, cuda_backend, rtol=1e-4, atol=2e-4)


@pytest.mark.parametrize("G,n_img,hw,C,act", [(2, 4, (8, 12), 24, 0), (1, 3, (4, 4), 7, 1)])
def test_conv_bwd_prep_half_resolution_residual(cuda_backend, G, n_img, hw, C, act):
    """Backward prologue of a conv that ran with DGMR_FLAG_RES_UP2: the residual in the <dY, Y - b - res> reduction is the
    half-resolution tensor read at (h/2, w/2)."""
    torch.manual_seed(18)
    uh, uw = hw
    rows = (n_img // G) * uh * uw if n_img % G == 0 else None
    assert rows is not None or G == 1
    rows = (n_img * uh * uw) // G
    dy, y = torch.randn(G * rows, C), torch.randn(G * rows, C)
    res = torch.randn(n_img * (uh // 2) * (uw // 2), C)
    bias, scale = torch.randn(C), torch.rand(G, C) + 0.5
    args = [dy, y, res, bias, scale, torch.empty(G * rows, C), torch.empty(G * rows, C), torch.zeros(C), torch.zeros(G, C), rows, G, C, act]
    _both("conv_bwd_prep", args, cuda_backend, rtol=1e-4, atol=1e-4, kwargs=dict(up_hw=(uh, uw)))


@pytest.mark.parametrize("geom,C,act", [((1, 2, 2, 1, 8, 12), 24, 0), ((2, 2, 2, 5, 6, 6), 8, 1), ((2, 2, 2, 11, 4, 4), 7, 1)])
def test_conv_bwd_prep_pooled_gradient(cuda_backend, geom, C, act):
    """Backward prologue of a conv whose output was average-pooled (DBlock): dy is the POOLED gradient, read at (d/pd, h/ph, w/pw) and divided
    by the window; odd depths leave a floor-dropped rim with zero gradient."""
    torch.manual_seed(22)
    pd, ph, pw, D, H, W = geom
    N, G = 4, 2
    rows = (N // G) * D * H * W
    dyp = torch.randn(N * (D // pd) * (H // ph) * (W // pw), C)
    y = torch.randn(G * rows, C)
    bias, scale = torch.randn(C), torch.rand(G, C) + 0.5
    args = [dyp, y, None, bias, scale, torch.empty(G * rows, C), torch.empty(G * rows, C), torch.zeros(C), torch.zeros(G, C), rows, G, C, act]
    _both("conv_bwd_prep", args, cuda_backend, rtol=1e-4, atol=1e-4, kwargs=dict(pool=geom))


def test_pack_weight_multi(cuda_backend):
    """dgmr_pack_weight_multi: several packs in one launch, padded input channels and a Cout window of a wider destination,
    against dgmr_pack_weight (bit-exact: pure index maps + the same cvt.rna rounding)."""
    import ctypes

    from skillful_nowcasting_b200._lib import PackItem

    be = cuda_backend
    torch.manual_seed(19)
    specs = [(24, 16, 0, 16, 9, 0), (24, 16, 4, 8, 9, 1 | 256), (48, 4, 0, 4, 27, 0 | 256), (40, 40, 8, 32, 1, 1)]
    items, checks = [], []
    for cout, cintot, ci0, cin, taps, mode in specs:
        w = torch.randn(cout, cintot, taps, device="cuda")
        pad = 8 if cin == 4 else cin
        cot, co0 = cout + 16, 8
        dst = torch.zeros(taps * cot * pad, device="cuda")
        items.append(dict(w=w, packed=dst, Cout=cout, CinTot=cintot, ci0=ci0, Cin=cin, taps=taps, mode=mode, CinPad=pad, co0=co0, CoutTot=cot))
        single = torch.empty(taps * cout * cin, device="cuda")
        be.pack_weight(w, single, cout, cintot, ci0, cin, taps, mode)
        checks.append((dst, single, cout, cin, taps, mode & 1, pad, co0, cot))
    be.pack_weight_multi(items)
    torch.cuda.synchronize()
    for dst, single, cout, cin, taps, mode, pad, co0, cot in checks:
        if mode == 0:
            got = dst.view(taps, cot, pad)[:, co0:co0 + cout, :cin]
            ref = single.view(taps, cout, cin)
            rest = dst.view(taps, cot, pad).clone(); rest[:, co0:co0 + cout, :cin] = 0
        else:
            got = dst.view(taps, pad, cot)[:, :cin, co0:co0 + cout]
            ref = single.view(taps, cin, cout)
            rest = dst.view(taps, pad, cot).clone(); rest[:, :cin, co0:co0 + cout] = 0
        assert torch.equal(got, ref)
        assert float(rest.abs().max()) == 0.0     # nothing outside the window was written


def test_sn_bwd_multi(cuda_backend):
    """dgmr_sn_bwd_multi against dgmr_sn_bwd weight by weight (same arithmetic order per element: bit-exact)."""
    be = cuda_backend
    torch.manual_seed(20)
    items, refs = [], []
    for R, K, G, acc in ((24, 72, 3, False), (8, 9, 1, True), (96, 864, 18, False), (16, 64, 5, True), (4, 36, 70, True)):
        t = [torch.randn(G, device="cuda"), torch.rand(G, device="cuda") + 0.5, torch.randn(G, R, device="cuda"), torch.randn(G, K, device="cuda")]
        dw0 = torch.randn(R, K, device="cuda")
        ref, dw = dw0.clone(), dw0.clone()
        be.sn_bwd(t[0], t[1], t[2], t[3], ref, R, K, G, acc)
        items.append(dict(d_inv_sigma=t[0], inv_sigma=t[1], u_hist=t[2], v_hist=t[3], dw=dw, R=R, K=K, G=G, accumulate=acc))
        refs.append(ref)
    be.sn_bwd_multi(items)
    torch.cuda.synchronize()
    for it, ref in zip(items, refs):
        assert torch.equal(it["dw"], ref)


def test_head_attention_losses_adam(cuda_backend):
    torch.manual_seed(9)
    x = torch.randn(6, 4, 40)
    _both("sumpool_relu_fwd", [x, torch.empty(6, 40), 6, 4, 40], cuda_backend)
    _both("sumpool_relu_bwd", [torch.randn(6, 40), x, torch.empty(6, 4, 40), 6, 4, 40], cuda_backend, rtol=0, atol=0)
    for (B, H, W, C) in [(1, 4, 4, 12), (2, 8, 8, 24)]:
        q, k, v, do = (torch.randn(B, 1, H, W, C) for _ in range(4))
        L = C * H
        beta = torch.empty(B, L, L)
        _both("attention_fwd", [q, k, v, torch.empty(B, 1, H, W, C), beta, B, H, W, C], cuda_backend, rtol=1e-4, atol=1e-5)
        EmuBackend().attention_fwd(q, k, v, torch.empty(B, 1, H, W, C), beta, B, H, W, C)
        _both("attention_bwd", [do, q, k, v, beta] + [torch.empty(B, 1, H, W, C) for _ in range(3)] + [torch.empty(B, L, L), B, H, W, C],
              cuda_backend, rtol=1e-4, atol=1e-5)
    s = torch.randn(8, 2) * 2
    _both("hinge_disc", [s, 4, 2, torch.empty(()), torch.empty(8, 2)], cuda_backend)
    _both("hinge_disc", [s.reshape(16, 1), 8, 1, torch.empty(()), torch.empty(16, 1)], cuda_backend)
    _both("hinge_gen", [s, 16, torch.empty(()), torch.empty(8, 2)], cuda_backend)
    gen, tgt = torch.rand(2, 4, 1, 32, 32) * 30, torch.rand(2, 4, 1, 32, 32) * 30
    _both("grid_cell_fwd", [gen, tgt, 24.0, 256.0, torch.empty(()), torch.zeros(1, dtype=torch.float64)], cuda_backend, check={4})
    _both("grid_cell_bwd", [gen, tgt, 24.0, 256.0, torch.tensor(0.7), torch.empty_like(gen)], cuda_backend)
    n = 5000
    p, g, m, vv = torch.randn(n), torch.randn(n), torch.rand(n), torch.rand(n)
    _both("adam", [p, g, m, vv, 2e-4, 0.0, 0.999, 1e-8, 3, 0.5], cuda_backend, rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("training", [True, False])
def test_spectral_norm_multi(cuda_backend, training):
    """All weights of a module in one launch (CTA groups per weight) vs the per-weight emulator."""
    torch.manual_seed(10)
    specs = [(768, 6912, 18), (384, 3456, 18), (384, 10368, 18), (192, 1728, 18), (96, 864, 18), (48, 432, 18), (768, 384, 18),
             (24, 36, 4), (1, 768, 8), (96, 2592, 1), (4, 48, 18)] + [(96, 96 * 9, 18)] * 20
    def make(dev):
        torch.manual_seed(11)
        items = []
        for (R, K, G) in specs:
            w = torch.randn(R, K) / K ** 0.5
            u = torch.nn.functional.normalize(torch.randn(R), dim=0)
            v = torch.nn.functional.normalize(torch.randn(K), dim=0)
            it = dict(w=w, u=u, v=v, R=R, K=K, G=G, eps=1e-4, training=training, inv_sigma=torch.empty(G), u_hist=torch.empty(G, R),
                      v_hist=torch.empty(G, K), ws=torch.zeros((G + 2) * R + 2 * G + 8))
            items.append({k: (t.to(dev) if torch.is_tensor(t) else t) for k, t in it.items()})
        return items
    ref, got = make("cpu"), make("cuda")
    EmuBackend().sn_power_iter_multi(ref)
    cuda_backend.sn_power_iter_multi(got)
    torch.cuda.synchronize()
    for i, (a, b) in enumerate(zip(ref, got)):
        for k in ("inv_sigma", "u_hist", "v_hist", "u", "v"):
            err = (a[k] - b[k].cpu()).abs().max().item()
            assert err <= 2e-5 + 2e-4 * a[k].abs().max().item(), (i, specs[i], k, err)
