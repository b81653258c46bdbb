"""Parity at BASELINE.json's full layer sizes (B = 16, T = 18 -> 288 images; temporal D 32 x 22 frames), where the oracle cannot
run in seconds.  Two arms per shape, both through the C ABI with AUTO dispatch -- i.e. on the kernels the benchmark uses
(CTA-pair halo-patch, plain, persistent, row / tap-wise wgrad):

  (1) against torch's own fp32 convolution (`allow_tf32 = False`; test-only use of cuDNN as a checker) on operands that are
      pre-rounded to TF32.  With TF32-exact operands every product is exact in fp32, so the two sides differ only by how the
      partial sums are accumulated: a dropped border pixel, a wrong tap, a mis-scaled split-K partial or an unrounded operand show up at
      1e-3 or worse, far outside the tolerance.
  (2) the adjoint identities of a linear map,  <conv(x,w), dy> == <x, dgrad(dy,w)> == <w, wgrad(x,dy)>,  which tie the three kernels
      to each other without any reference.

Tolerances are DERIVED from the measured accumulation behaviour of tcgen05 kind::tf32 (test_tensor_core_accumulation below,
profiles/accum_probe_r02.txt): inside one MMA chain every addend down to 2^-23 of the running sum survives (the accumulator is a
genuine fp32), but the sum is TRUNCATED -- not rounded -- to fp32 after each MMA: a bias of about -0.5 ulp (3e-8 relative) per MMA in the
chain, towards zero.  A forward / dgrad output accumulates K/8 MMAs (K = Cin*taps), so its relative error against an exactly rounded
fp32 sum is ~ (K/8) * 2^-25: 3e-6 at K = 864 and 2.6e-5 at K = 6912 -- what round 2 measured (2.6e-6 and 1.96e-5).  A wgrad element
accumulates (pixels / 8 / split-K CTAs) MMAs per CTA, ~2000 with the two-wave split of the big layers: -1.0e-4 .. -2.2e-4 relative,
coherent -- that, not a dropped tap, is the 1.03e-4 adjoint mismatch round 1 left open (<w,dw> is low by exactly that bias while
<y,dy> and <x,dx> agree to 4e-8).  Parity mode (3xTF32) issues three MMAs per k-step and caps wgrad chains at 192 MMAs.
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

# N, D, H, W, Cin, Cout, kd, k  -> which kernels serve it at this size
FULL_SHAPES = [
    (288, 1, 128, 128, 96, 96, 1, 3),   # CTA-pair halo-patch kernel (fwd, dgrad), row wgrad
    (288, 1, 128, 128, 96, 48, 1, 3),   # plain kernel fwd (Cout 48), patch dgrad (48 -> 96 tail chunk), row wgrad
    (32, 22, 64, 64, 48, 48, 3, 3),     # temporal D: 3-D, channel-tail blocks, row wgrad with 9 (kd,kh) rows
    (128, 1, 128, 128, 8, 48, 1, 3),    # first D block: 8-channel (zero-padded) input, persistent short-K kernel
    (288, 1, 64, 64, 48, 96, 1, 1),     # 1x1 conv: persistent plain kernel, tap-wise wgrad
    (288, 1, 16, 16, 768, 768, 1, 3),   # widest layer: 3 N tiles, patch n/a (H*W < 1024) -> plain kernel, tap-wise wgrad
    (16, 1, 8, 8, 384, 768, 1, 3),      # ConvGRU level-1 gate conv of one step (tap-split accumulate mode is chosen by the host)
]
IDS = ["patchpair_96x96_128", "plain_96x48_128", "conv3d_48x48", "first_d_8x48", "conv1x1_48x96", "wide_768_16", "gru_l1"]

ULP_BIAS = 2.0 ** -25   # mean truncation loss per MMA, relative to the running sum


def tol_chain(mmas: float, floor: float = 1e-5) -> float:
    """max |y - ref| / max |ref| allowed for an accumulation chain of `mmas` truncating MMAs: 3x the mean bias (elements whose running sum
    stayed near its final magnitude for the whole chain lose the full 1 ulp per MMA, and max|err|/max|ref| picks the worst), plus the
    fp32 noise of the checker itself."""
    return max(floor, 3.0 * mmas * ULP_BIAS)


def wgrad_chain(n, d, h, w, cin, cout, kd, k, x3=False):
    """MMAs in one CTA's wgrad accumulation chain (mirrors the split-K rules of launch_conv_umma_wgrad[_row]; 148 SMs)."""
    pixels = n * d * h * w
    if x3:
        return 3 * 16 * 4
    if k == 3 and w % 32 == 0 and pixels >= 16384:            # row kernel: two waves of CTAs
        base = kd * 3 * -(-cout // 128) * -(-cin // 160)
        ksplit = max(1, min((148 * 2) // base, pixels // 32 // 8))
    else:                                                     # tap-wise kernel: three waves
        base = kd * k * k * -(-cout // 128) * -(-cin // 256)
        ksplit = max(1, min((148 * 3) // base, pixels // 32 // 4))
    return pixels / 8.0 / ksplit


def _dot(a, b):
    return float((a.double().flatten() * b.double().flatten()).sum())


def _torch_conv(x_cl, wt, kd, k):
    """fp32 reference on the channels-last tensor viewed as NC(D)HW (no copy)."""
    pad = k // 2
    if kd > 1:
        return F.conv3d(x_cl.permute(0, 4, 1, 2, 3), wt, padding=(kd // 2, pad, pad)).permute(0, 2, 3, 4, 1)
    return F.conv2d(x_cl[:, 0].permute(0, 3, 1, 2), wt, padding=pad).permute(0, 2, 3, 1).unsqueeze(1)


@pytest.fixture
def true_fp32():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _setup(be, shape):
    n, d, h, w, cin, cout, kd, k = shape
    taps = kd * k * k
    torch.manual_seed(5)
    x = torch.randn(n, d, h, w, cin, device="cuda")
    ks = (kd, k, k) if kd > 1 else (k, k)
    wt = torch.randn(cout, cin, *ks, device="cuda") / (taps * cin) ** 0.5
    be.round_tf32(x)
    be.round_tf32(wt)
    wp = torch.empty(taps * cout * cin, device="cuda")
    wpt = torch.empty(taps * cout * cin, device="cuda")
    be.pack_weight(wt, wp, cout, cin, 0, cin, taps, 0)
    be.pack_weight(wt, wpt, cout, cin, 0, cin, taps, 1)
    return x, wt, wp, wpt, taps


@pytest.mark.parametrize("shape", FULL_SHAPES, ids=IDS)
def test_full_size_layer_against_fp32_torch_and_adjoints(cuda_backend, true_fp32, shape):
    be = cuda_backend
    n, d, h, w, cin, cout, kd, k = shape
    x, wt, wp, wpt, taps = _setup(be, shape)
    # ---- forward
    y = torch.full((n, d, h, w, cout), float("nan"), device="cuda")
    be.conv_fwd(x, wp, None, None, None, y, n, d, h, w, cin, cout, kd, k, k, 1, 0)
    ref = _torch_conv(x, wt, kd, k)
    assert not torch.isnan(y).any(), "forward left outputs unwritten"
    e_fwd = float((y - ref).abs().max() / ref.abs().max())
    # ---- dgrad (the forward kernel on the flipped-transposed pack) and wgrad, against autograd of the fp32 torch conv
    dy = torch.empty_like(y)
    be.round_tf32(y, dy)
    dx = torch.full((n, d, h, w, cin), float("nan"), device="cuda")
    be.conv_fwd(dy, wpt, None, None, None, dx, n, d, h, w, cout, cin, kd, k, k, 1, 0)
    dwp = torch.empty(taps * cout * cin, device="cuda")
    be.conv_wgrad(x, dy, dwp, n, d, h, w, cin, cout, kd, k, k)
    del ref
    xr = x.clone().requires_grad_(True)
    wr = wt.clone().requires_grad_(True)
    gx, gw = torch.autograd.grad(_torch_conv(xr, wr, kd, k), [xr, wr], dy)
    assert not torch.isnan(dx).any(), "dgrad left outputs unwritten"
    e_dx = float((dx - gx).abs().max() / gx.abs().max())
    gwp = torch.empty_like(dwp)
    be.pack_weight(gw.contiguous(), gwp, cout, cin, 0, cin, taps, 0)
    e_dw = float((dwp - gwp).abs().max() / gwp.abs().max())
    # ---- adjoint identities
    a, b, c = _dot(y, dy), _dot(x, dx), _dot(wp, dwp)
    e_ab, e_ac = abs(a - b) / a, abs(a - c) / a
    print(f"\nFULLSIZE {shape}: fwd {e_fwd:.2e} dgrad {e_dx:.2e} wgrad {e_dw:.2e} | <y,dy> {a:.9e} <x,dx> {b:.9e} <w,dw> {c:.9e} "
          f"adj {e_ab:.2e} {e_ac:.2e}")
    t_f, t_d, t_w = tol_chain(cin * taps / 8), tol_chain(cout * taps / 8), tol_chain(wgrad_chain(*shape))
    print(f"   bounds: fwd {t_f:.1e} dgrad {t_d:.1e} wgrad {t_w:.1e}")
    assert e_fwd < t_f, f"forward vs fp32 torch: {e_fwd:.3e} (bound {t_f:.1e})"
    assert e_dx < t_d, f"dgrad vs fp32 torch: {e_dx:.3e} (bound {t_d:.1e})"
    assert e_dw < t_w, f"wgrad vs fp32 torch: {e_dw:.3e} (bound {t_w:.1e})"
    # <y,dy> and <x,dx> carry the (tiny) forward / dgrad bias, <w,dw> the wgrad chain bias
    assert a > 0 and e_ab < t_f + t_d and e_ac < t_f + t_w, f"adjoint identities: {e_ab:.3e} {e_ac:.3e}"


@pytest.mark.parametrize("shape", [FULL_SHAPES[0], FULL_SHAPES[2], FULL_SHAPES[5]], ids=[IDS[0], IDS[2], IDS[5]])
def test_full_size_layer_3xtf32_matches_fp32_on_unrounded_operands(cuda_backend, true_fp32, shape):
    """Parity mode (DGMR_PREC_3XTF32): arbitrary fp32 operands split into (hi, lo) pairs; forward, dgrad and wgrad must agree with
    torch's fp32 convolution to the accumulation bound (three truncating MMAs per k-step) -- 10-100x tighter than the 2^-11 operand
    rounding of 1xTF32.  Round 2 measured fwd 7e-6 / 1.1e-5 / 5.1e-5 at K = 864 / 1296 / 6912."""
    be = cuda_backend
    n, d, h, w, cin, cout, kd, k = shape
    taps = kd * k * k
    torch.manual_seed(7)
    x = torch.randn(n, d, h, w, cin, device="cuda")
    ks = (kd, k, k) if kd > 1 else (k, k)
    wt = torch.randn(cout, cin, *ks, device="cuda") / (taps * cin) ** 0.5
    wp = torch.empty(taps * cout * cin, device="cuda")
    wpt = torch.empty_like(wp)
    be.pack_weight(wt, wp, cout, cin, 0, cin, taps, 0)
    be.pack_weight(wt, wpt, cout, cin, 0, cin, taps, 1)

    def split(t):
        hi, lo = torch.empty_like(t), torch.empty_like(t)
        be.split_tf32(t, hi, lo)
        return hi, lo

    xh, xl = split(x)
    wh, wl = split(wp)
    y = torch.full((n, d, h, w, cout), float("nan"), device="cuda")
    be.conv_fwd(xh, wh, None, None, None, y, n, d, h, w, cin, cout, kd, k, k, 1, 0, precision=1, x_lo=xl, wp_lo=wl)
    ref = _torch_conv(x, wt, kd, k)
    e_fwd = float((y - ref).abs().max() / ref.abs().max())
    # dy = y: every gradient element is then a COHERENT sum (~ ||y||^2-like), which is what makes a relative comparison meaningful --
    # with an uncorrelated dy a weight-gradient element is a zero-mean sum of up to 4.7 M terms and any fp32 accumulation (torch's own
    # included) is only good to ~1e-4 of it (round 2 measured 3.5e-4 between this path and cuDNN's fp32 wgrad that way).
    dy = y.clone()
    dyh, dyl = split(dy)
    wth, wtl = split(wpt)
    dx = torch.full((n, d, h, w, cin), float("nan"), device="cuda")
    be.conv_fwd(dyh, wth, None, None, None, dx, n, d, h, w, cout, cin, kd, k, k, 1, 0, precision=1, x_lo=dyl, wp_lo=wtl)
    dwp = torch.empty(taps * cout * cin, device="cuda")
    be.conv_wgrad(xh, dyh, dwp, n, d, h, w, cin, cout, kd, k, k, precision=1, x_lo=xl, dz_lo=dyl)
    del ref
    xr = x.clone().requires_grad_(True)
    wr = wt.clone().requires_grad_(True)
    gx, gw = torch.autograd.grad(_torch_conv(xr, wr, kd, k), [xr, wr], dy)
    e_dx = float((dx - gx).abs().max() / gx.abs().max())
    gwp = torch.empty_like(dwp)
    be.pack_weight(gw.contiguous(), gwp, cout, cin, 0, cin, taps, 0)
    e_dw = float((dwp - gwp).abs().max() / gwp.abs().max())
    a, b, c = _dot(y, dy), _dot(x, dx), _dot(wp, dwp)
    e_ab, e_ac = abs(a - b) / a, abs(a - c) / a
    t_f, t_d, t_w = tol_chain(3 * cin * taps / 8), tol_chain(3 * cout * taps / 8), tol_chain(wgrad_chain(*shape, x3=True), floor=3e-5)
    print(f"\nFULLSIZE-3X {shape}: fwd {e_fwd:.2e} dgrad {e_dx:.2e} wgrad {e_dw:.2e} adj {e_ab:.2e} {e_ac:.2e}  bounds {t_f:.1e} {t_d:.1e} {t_w:.1e}")
    assert e_fwd < t_f and e_dx < t_d, (e_fwd, e_dx)
    assert e_dw < max(t_w, 1e-4), e_dw          # (the checker's own fp32 wgrad is not better than that over millions of pixels)
    assert e_ab < t_f + t_d and e_ac < t_f + t_w, (e_ab, e_ac)   # our three kernels against each other, no checker involved


def test_tensor_core_accumulation(cuda_backend):
    """How tcgen05 kind::tf32 accumulates (this is what the tolerances above are derived from).  One chain of 4 MMAs (K = 32) on
    exactly representable operands: row = [1, 2^-s, 2^-s, ...] . ones.  fp32 round-to-nearest keeps the 31 small addends down to
    s = 23; the tensor core must keep them at least down to s = 20 (i.e. >= 21 bits below the leading one survive the alignment)."""
    import ctypes

    be = cuda_backend
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    N = 16
    lines = []
    worst_kept = None
    for sign in ("same", "alternating"):
        for s in range(8, 26):
            A = torch.zeros(256, 32, device="cuda")
            A[:, 0] = 1.0
            A[:, 1:] = 2.0 ** -s
            if sign == "alternating":
                A[:, 2::2] *= -1.0
            Bm = torch.ones(N, 32, device="cuda")
            C = torch.full((128, N), float("nan"), device="cuda")
            assert be.lib.dgmr_debug_umma_shift(A.data_ptr(), Bm.data_ptr(), C.data_ptr(), N, 0, 0, st) == 0
            torch.cuda.synchronize()
            exact = float((A[:1].double() @ Bm.double().t())[0, 0])
            got = float(C[0, 0])
            lines.append(f"signs {sign:11s} s={s:2d} exact-1 {exact - 1:+.6e} got-1 {got - 1:+.6e} err {got - exact:+.3e}")
            if sign == "same" and abs(got - exact) <= 0.5 * (exact - 1):
                worst_kept = s
    print("\nACCUM\n" + "\n".join(lines))
    assert worst_kept is not None and worst_kept >= 20, f"addends survive only down to 2^-{worst_kept}"


def test_depth_to_space_round_trip_is_bit_exact_at_full_size(cuda_backend):
    """The sampler's final pixel shuffle (ref: dgmr/generators.py:173-176) and its inverse as strided permutes over the full
    18 x 16 x 256 x 256 output: pure index maps, so the round trip must reproduce the input bit for bit."""
    be = cuda_backend
    t, b, hh = 18, 16, 128
    torch.manual_seed(6)
    x = torch.randn(t * b * hh * hh * 4, device="cuda")            # [T*B, 128, 128, 4] channels-last, 4 = 2x2 sub-pixels
    y = torch.empty_like(x)                                        # [B, T, 256, 256]
    z = torch.empty_like(x)
    shape = (t, b, hh, hh, 1, 2, 2)
    s_cl = (b * hh * hh * 4, hh * hh * 4, hh * 4, 4, 4, 2, 1)      # (t, b, h, w, c=1, i, j) of the channels-last tensor
    s_img = (4 * hh * hh, t * 4 * hh * hh, 4 * hh, 2, 4 * hh * hh, 2 * hh, 1)   # the same element in [B, T, 2h+i, 2w+j]
    be.permute(x, y, shape, s_cl, s_img)
    be.permute(y, z, shape, s_img, s_cl)
    torch.cuda.synchronize()
    assert torch.equal(x, z)
    assert not torch.equal(x, y)
