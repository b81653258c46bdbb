"""DRAFT (not collected by pytest: no test_ prefix) -- to be finished with GPU time in the next round.
Only run so far: shape 0, first identity: <conv(x),dy> = 4.50643183e8 vs <x,dgrad(dy)> = 4.50596807e8, i.e. they agree to 1.03e-4
relative -- an order of magnitude looser than the fp32-accumulation estimate below, which needs explaining (tensor-core internal
accumulation width?) before a tolerance is fixed.  Run with:  python -m pytest tests/fullsize_props_draft.py -m gpu

Size-independent properties at BASELINE.json's full layer sizes (B = 16, T = 18 -> 288 images; temporal D 32 x 22 frames).

The oracle cannot finish these sizes in seconds, so the three convolution kernels (forward, dgrad = forward with the
flipped-transposed pack, wgrad) are checked against EACH OTHER through the adjoint identities of a linear map:

    <conv(x, w), dy>  ==  <x, dgrad(dy, w)>  ==  <w, wgrad(x, dy)>

With every operand rounded to TF32 beforehand the tensor cores multiply exactly, so the three numbers differ only by fp32
accumulation order (expected ~1e-6 relative; a dropped tap, a wrong border, an unrounded operand or a mis-scaled split-K
partial shows up at 1e-3 or worse).  dy is chosen as tf32(y), which makes all three inner products ~ ||y||^2 (large and
positive), so relative comparison is meaningful.  Tolerance: 1e-4 relative.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

# N, D, H, W, Cin, Cout, kd, k  -> which kernels serve it at this size
FULL_SHAPES = [
    (288, 1, 128, 128, 96, 96, 1, 3),   # CTA-pair halo-patch kernel (fwd, dgrad), row wgrad
    (288, 1, 128, 128, 96, 48, 1, 3),   # plain kernel fwd (Cout 48), patch dgrad (48 -> 96 tail chunk), row wgrad
    (32, 22, 64, 64, 48, 48, 3, 3),     # temporal D: 3-D, channel-tail blocks, row wgrad with 9 (kd,kh) rows
    (128, 1, 128, 128, 8, 48, 1, 3),    # first D block: 8-channel (zero-padded) input, persistent short-K kernel
    (288, 1, 64, 64, 48, 96, 1, 1),     # 1x1 conv: persistent plain kernel, tap-wise wgrad
    (288, 1, 16, 16, 768, 768, 1, 3),   # widest layer: 3 N tiles, patch n/a (H*W < 1024) -> plain kernel, tap-wise wgrad
]


def _dot(a, b):
    return float((a.double().flatten() * b.double().flatten()).sum())


@pytest.mark.parametrize("shape", FULL_SHAPES)
def test_conv_adjoint_identities_at_full_size(cuda_backend, shape):
    be = cuda_backend
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
    y = torch.full((n, d, h, w, cout), float("nan"), device="cuda")
    be.conv_fwd(x, wp, None, None, None, y, n, d, h, w, cin, cout, kd, k, k, 1, 0)
    dy = torch.empty_like(y)
    be.round_tf32(y, dy)
    dx = torch.full((n, d, h, w, cin), float("nan"), device="cuda")
    be.conv_fwd(dy, wpt, None, None, None, dx, n, d, h, w, cout, cin, kd, k, k, 1, 0)
    dwp = torch.empty(taps * cout * cin, device="cuda")
    be.conv_wgrad(x, dy, dwp, n, d, h, w, cin, cout, kd, k, k)
    torch.cuda.synchronize()
    assert not torch.isnan(y).any() and not torch.isnan(dx).any(), "kernel left outputs unwritten"
    a, b, c = _dot(y, dy), _dot(x, dx), _dot(wp, dwp)
    assert a > 0
    assert abs(a - b) <= 1e-4 * a, f"<conv(x),dy> = {a:.9e} vs <x,dgrad(dy)> = {b:.9e}"
    assert abs(a - c) <= 1e-4 * a, f"<conv(x),dy> = {a:.9e} vs <w,wgrad(x,dy)> = {c:.9e}"


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
