"""tcgen05 / TMA implicit-GEMM convolution against the fp32 emulator and against the SIMT kernel.

Tolerance: kind::tf32 keeps 10 explicit mantissa bits of each operand (the same operand precision cuDNN uses
by default for the reference's convs); with unit-variance data and 1/sqrt(K) weights the expected error of a
K-term dot product is ~2^-11 * sqrt(2) relative to the output RMS, so 4e-3 * max|y| bounds it with margin.
"""
import pytest
import torch

from emu_backend import EmuBackend

pytestmark = pytest.mark.gpu

# N, D, H, W, Cin, Cout, kd, kh, kw, G
UMMA_SHAPES = [
    (2, 1, 16, 16, 32, 64, 1, 3, 3, 1),     # BK=32, single k-chunk per tap
    (2, 1, 16, 16, 96, 48, 1, 3, 3, 2),     # 3 k-chunks, N=48
    (4, 1, 8, 8, 64, 192, 1, 3, 3, 4),      # 8x8 images: box spans 2 images; groups
    (8, 1, 4, 4, 128, 256, 1, 3, 3, 2),     # 4x4 images: 8 images per tile
    (32, 1, 2, 2, 64, 32, 1, 3, 3, 1),      # 2x2 images
    (2, 1, 32, 32, 48, 96, 1, 3, 3, 1),     # Cin=48: 32-channel block + 16-channel tail block (TMA zero fill, 2 k-steps)
    (2, 1, 32, 32, 16, 96, 1, 3, 3, 1),     # BK=16 (64-byte swizzle)
    (2, 1, 16, 16, 72, 64, 1, 3, 3, 1),     # Cin=72: two full blocks + 8-channel tail
    (1, 1, 8, 8, 768, 768, 1, 3, 3, 1),     # one 8x8 image (latent stack): half-empty 128-pixel tile, second image out of bounds
    (3, 1, 8, 8, 64, 48, 1, 3, 3, 3),       # N=3 not a multiple of the 2-image box
    (1, 1, 4, 4, 32, 32, 1, 3, 3, 1),       # 16 pixels in a 128-pixel tile
    (2, 1, 16, 16, 24, 24, 1, 3, 3, 1),     # BK=8 (32-byte swizzle), Cout=24 -> BN=32 with OOB weight rows
    (2, 1, 16, 16, 8, 48, 1, 3, 3, 1),      # Cin=8
    (2, 1, 16, 16, 192, 384, 1, 1, 1, 1),   # 1x1, two N tiles of 192
    (1, 4, 16, 16, 32, 48, 3, 3, 3, 1),     # 3-D conv
    (2, 5, 8, 8, 48, 96, 3, 3, 3, 1),       # 3-D conv, odd depth, box spans 2 images
    (2, 1, 64, 64, 96, 96, 1, 3, 3, 1),     # bigger
    (2, 1, 8, 8, 768, 768, 1, 3, 3, 1),     # K = 6912, 3 N tiles
    (2, 1, 32, 32, 48, 4, 1, 1, 1, 2),      # Cout=4 (sampler output conv / dgrad into a 4-channel input): N tile 16, 12 OOB rows
    (2, 3, 16, 16, 48, 8, 3, 3, 3, 1),      # 3-D dgrad into the zero-padded 8-channel temporal-D input
]


@pytest.mark.parametrize("shape", UMMA_SHAPES)
@pytest.mark.parametrize("variant", ["plain", "fused", "fused-persistent", "fusedup2", "fusedup2-persistent"])
def test_conv_umma(cuda_backend, shape, variant):
    n, d, h, w, cin, cout, kd, kh, kw, g = shape
    # "persistent": force the many-tile persistent kernel (double-buffered TMEM, continuous K stream) on these small shapes
    cuda_backend.set_option("umma_persist", 2 if variant.endswith("persistent") else 0)
    variant = variant.split("-")[0]
    assert cuda_backend.conv_umma_supported(n, d, h, w, cin, cout, kd, kh, kw), "shape should be served by tcgen05 path"
    torch.manual_seed(11)
    taps = kd * kh * kw
    x = torch.randn(n, d, h, w, cin)
    wp = torch.randn(taps * cout * cin) / (taps * cin) ** 0.5
    fused = variant.startswith("fused")
    up2 = variant == "fusedup2"     # epilogue flags: half-resolution residual read at (h/2, w/2) + tf32-rounded output
    if up2 and (h % 2 or w % 2):
        pytest.skip("RES_UP2 needs even H, W")
    bias = torch.randn(cout) if fused else None
    scale = (torch.rand(g, cout) + 0.5) if fused else None
    res = (torch.randn(n, d, h // 2, w // 2, cout) if up2 else torch.randn(n, d, h, w, cout)) if fused else None
    act = (1 if fused else 0) | ((1024 | 2048) if up2 else 0)
    y_ref = torch.empty(n, d, h, w, cout)
    EmuBackend().conv_fwd(x, wp, bias, scale, res, y_ref, n, d, h, w, cin, cout, kd, kh, kw, g, act)
    dev = lambda t: None if t is None else t.cuda()
    y_umma = torch.full((n, d, h, w, cout), float("nan"), device="cuda")
    cuda_backend.conv_fwd(dev(x), dev(wp), dev(bias), dev(scale), dev(res), y_umma, n, d, h, w, cin, cout, kd, kh, kw, g, act, algo=2)
    y_simt = torch.empty(n, d, h, w, cout, device="cuda")
    cuda_backend.conv_fwd(dev(x), dev(wp), dev(bias), dev(scale), dev(res), y_simt, n, d, h, w, cin, cout, kd, kh, kw, g, act, algo=1)
    torch.cuda.synchronize()
    ref_max = y_ref.abs().max().item()
    e_simt = (y_simt.cpu() - y_ref).abs().max().item()
    e_umma = (y_umma.cpu() - y_ref).abs().max().item()
    # (ROUND_OUT: an fp32 difference of 1e-7 that straddles a tf32 rounding boundary becomes one tf32 ulp, 2^-10 of the value)
    assert e_simt <= (1.1e-3 if up2 else 2e-5) * max(ref_max, 1), f"SIMT err {e_simt:.3e}"
    assert not torch.isnan(y_umma).any(), "tcgen05 path left outputs unwritten"
    assert e_umma <= 4e-3 * max(ref_max, 1), f"tcgen05 err {e_umma:.3e} (ref max {ref_max:.3e}, simt err {e_simt:.3e})"
    if up2:   # ROUND_OUT: every output is a tf32 value (low 13 mantissa bits zero)
        for t in (y_umma, y_simt):
            assert int((t.view(torch.int32) & 0x1fff).abs().max()) == 0


# N, D, H, W, Cin, Cout, kd, kh, kw
WGRAD_SHAPES = [
    (4, 1, 32, 32, 32, 64, 1, 3, 3),      # aw=32, one ci tile
    (4, 1, 32, 32, 96, 96, 1, 3, 3),      # M tile 128 > Cout=96 (OOB channel block), BN=96
    (8, 1, 16, 16, 48, 96, 1, 3, 3),      # Cin=48: second 32-channel block half out of bounds; pixel box 16x2
    (16, 1, 8, 8, 64, 192, 1, 3, 3),      # 8x8 images: pixel box 8x4, two co tiles
    (64, 1, 4, 4, 128, 256, 1, 3, 3),     # 4x4 images: box spans 2 images
    (4, 1, 32, 32, 24, 24, 1, 3, 3),      # 24 channels inside one 32-channel block
    (2, 1, 64, 64, 192, 384, 1, 1, 1),    # 1x1
    (2, 6, 16, 16, 48, 96, 3, 3, 3),      # 3-D
    (4, 1, 16, 16, 768, 768, 1, 3, 3),    # 3 ci tiles x 6 co tiles
    (4, 1, 32, 32, 8, 48, 1, 3, 3),       # zero-padded 4->8 channel input (first DBlocks)
    (2, 4, 16, 16, 8, 48, 3, 3, 3),       # same, 3-D
    (4, 1, 32, 32, 48, 4, 1, 1, 1),       # Cout=4
]


@pytest.mark.parametrize("shape", WGRAD_SHAPES)
def test_wgrad_umma(cuda_backend, shape):
    """tcgen05 wgrad (MN-major operands read straight from the channels-last tensors, split-K) vs fp32 emulator and SIMT."""
    n, d, h, w, cin, cout, kd, kh, kw = shape
    assert cuda_backend.wgrad_umma_supported(n, d, h, w, cin, cout, kd, kh, kw)
    torch.manual_seed(12)
    taps = kd * kh * kw
    x, dz = torch.randn(n, d, h, w, cin), torch.randn(n, d, h, w, cout)
    ref = torch.empty(taps * cout * cin)
    EmuBackend().conv_wgrad(x, dz, ref, n, d, h, w, cin, cout, kd, kh, kw)
    g_umma = torch.full((taps * cout * cin,), float("nan"), device="cuda")
    g_simt = torch.empty(taps * cout * cin, device="cuda")
    cuda_backend.conv_wgrad(x.cuda(), dz.cuda(), g_umma, n, d, h, w, cin, cout, kd, kh, kw, algo=2)
    cuda_backend.conv_wgrad(x.cuda(), dz.cuda(), g_simt, n, d, h, w, cin, cout, kd, kh, kw, algo=1)
    torch.cuda.synchronize()
    ref_max = ref.abs().max().item()
    e_simt = (g_simt.cpu() - ref).abs().max().item()
    e_umma = (g_umma.cpu() - ref).abs().max().item()
    assert e_simt <= 1e-4 * ref_max, f"SIMT err {e_simt:.3e}"
    assert not torch.isnan(g_umma).any()
    assert e_umma <= 4e-3 * ref_max, f"tcgen05 wgrad err {e_umma:.3e} (ref max {ref_max:.3e})"


@pytest.mark.parametrize("shape", [(4, 1, 8, 8, 384, 384, 1, 3, 3, 1), (2, 1, 16, 16, 192, 192, 1, 3, 3, 2), (2, 3, 8, 8, 64, 96, 3, 3, 3, 1)])
def test_conv_umma_accumulate_mode(cuda_backend, shape):
    """DGMR_FLAG_ACCUMULATE: y += conv(x)*scale with the K loop split over filter taps across CTAs (ConvGRU step convs)."""
    n, d, h, w, cin, cout, kd, kh, kw, g = shape
    torch.manual_seed(13)
    taps = kd * kh * kw
    x = torch.randn(n, d, h, w, cin)
    wp = torch.randn(taps * cout * cin) / (taps * cin) ** 0.5
    scale = torch.rand(g, cout) + 0.5
    y0 = torch.randn(n, d, h, w, cout)
    y_ref = y0.clone()
    EmuBackend().conv_fwd(x, wp, None, scale, None, y_ref, n, d, h, w, cin, cout, kd, kh, kw, g, 512)
    y = y0.cuda()
    cuda_backend.conv_fwd(x.cuda(), wp.cuda(), None, scale.cuda(), None, y, n, d, h, w, cin, cout, kd, kh, kw, g, 512)
    torch.cuda.synchronize()
    assert (y.cpu() - y_ref).abs().max().item() <= 4e-3 * y_ref.abs().max().item()


# N, D, H, W, Cin, Cout, kd, G
PATCH_SHAPES = [
    (2, 1, 32, 32, 32, 64, 1, 1),      # P=34
    (9, 1, 64, 64, 96, 96, 1, 3),      # P=66, groups, odd batch, two sub-tiles per item
    (4, 1, 128, 128, 96, 48, 1, 2),    # P=130, two sub-tiles per item, double-buffered accumulators
    (2, 1, 32, 32, 384, 384, 1, 1),    # two N tiles of 192, single accumulator buffer
    (1, 1, 64, 32, 64, 20, 1, 1),      # H != W, Cout not a multiple of 16
    (2, 5, 32, 32, 96, 96, 3, 1),      # 3-D: 3 depth taps, odd depth
    (1, 1, 40, 48, 32, 32, 1, 1),      # W not a power of two
    (4, 1, 64, 64, 48, 48, 1, 1),      # Cin = 48: 32-channel chunk + 16-channel tail chunk
    (4, 1, 64, 64, 16, 48, 1, 1),      # Cin = 16: 16-channel chunks (64-byte rows)
    (2, 6, 32, 32, 48, 96, 3, 1),      # same, 3-D
    (16, 1, 16, 16, 192, 192, 1, 1),   # small images (ConvGRU step): Cout split over more CTAs
    (16, 1, 32, 32, 96, 96, 1, 1),
]


@pytest.mark.parametrize("shape", PATCH_SHAPES)
@pytest.mark.parametrize("variant", ["plain", "fused", "fusedup2"])
def test_conv_umma_patch(cuda_backend, shape, variant):
    """Halo-patch persistent kernel (one activation patch feeds all 9 taps) vs fp32 emulator and the plain tcgen05 kernel."""
    n, d, h, w, cin, cout, kd, g = shape
    torch.manual_seed(14)
    taps = kd * 9
    x = torch.randn(n, d, h, w, cin)
    wp = torch.randn(taps * cout * cin) / (taps * cin) ** 0.5
    fused = variant.startswith("fused")
    up2 = variant == "fusedup2"
    bias = torch.randn(cout) if fused else None
    scale = (torch.rand(g, cout) + 0.5) if fused else None
    res = (torch.randn(n, d, h // 2, w // 2, cout) if up2 else torch.randn(n, d, h, w, cout)) if fused else None
    act = (1 if fused else 0) | ((1024 | 2048) if up2 else 0)
    y_ref = torch.empty(n, d, h, w, cout)
    EmuBackend().conv_fwd(x, wp, bias, scale, res, y_ref, n, d, h, w, cin, cout, kd, 3, 3, g, act)
    dev = lambda t: None if t is None else t.cuda()
    y = torch.full((n, d, h, w, cout), float("nan"), device="cuda")
    cuda_backend.conv_fwd(dev(x), dev(wp), dev(bias), dev(scale), dev(res), y, n, d, h, w, cin, cout, kd, 3, 3, g, act, algo=3)
    torch.cuda.synchronize()
    assert not torch.isnan(y).any(), "halo-patch kernel left outputs unwritten"
    e = (y.cpu() - y_ref).abs().max().item()
    assert e <= 4e-3 * max(y_ref.abs().max().item(), 1), f"halo-patch kernel err {e:.3e}"


# N, D, H, W, Cin, Cout, kd, G
KWSTACK_SHAPES = [
    (3, 1, 64, 64, 48, 48, 1, 1),      # W = 64: two image rows per tile, warp-boundary exchange; 32-channel chunk + 16-channel tail
    (4, 1, 128, 128, 96, 48, 1, 2),    # W = 128: one row per tile, three warp boundaries; groups
    (2, 5, 32, 32, 48, 48, 3, 1),      # W = 32: four rows per tile, no exchange; 3-D, odd depth
    (2, 6, 64, 64, 48, 48, 3, 1),      # the temporal discriminator's 48 -> 48 (small)
    (2, 1, 64, 64, 16, 48, 1, 1),      # 16-channel chunks (64-byte rows)
    (2, 3, 32, 32, 96, 16, 3, 1),      # Cout = 16 (N = 48)
    (5, 1, 64, 64, 96, 80, 1, 5),      # Cout = 80 (N = 240), odd tile count
    (1, 1, 2, 64, 32, 32, 1, 1),       # a single tile
]


@pytest.mark.parametrize("shape", KWSTACK_SHAPES)
@pytest.mark.parametrize("variant", ["plain", "fused", "fusedup2"])
@pytest.mark.parametrize("pair", [0, 1])
def test_conv_umma_kwstack(cuda_backend, shape, variant, pair):
    """Column-stacked kernel (three kw taps along the MMA's N, shift-add in the epilogue) vs fp32 emulator; same packed weights as every other kernel.
    pair = 1: CTA pairs (cta_group::2, half of each weight tile per CTA) forced on these small shapes, odd tile counts included."""
    n, d, h, w, cin, cout, kd, g = shape
    cuda_backend.set_option("kwstack_pair", pair)
    torch.manual_seed(24)
    taps = kd * 9
    x = torch.randn(n, d, h, w, cin)
    wp = torch.randn(taps * cout * cin) / (taps * cin) ** 0.5
    fused = variant.startswith("fused")
    up2 = variant == "fusedup2"
    bias = torch.randn(cout) if fused else None
    scale = (torch.rand(g, cout) + 0.5) if fused else None
    res = (torch.randn(n, d, h // 2, w // 2, cout) if up2 else torch.randn(n, d, h, w, cout)) if fused else None
    act = (1 if fused else 0) | ((1024 | 2048) if up2 else 0)
    y_ref = torch.empty(n, d, h, w, cout)
    EmuBackend().conv_fwd(x, wp, bias, scale, res, y_ref, n, d, h, w, cin, cout, kd, 3, 3, g, act)
    dev = lambda t: None if t is None else t.cuda()
    y = torch.full((n, d, h, w, cout), float("nan"), device="cuda")
    cuda_backend.conv_fwd(dev(x), dev(wp), dev(bias), dev(scale), dev(res), y, n, d, h, w, cin, cout, kd, 3, 3, g, act, algo=4)
    torch.cuda.synchronize()
    assert not torch.isnan(y).any(), "column-stacked kernel left outputs unwritten"
    e = (y.cpu() - y_ref).abs().max().item()
    assert e <= 4e-3 * max(y_ref.abs().max().item(), 1), f"column-stacked kernel err {e:.3e}"
    # the same call through AUTO with the option forced must take this kernel too (bit-identical result)
    cuda_backend.set_option("kwstack", 1)
    y2 = torch.empty_like(y)
    cuda_backend.conv_fwd(dev(x), dev(wp), dev(bias), dev(scale), dev(res), y2, n, d, h, w, cin, cout, kd, 3, 3, g, act)
    torch.cuda.synchronize()
    if cout < 64 and cin >= 32:     # (AUTO leaves 16-channel inputs to the plain kernel: measured slower here)
        assert torch.equal(y, y2)


# N, D, H, W, Cin, Cout, kd, G
PAIRCONV_SHAPES = [
    (6, 1, 16, 16, 192, 384, 1, 3),    # W = 16: eight image rows per tile; two Cout tiles of 192
    (4, 1, 16, 16, 768, 768, 1, 2),    # the weight-heavy sampler layer (small batch): three Cout tiles of 256
    (3, 1, 32, 32, 96, 80, 1, 1),      # odd tile count (the pair's last tile past the end), Cout = 80
    (2, 3, 16, 16, 64, 32, 3, 1),      # 3-D
    (2, 1, 64, 64, 48, 96, 1, 2),      # channel tail (32 + 16), W = 64
]


@pytest.mark.parametrize("shape", PAIRCONV_SHAPES)
@pytest.mark.parametrize("variant", ["plain", "fused", "fusedup2"])
@pytest.mark.parametrize("pair", [0, 1])
def test_conv_umma_pairconv(cuda_backend, shape, variant, pair):
    """Persistent whole-row-tile kernel with ordinary taps (conv_kwstack.cu, STACK = false), single CTAs and CTA pairs sharing each weight tile."""
    n, d, h, w, cin, cout, kd, g = shape
    cuda_backend.set_option("kwstack_pair", pair)
    torch.manual_seed(25)
    taps = kd * 9
    x = torch.randn(n, d, h, w, cin)
    wp = torch.randn(taps * cout * cin) / (taps * cin) ** 0.5
    fused = variant.startswith("fused")
    up2 = variant == "fusedup2"
    bias = torch.randn(cout) if fused else None
    scale = (torch.rand(g, cout) + 0.5) if fused else None
    res = (torch.randn(n, d, h // 2, w // 2, cout) if up2 else torch.randn(n, d, h, w, cout)) if fused else None
    act = (1 if fused else 0) | ((1024 | 2048) if up2 else 0)
    y_ref = torch.empty(n, d, h, w, cout)
    EmuBackend().conv_fwd(x, wp, bias, scale, res, y_ref, n, d, h, w, cin, cout, kd, 3, 3, g, act)
    dev = lambda t: None if t is None else t.cuda()
    y = torch.full((n, d, h, w, cout), float("nan"), device="cuda")
    cuda_backend.conv_fwd(dev(x), dev(wp), dev(bias), dev(scale), dev(res), y, n, d, h, w, cin, cout, kd, 3, 3, g, act, algo=5)
    torch.cuda.synchronize()
    assert not torch.isnan(y).any(), "pair-persistent kernel left outputs unwritten"
    e = (y.cpu() - y_ref).abs().max().item()
    assert e <= 4e-3 * max(y_ref.abs().max().item(), 1), f"pair-persistent kernel err {e:.3e}"


# N, D, H, W, Cin, Cout, kd, kh
WGRAD_ROW_SHAPES = [
    (4, 1, 32, 32, 32, 64, 1, 3),
    (2, 1, 64, 64, 96, 96, 1, 3),      # M tile 128 > Cout, BN = 96 (3 patch blocks)
    (2, 1, 32, 64, 192, 48, 1, 3),     # two ci tiles of 96, W > H
    (2, 1, 128, 128, 48, 24, 1, 3),    # channel tails inside a 32-channel block
    (2, 4, 32, 32, 48, 96, 3, 3),      # 3-D
    (1, 1, 32, 32, 768, 384, 1, 3),    # 5 ci tiles of 160, 3 co tiles
    (2, 1, 32, 32, 8, 48, 1, 3),       # zero-padded 8-channel input
    (4, 1, 16, 16, 96, 64, 1, 3),      # W = 16: a K block is two whole image rows (36-row patch)
    (3, 1, 16, 16, 384, 192, 1, 3),    # same, three ci tiles
    (6, 1, 8, 8, 64, 96, 1, 3),        # W = 8: four image rows per K block (40-row patch)
    (2, 3, 8, 16, 48, 48, 3, 3),       # W = 16, 3-D, H = 8
]


@pytest.mark.parametrize("shape", WGRAD_ROW_SHAPES)
def test_wgrad_umma_row(cuda_backend, shape):
    """Row variant: the three kw taps of a filter row share one dz tile and one 34-pixel x patch (shifted B descriptors)."""
    n, d, h, w, cin, cout, kd, kh = shape
    torch.manual_seed(15)
    taps = kd * kh * 3
    x, dz = torch.randn(n, d, h, w, cin), torch.randn(n, d, h, w, cout)
    ref = torch.empty(taps * cout * cin)
    EmuBackend().conv_wgrad(x, dz, ref, n, d, h, w, cin, cout, kd, kh, 3)
    g = torch.full((taps * cout * cin,), float("nan"), device="cuda")
    cuda_backend.conv_wgrad(x.cuda(), dz.cuda(), g, n, d, h, w, cin, cout, kd, kh, 3, algo=3)
    torch.cuda.synchronize()
    assert not torch.isnan(g).any()
    e = (g.cpu() - ref).abs().max().item()
    assert e <= 4e-3 * ref.abs().max().item(), f"row wgrad err {e:.3e} (ref max {ref.abs().max().item():.3e})"


# N, H, W (low resolution), Cin, Cout, G
UPCONV_SHAPES = [
    (4, 16, 16, 64, 64, 2),       # smallest supported image (H*W = 256), CTA pairs, groups
    (3, 32, 32, 96, 96, 3),       # odd batch: the pair's second image is past the end for the last item
    (2, 64, 64, 96, 96, 1),       # P = 66
    (2, 32, 32, 384, 192, 1),     # two Cout tiles of 96, 12 channel chunks
    (1, 32, 32, 64, 160, 1),      # single image: no CTA pairs; Cout = 160 -> two tiles of 80
    (2, 16, 16, 768, 256, 1),     # widest K
    (2, 128, 128, 32, 64, 2),     # P = 130
]


@pytest.mark.parametrize("shape", UPCONV_SHAPES)
def test_upconv_subpixel_kernels(cuda_backend, shape):
    """csrc/conv_subpix.cu: forward (plain and fused epilogue), dgrad (four strided phase views of dz) and the 16-tile weight gradient of the
    sub-pixel up-convolution against the host emulator's restatement of the phase formula; pre-summed packs bit-exact."""
    be = cuda_backend
    n, h, w, cin, cout, g = shape
    assert be.upconv_supported(n, h, w, cin, cout)
    emu = EmuBackend()
    torch.manual_seed(21)
    x = torch.randn(n, 1, h, w, cin)
    wt = torch.randn(cout, cin, 3, 3) / (9 * cin) ** 0.5
    # packs: pure sums + the same rounding instruction -> equal up to the order of the <= 4-term fp32 sums
    for mode in (0, 1, 0 | 256, 1 | 256):
        ref = torch.empty(16 * cout * cin)
        emu.pack_weight_subpix(wt, ref, cout, cin, 0, cin, mode)
        got = torch.empty(16 * cout * cin, device="cuda")
        be.pack_weight_subpix(wt.cuda(), got, cout, cin, 0, cin, mode)
        assert (got.cpu() - ref).abs().max().item() <= (1e-6 if not mode & 256 else 1e-3) * ref.abs().max().item(), mode
    wsp = torch.empty(16 * cout * cin)
    emu.pack_weight_subpix(wt, wsp, cout, cin, 0, cin, 0)
    wspt = torch.empty(16 * cout * cin)
    emu.pack_weight_subpix(wt, wspt, cout, cin, 0, cin, 1)
    # forward: the halo-patch form (subpix_rows = 0) and, on narrow images, the whole-row CTA-pair form (1; single CTAs with kwstack_pair = 0)
    for fused, rows, pair in [(f, r, pr) for f in (False, True) for (r, pr) in ((0, -1), (1, 1), (1, 0))]:
        if rows and not (w <= 32 and cout % 16 == 0):
            continue
        be.set_option("subpix_rows", rows); be.set_option("kwstack_pair", pair)
        bias = torch.randn(cout) if fused else None
        scale = (torch.rand(g, cout) + 0.5) if fused else None
        res = torch.randn(n, 1, 2 * h, 2 * w, cout) if fused else None
        act = (1 | 1024) if fused else 0
        y_ref = torch.empty(n, 1, 2 * h, 2 * w, cout)
        emu.upconv_fwd(x, wsp, bias, scale, res, y_ref, n, h, w, cin, cout, g, act)
        dev = lambda t: None if t is None else t.cuda()
        y = torch.full((n, 1, 2 * h, 2 * w, cout), float("nan"), device="cuda")
        be.upconv_fwd(x.cuda(), wsp.cuda(), dev(bias), dev(scale), dev(res), y, n, h, w, cin, cout, g, act)
        torch.cuda.synchronize()
        assert not torch.isnan(y).any(), "sub-pixel forward left outputs unwritten"
        e = (y.cpu() - y_ref).abs().max().item()
        assert e <= 4e-3 * max(y_ref.abs().max().item(), 1), f"sub-pixel forward err {e:.3e} (fused={fused}, rows={rows}, pair={pair})"
    be.set_option("subpix_rows", -1); be.set_option("kwstack_pair", -1)
    dz = torch.randn(n, 1, 2 * h, 2 * w, cout)
    dx_ref = torch.empty(n, 1, h, w, cin)
    emu.upconv_dgrad(dz, wspt, dx_ref, n, h, w, cin, cout)
    dx = torch.full((n, 1, h, w, cin), float("nan"), device="cuda")
    be.upconv_dgrad(dz.cuda(), wspt.cuda(), dx, n, h, w, cin, cout)
    dw_ref = torch.empty(16 * cout * cin)
    emu.upconv_wgrad(x, dz, dw_ref, n, h, w, cin, cout)
    torch.cuda.synchronize()
    assert not torch.isnan(dx).any()
    e = (dx.cpu() - dx_ref).abs().max().item()
    assert e <= 4e-3 * dx_ref.abs().max().item(), f"sub-pixel dgrad err {e:.3e}"
    for row in ((0, 1) if (w % 32 == 0 or w == 16) else (0,)):     # tap-wise kernel / row kernel (two column taps per CTA; W = 16: two image rows per K block)
        be.set_option("subpix_wgrad_row", row)
        dw = torch.full((16 * cout * cin,), float("nan"), device="cuda")
        be.upconv_wgrad(x.cuda(), dz.cuda(), dw, n, h, w, cin, cout)
        torch.cuda.synchronize()
        assert not torch.isnan(dw).any()
        e = (dw.cpu() - dw_ref).abs().max().item()
        assert e <= 4e-3 * dw_ref.abs().max().item(), f"sub-pixel wgrad err {e:.3e} (row kernel: {row})"
    # folding the 16 tile gradients back onto the 3x3 taps
    gw_ref = torch.zeros(cout, cin, 3, 3)
    emu.unpack_wgrad_subpix(dw_ref, gw_ref, cout, cin, 0, cin, False)
    gw = torch.empty(cout, cin, 3, 3, device="cuda")
    be.unpack_wgrad_subpix(dw_ref.cuda(), gw, cout, cin, 0, cin, False)
    assert (gw.cpu() - gw_ref).abs().max().item() <= 1e-5 * gw_ref.abs().max().item()


def test_upsample_gblock_subpixel_on_gpu(cuda_backend):
    """The wide UpsampleGBlock case takes the sub-pixel kernels in AUTO mode (S = 16: H*W = 256): block output and gradients vs the oracle."""
    from block_cases import block_cases, run_block_case
    from skillful_nowcasting_b200 import ops

    case = [c for c in block_cases(True) if c[0] == "upg"][0]
    n0 = cuda_backend.launches
    assert ops.upconv_available(2, 16, 16, 64, 64)
    run_block_case(case, True, "cuda", 1e-3, 1.5e-1, tol_buf=1e-3, tol_l2=5e-2)
    run_block_case(case, False, "cuda", 1e-3, 1.5e-1, tol_buf=1e-3, tol_l2=5e-2)
