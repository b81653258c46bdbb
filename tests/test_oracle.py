"""The CPU oracle against (a) the committed golden fixture produced by the unmodified reference
(tests/golden/make_golden.py) and (b) the live reference modules when /root/reference is present."""
import pytest
import torch

from oracle import dgmr_oracle as O
from parity_util import C1, GOLDEN, build_gan, c1_inputs, compare_grads, oracle_gan_forward, rel_err, state_checksum


@pytest.fixture(scope="module")
def c1():
    gen, disc = build_gan(C1, seed=0, gamma=0.5)
    return ({k: v.clone() for k, v in gen.state_dict().items()}, {k: v.clone() for k, v in disc.state_dict().items()})


def test_seeded_construction_matches_fixture_checksum(c1):
    fix = torch.load(GOLDEN)
    assert state_checksum(c1[0]) == pytest.approx(fix["g_checksum"], rel=1e-8)
    assert state_checksum(c1[1]) == pytest.approx(fix["d_checksum"], rel=1e-8)


@pytest.mark.parametrize("mode", ["eval", "train"])
def test_oracle_reproduces_reference_fixture(c1, mode):
    fix = torch.load(GOLDEN)
    x, y = c1_inputs()
    training = mode == "train"
    res = oracle_gan_forward(c1[0], c1[1], x, y, C1, training, seed=fix["seed"])
    # eval is deterministic op-for-op; train mode amplifies fp32 summation-order noise (thread count) ~1e-4
    tol = 2e-4 if training else 1e-6
    assert rel_err(res["out"], fix[mode]["out"]) < tol
    assert rel_err(res["scores"], fix[mode]["scores"]) < tol * 5
    for k in ("d_loss", "grid", "g_loss"):
        assert rel_err(res[k], fix[mode][k]) < tol * 5
    if training:
        for sd, after in ((res["g_state"], fix[mode]["g_state_after"]), (res["d_state"], fix[mode]["d_state_after"])):
            for k, v in after.items():
                if "num_batches" in k:
                    assert int(sd[k]) == int(v)
                else:
                    assert rel_err(sd[k], v) < 1e-4, k
        compare_grads(res["d_grads"], fix[mode]["d_grads"], 2e-3, 2e-2, zero_floor=1e-6)
        compare_grads(res["g_grads"], fix[mode]["g_grads"], 5e-2, 2e-1, zero_floor=1e-5)


def test_rng_draws_match_reference_record():
    fix = torch.load(GOLDEN)
    torch.manual_seed(fix["seed"])
    z = torch.normal(torch.zeros(8, 4, 4, 1), torch.ones(8, 4, 4, 1))
    idx = torch.randint(low=0, high=8, size=(8,))
    assert torch.equal(z, fix["z"]) and torch.equal(idx, fix["idxs"])


def test_pixel_shuffle_roundtrip_is_exact():
    x = torch.randn(2, 3, 8, 6, 10)
    assert torch.equal(O.pixel_shuffle(O.pixel_unshuffle(x)), x)
    assert torch.equal(O.pixel_unshuffle(x[0]), torch.nn.PixelUnshuffle(2)(x[0]))
    assert torch.equal(O.pixel_shuffle(O.pixel_unshuffle(x[0])), torch.nn.PixelShuffle(2)(torch.nn.PixelUnshuffle(2)(x[0])))


@pytest.mark.reference
def test_oracle_blocks_against_live_reference():
    import make_golden as mg

    mg.import_reference()
    from dgmr.common import DBlock, GBlock, LBlock, UpsampleGBlock
    from dgmr.layers import AttentionLayer, ConvGRU

    torch.manual_seed(3)
    cases = [
        (GBlock(16, 16), lambda st, x, tr: O.g_block(st, "m", x, tr), torch.rand(2, 16, 8, 8)),
        (UpsampleGBlock(16, 8), lambda st, x, tr: O.upsample_g_block(st, "m", x, tr), torch.rand(2, 16, 8, 8)),
        (DBlock(8, 16), lambda st, x, tr: O.d_block(st, "m", x, tr), torch.rand(2, 8, 8, 8)),
        (DBlock(4, 8, conv_type="3d", first_relu=False), lambda st, x, tr: O.d_block(st, "m", x, tr, first_relu=False), torch.rand(2, 4, 6, 8, 8)),
        (DBlock(8, 8, keep_same_output=True), lambda st, x, tr: O.d_block(st, "m", x, tr, keep_same_output=True), torch.rand(2, 8, 4, 4)),
        (LBlock(8, 24), lambda st, x, tr: O.l_block(st, "m", x), torch.rand(1, 8, 4, 4)),
    ]
    for mod, fn, x in cases:
        for tr in (True, False):
            mod.train(tr)
            st = {"m." + k: v.clone() for k, v in mod.state_dict().items()}
            ref = mod(x)
            got = fn(st, x, tr)
            assert rel_err(got, ref) < 1e-5, type(mod).__name__
            for k, v in mod.state_dict().items():
                assert rel_err(st["m." + k], v) < 1e-5 or v.numel() == 0, k
    att = AttentionLayer(48, 48)
    with torch.no_grad():
        att.gamma.fill_(0.7)
    x = torch.randn(1, 48, 4, 4)
    assert rel_err(O.attention({"a." + k: v for k, v in att.state_dict().items()}, "a", x), att(x)) < 1e-5
    gru = ConvGRU(24 + 8, 8)
    xs, h = [torch.rand(2, 24, 8, 8) for _ in range(3)], torch.rand(2, 8, 8, 8)
    st = {"g." + k: v.clone() for k, v in gru.state_dict().items()}
    assert rel_err(O.conv_gru(st, "g", xs, h, True), gru(xs, h)) < 1e-5


@pytest.mark.reference
def test_seeded_construction_equals_reference_bitwise(c1):
    import make_golden as mg

    rg, rd = mg.build_reference_gan(C1, seed=0)
    with torch.no_grad():
        rg.latent_stack.att_block.gamma.fill_(0.5)
    for ref_sd, mine in ((rg.state_dict(), c1[0]), (rd.state_dict(), c1[1])):
        assert list(ref_sd.keys()) == list(mine.keys())
        for k in ref_sd:
            assert torch.equal(ref_sd[k], mine[k]), k


def test_tf32_operand_rounding_alone_moves_spatial_scores():
    """Documents the conditioning of the discriminator scores ('D scores exceed the stated tolerance').  Rounding the
    convolution operands of the fp32 CPU ORACLE to TF32 (round-to-nearest, fp32 accumulate -- what cuDNN does by default for the
    reference's convolutions and what the 1xTF32 tcgen05 kernels do) moves the eval-mode SPATIAL score by ~1e-2 relative: the score is
    a cancelling sum (8 frames x 768 normalised features, result 4e-3).  The temporal score moves by < 1e-3.  No GPU involved."""
    import torch.nn.functional as F

    import skillful_nowcasting_b200 as B
    from oracle import dgmr_oracle as O

    def rna(x):
        i = x.contiguous().view(torch.int32)
        return ((i + 0x1000) & ~0x1fff).view(torch.float32).view_as(x)

    res = {}
    orig2, orig3 = F.conv2d, F.conv3d
    for which in ("spatial", "temporal"):
        torch.manual_seed(3)
        mod = B.SpatialDiscriminator(input_channels=1) if which == "spatial" else B.TemporalDiscriminator(input_channels=1)
        pfx = which + "_discriminator"
        st = O.clone_state({pfx + "." + k: v for k, v in mod.state_dict().items()})
        x = torch.rand(4, 8, 1, 128, 128)
        fn = O.spatial_discriminator if which == "spatial" else O.temporal_discriminator
        torch.manual_seed(4)
        ref = fn(st, pfx, x, False)
        try:
            F.conv2d = lambda a, w, b=None, *r, **k: orig2(rna(a), rna(w), b, *r, **k)
            F.conv3d = lambda a, w, b=None, *r, **k: orig3(rna(a), rna(w), b, *r, **k)
            torch.manual_seed(4)
            got = fn(st, pfx, x, False)
        finally:
            F.conv2d, F.conv3d = orig2, orig3
        res[which] = float((got - ref).abs().max() / ref.abs().max())
    assert 3e-3 < res["spatial"] < 4e-2, res       # measured 1.27e-2: inherent to 2^-11 operands, not to the kernels
    assert res["temporal"] < 1e-3, res             # measured 4.9e-4
