"""Parity of the B200 CUDA path (through the C ABI) against the CPU oracle and the committed golden fixture
(tests/golden/c1_gan.pt: outputs of the unmodified reference on BASELINE config C1).

Tolerances (rel = max|a-b| / max|b|):
  * eval-mode forward, SIMT fp32 convs: 2e-5; tensor-core (1xTF32 operands, like cuDNN's default for the reference): 1e-3.
  * train-mode end-to-end: the freshly initialised net with batch-stat BatchNorm amplifies rounding chaotically
    (the reference itself moves by 1e-4 rel when the CPU thread count changes; SURVEY.md section 7 measures 1.3e-2 for
    TF32 operands), so e2e train tolerances are 2e-3 (fp32 SIMT) / 5e-2 (1xTF32) and the tight checks live in the per-block tests.
"""
import pytest
import torch

from parity_util import (C1, GOLDEN, build_gan, c1_inputs, compare_grads, global_grad_error, module_gan_forward,
                         oracle_gan_forward, rel_err, state_checksum)

pytestmark = pytest.mark.gpu

ALGOS = {"simt": 1, "auto": 0}


@pytest.fixture(scope="module")
def c1_state():
    gen, disc = build_gan(C1, seed=0, gamma=0.5)
    g0 = {k: v.clone() for k, v in gen.state_dict().items()}
    d0 = {k: v.clone() for k, v in disc.state_dict().items()}
    return gen, disc, g0, d0


@pytest.mark.parametrize("algo", ["simt", "auto"])
@pytest.mark.parametrize("mode", ["eval", "train"])
def test_c1_gan_against_fixture_and_oracle(cuda_backend, c1_state, algo, mode):
    from skillful_nowcasting_b200 import ops

    gen, disc, g0, d0 = c1_state
    fix = torch.load(GOLDEN)
    training = mode == "train"
    x, y = c1_inputs()
    # seeded construction reproduces the weights the fixture was generated with (up to thread-count rounding of the
    # 16 warm-up power iterations)
    close = lambda a, b: all(abs(x - y) <= 1e-7 * abs(y) for x, y in zip(a, b))
    have_fixture = close(state_checksum(g0), fix["g_checksum"]) and close(state_checksum(d0), fix["d_checksum"])
    ref = oracle_gan_forward(g0, d0, x, y, C1, training, seed=2)
    if have_fixture:  # the oracle itself must reproduce the reference's recorded outputs
        assert rel_err(ref["out"], fix[mode]["out"]) < (1e-3 if training else 1e-5)
        assert rel_err(ref["scores"], fix[mode]["scores"]) < (1e-3 if training else 1e-5)
    gen.load_state_dict(g0); disc.load_state_dict(d0)
    gen.cuda(); disc.cuda()
    ops.clear_pack_cache()
    ops.config.conv_algo = ops.config.wgrad_algo = ALGOS[algo]
    try:
        got = module_gan_forward(gen, disc, x, y, training, seed=2, device="cuda")
    finally:
        ops.config.conv_algo = ops.config.wgrad_algo = 0
    tc = algo == "auto"
    tol_out = (5e-2 if tc else 2e-3) if training else (1e-3 if tc else 2e-5)
    for name, r in (("oracle", ref), ("fixture", fix[mode] if have_fixture else None)):
        if r is None:
            continue
        assert rel_err(got["out"], r["out"]) < tol_out, (name, rel_err(got["out"], r["out"]))
        assert rel_err(got["scores"], r["scores"]) < tol_out * 5, (name, rel_err(got["scores"], r["scores"]))
        for k in ("d_loss", "grid", "g_loss"):
            assert rel_err(got[k], r[k]) < tol_out, (name, k, float(got[k]), float(r[k]))
    if training:
        # buffers the reference mutates in training mode: u/v depend on weights only -> tight; BN running stats on data
        gsd, dsd = gen.state_dict(), disc.state_dict()
        for sd, rs in ((gsd, ref["g_state"]), (dsd, ref["d_state"])):
            for k, v in rs.items():
                if k.endswith("._u") or k.endswith("._v"):
                    assert rel_err(sd[k], v) < 2e-4, k
                elif "running_" in k:
                    assert rel_err(sd[k], v) < tol_out, k
                elif "num_batches" in k:
                    assert int(sd[k]) == int(v), k
        if tc:
            # 1xTF32 operands end to end through batch-stat BatchNorm at fresh init: per-parameter agreement is chaotic
            # (1.3e-2 on the forward already); check the whole gradient vector, tight checks are per block
            assert global_grad_error(got["d_grads"], ref["d_grads"]) < 0.25
            assert global_grad_error(got["g_grads"], ref["g_grads"]) < 0.3
        else:
            compare_grads(got["d_grads"], ref["d_grads"], 2e-3, 5e-2, zero_floor=1e-6)
            # G gradients through the whole train-mode net are chaotic at the 1e-2 level even reference-vs-reference
            # (tests/test_oracle.py); the tight gradient checks are the per-block tests below
            compare_grads(got["g_grads"], ref["g_grads"], 5e-2, 2e-1, zero_floor=1e-5)
    gen.cpu(); disc.cpu()


# ------------------------------------------------------------------------------------------------ per-block parity
from block_cases import block_cases, run_block_case, run_conv_gru_case  # noqa: E402


@pytest.mark.parametrize("case", block_cases(False), ids=lambda c: c[0])
@pytest.mark.parametrize("training", [True, False])
def test_block_simt_fp32(cuda_backend, case, training):
    """Narrow channel counts: served by the fp32 SIMT kernels -> fp32-level agreement with the oracle."""
    from skillful_nowcasting_b200 import ops

    ops.config.conv_algo = ops.config.wgrad_algo = 1
    try:
        run_block_case(case, training, "cuda", 2e-5, 3e-4)
    finally:
        ops.config.conv_algo = ops.config.wgrad_algo = 0


@pytest.mark.parametrize("case", block_cases(True), ids=lambda c: c[0])
@pytest.mark.parametrize("training", [True, False])
def test_block_tensor_core(cuda_backend, case, training):
    """Wide channel counts: served by the tcgen05 kind::tf32 kernels (1xTF32 operands, fp32 accumulate).
    Forward within 1e-3 rel (the north-star tolerance); gradients of a freshly initialised block with batch-stat
    BatchNorm amplify the 2^-11 operand rounding (SURVEY.md section 7 measures 4e-3..1.2e-2 per block), and a few ReLU masks
    of near-zero pre-activations flip: 2e-2 in L2 norm, 1.5e-1 on the worst single element."""
    run_block_case(case, training, "cuda", 1e-3, 1.5e-1, tol_buf=1e-3, tol_l2=5e-2)


@pytest.mark.parametrize("fused", [True, False])
def test_conv_gru_simt_and_tensor_core(cuda_backend, fused):
    """fused: ops.gru_sequence (one conv for both gates, one autograd node, batched weight gradients); False: per-step wiring."""
    from skillful_nowcasting_b200 import ops

    old = ops.config.gru_sequence
    ops.config.gru_sequence = fused
    ops.config.conv_algo = ops.config.wgrad_algo = 1
    try:
        run_conv_gru_case("cuda", 2e-5, 3e-4)
        ops.config.conv_algo = ops.config.wgrad_algo = 0
        run_conv_gru_case("cuda", 1e-3, 1.5e-1, cx=64, ch=32, s=16, T=4, tol_l2=5e-2)
    finally:
        ops.config.conv_algo = ops.config.wgrad_algo = 0
        ops.config.gru_sequence = old


# ------------------------------------------------------------------------------------------------ modes
# "simt": fp32 FMA kernels; "tf32": AUTO dispatch, 1xTF32 operands (the benchmark's mode); "tf32-patch": AUTO dispatch preferring the
# halo-patch / CTA-pair kernels the benchmark shapes run on (they are not profitable, hence not chosen, at C1's sizes);
# "3xtf32": parity mode, error-compensated operand pairs on the plain tcgen05 kernels.
def _set_mode(be, mode):
    from skillful_nowcasting_b200 import ops

    ops.clear_pack_cache()
    ops.config.conv_algo = ops.config.wgrad_algo = 1 if mode == "simt" else 0
    ops.config.precision = 1 if mode == "3xtf32" else 0
    be.set_option("prefer_patch", 1 if mode == "tf32-patch" else -1)


@pytest.mark.parametrize("mode", ["tf32-patch", "3xtf32"])
@pytest.mark.parametrize("training", [False, True], ids=["eval", "train"])
def test_c1_gan_patch_kernels_and_parity_mode(cuda_backend, c1_state, mode, training):
    """The C1 GAN forward / losses / gradients against the oracle (a) with the halo-patch + CTA-pair kernels forced into AUTO
    dispatch, so the kernels the benchmark relies on see oracle data, and (b) in 3xTF32 parity mode, whose tolerances are the
    fp32 ones: eval 1e-4 (north star: 1e-3), train-mode output 2e-3, gradients as for the fp32 SIMT path."""
    gen, disc, g0, d0 = c1_state
    x, y = c1_inputs()
    ref = oracle_gan_forward(g0, d0, x, y, C1, training, seed=2)
    gen.load_state_dict(g0); disc.load_state_dict(d0)
    gen.cuda(); disc.cuda()
    _set_mode(cuda_backend, mode)
    got = module_gan_forward(gen, disc, x, y, training, seed=2, device="cuda")
    x3 = mode == "3xtf32"
    tol_out = (2e-3 if x3 else 5e-2) if training else (1e-4 if x3 else 1e-3)
    e_out, e_sc = rel_err(got["out"], ref["out"]), rel_err(got["scores"], ref["scores"])
    print(f"\nC1 {mode} {'train' if training else 'eval'}: out {e_out:.2e} scores {e_sc:.2e}")
    assert e_out < tol_out, e_out
    assert e_sc < tol_out * (1 if x3 else 5), e_sc
    for k in ("d_loss", "grid", "g_loss"):
        assert rel_err(got[k], ref[k]) < tol_out, (k, float(got[k]), float(ref[k]))
    if training:
        gd, gg = global_grad_error(got["d_grads"], ref["d_grads"]), global_grad_error(got["g_grads"], ref["g_grads"])
        print(f"   global gradient error D {gd:.2e} G {gg:.2e}")
        if x3:
            compare_grads(got["d_grads"], ref["d_grads"], 2e-3, 5e-2, zero_floor=1e-6)
            # G gradients through the whole train-mode net are chaotic at the 1e-2 level even reference-vs-reference (tests/test_oracle.py:
            # single parameters such as att_block.gamma move by 6e-2): whole-vector bound.  Round 2 measured D 5.5e-4, G 2.9e-2.
            assert gd < 2e-3 and gg < 6e-2, (gd, gg)
        else:
            assert gd < 0.25 and gg < 0.3, (gd, gg)
    gen.cpu(); disc.cpu()


@pytest.mark.parametrize("mode", ["simt", "tf32", "3xtf32"])
def test_gan_step_against_oracle(cuda_backend, c1_state, mode):
    """ONE full training.gan_step on C1 (2 discriminator updates + 1 generator update through the fused Adam) against
    oracle.gan_step from the same state, seeds and RNG order: losses, every parameter, Adam exp_avg / exp_avg_sq / step, all
    mutated buffers, and the requires_grad toggling of the discriminator (ref: dgmr/dgmr.py:137-218, :292-300)."""
    from parity_util import GAN_STEP_TOL_FP32, GAN_STEP_TOL_TF32, compare_gan_step, module_gan_step, oracle_gan_step

    gen, disc, g0, d0 = c1_state
    x, y = c1_inputs()
    ref = oracle_gan_step(g0, d0, x, y, C1, seed=4)
    gen.load_state_dict(g0); disc.load_state_dict(d0)
    gen.cuda(); disc.cuda()
    _set_mode(cuda_backend, mode)
    got = module_gan_step(gen, disc, x, y, seed=4, device="cuda")
    # 1xTF32 end to end through batch-stat BatchNorm at fresh init is chaotic (see module docstring): loose there, tight elsewhere
    rep = compare_gan_step(got, ref, g0, d0, GAN_STEP_TOL_TF32 if mode == "tf32" else GAN_STEP_TOL_FP32)
    print(f"\nGAN STEP {mode}: " + " ".join(f"{k} {v:.2e}" for k, v in rep.items()))
    gen.cpu(); disc.cpu()
    gen.load_state_dict(g0); disc.load_state_dict(d0)


@pytest.mark.parametrize("mode", ["simt", "tf32", "tf32-patch", "3xtf32"])
@pytest.mark.parametrize("which", ["spatial", "temporal"])
@pytest.mark.parametrize("training", [False, True], ids=["eval", "train"])
def test_discriminators_separately(cuda_backend, which, training, mode):
    """Spatial and temporal discriminator each on its own, so that a score error is attributable.

    What round 2 found (this is the 1.3e-3 the round-1 smoke test printed): the TEMPORAL scores meet the north-star 1e-3 in every mode
    (1xTF32 3.9e-4, 3xTF32 1.0e-5).  The SPATIAL eval score is a heavily cancelling sum -- 8 frames x 768 BatchNorm-ed features x a
    spectrally normalised weight vector add up to 4e-3 while the terms are O(1e-2..1e-1) each -- so ANY 2^-11 operand rounding moves it by
    ~1e-2: rounding the conv operands to TF32 inside the fp32 CPU oracle (no GPU involved, tests/test_oracle.py::
    test_tf32_operand_rounding_alone_moves_spatial_scores) shifts it by 1.3e-2, the 1xTF32 kernels by 1.9e-2, in the same direction.
    cuDNN's default TF32 convolutions give the reference the same sensitivity.  Bounds: 1xTF32 4e-2 (3x the emulated shift), parity mode
    1e-3 (measured 3.4e-4), fp32 SIMT 2e-4 (measured 3.4e-5).  Train mode: BatchNorm1d over 4 rows amplifies further (fp32 emulator vs oracle
    gradients already differ by 3e-3); 1xTF32 gradients agree only as a whole vector (measured L2 0.17 - 0.32), parity mode to 4e-3 / 8e-3."""
    from parity_util import run_discriminator_case

    _set_mode(cuda_backend, mode)
    if not training:
        tol = {"simt": 2e-4, "3xtf32": 1e-3}.get(mode, 4e-2 if which == "spatial" else 1e-3)
    else:
        # 1xTF32 train-mode scores: BatchNorm1d over 4 rows amplifies the 2^-11 operand rounding (measured 2.2e-2 .. 7.4e-2 over seeds and kernels)
        tol = {"simt": 1e-3, "3xtf32": 1e-3}.get(mode, 1.5e-1)
    out = run_discriminator_case(which, training, "cuda", tol, tol_grad_l2=(2e-2 if mode in ("simt", "3xtf32") else 0.5))
    print(f"\nDISC {which} {mode} {'train' if training else 'eval'}: scores rel err {out['fwd']:.2e}" +
          (f" grad L2 {out['grad_l2']:.2e}" if training else ""))


@pytest.mark.parametrize("case", block_cases(True), ids=lambda c: c[0])
def test_block_parity_mode(cuda_backend, case):
    """Wide blocks in 3xTF32 parity mode: forward and gradients at fp32-level tolerances (per-block gradients <= 1e-3 rel)."""
    _set_mode(cuda_backend, "3xtf32")
    run_block_case(case, True, "cuda", 2e-5, 1e-3, tol_buf=1e-4, tol_l2=3e-4)


def test_conv_gru_parity_mode(cuda_backend):
    _set_mode(cuda_backend, "3xtf32")
    run_conv_gru_case("cuda", 2e-5, 1e-3, cx=64, ch=32, s=16, T=4, tol_l2=3e-4)


def test_pretrained_round_trip_on_gpu(cuda_backend, c1_state, tmp_path):
    """save_pretrained / from_pretrained (ref: README.md:57-69, tests/test_model.py:341-399) with the model on the GPU: same keys,
    same tensors, same eval forward afterwards."""
    import skillful_nowcasting_b200 as B

    gen, disc, g0, d0 = c1_state
    gen.load_state_dict(g0)
    gen.cuda().eval()
    x, _ = c1_inputs()
    torch.manual_seed(9)
    a = gen(x.cuda())
    for name, mod, cls in (("sampler", gen.sampler, B.Sampler), ("ctx", gen.conditioning_stack, B.ContextConditioningStack),
                           ("lat", gen.latent_stack, B.LatentConditioningStack)):
        mod.save_pretrained(tmp_path / name)
        new = cls.from_pretrained(tmp_path / name)
        assert not new.training
        sd_a, sd_b = mod.state_dict(), new.state_dict()
        assert list(sd_a) == list(sd_b)
        for k in sd_a:
            assert torch.equal(sd_a[k].cpu(), sd_b[k].cpu()), k
    g2 = B.Generator(B.ContextConditioningStack.from_pretrained(tmp_path / "ctx"), B.LatentConditioningStack.from_pretrained(tmp_path / "lat"),
                     B.Sampler.from_pretrained(tmp_path / "sampler")).cuda().eval()
    torch.manual_seed(9)
    b = g2(x.cuda())
    # Identical weights and latent draw, yet not bit-equal: the tap-split ConvGRU convolutions accumulate with fp32 atomics whose order is
    # not fixed, and a 1e-7 difference that lands on a tf32 rounding boundary of a conv operand becomes a 2^-11 (4.9e-4) difference of
    # that element.  Round 2 measured 6.0e-4 between two runs -- this is the run-to-run noise floor of the 1xTF32 path.
    assert rel_err(b, a) < 3e-3
    gen.cpu()


def test_graphed_generator_matches_eager_and_oracle(cuda_backend, c1_state):
    """Inference path: the eval-mode generator forward replayed from a CUDA graph equals the eager forward from the same
    seed (up to the fp32-atomic ordering of the tap-split ConvGRU convolutions) and the oracle within the 1xTF32 eval tolerance."""
    from skillful_nowcasting_b200.inference import GraphedGenerator

    gen, disc, g0, d0 = c1_state
    gen.load_state_dict(g0)
    gen.cuda().eval()
    x, y = c1_inputs()
    xc = x.cuda()
    runner = GraphedGenerator(gen, xc)
    n0 = cuda_backend.launches
    torch.manual_seed(2)
    out_g = runner(xc).clone()
    assert cuda_backend.launches == n0, "a graph replay must not issue C-ABI launches from the host"
    torch.manual_seed(2)
    with torch.no_grad():
        out_e = gen(xc)
    assert rel_err(out_g, out_e) < 3e-3          # run-to-run noise floor of the 1xTF32 path (see test_pretrained_round_trip_on_gpu): 5.4e-4
    ref = oracle_gan_forward(g0, d0, x, y, C1, False, seed=2)
    assert rel_err(out_g, ref["out"]) < 1e-3
    torch.manual_seed(3)                      # a second replay takes a fresh latent draw (the reference's CPU RNG order) into the static buffer
    runner(xc)
    torch.manual_seed(3)
    assert torch.equal(runner.z, gen.latent_stack.sample_z(xc))
    gen.cpu()


def test_graphed_train_mode_generator_matches_eager(cuda_backend, c1_state):
    """The D phase's gradient-free TRAIN-mode generator forward (ref: dgmr/dgmr.py:159-160) replayed from a CUDA graph: outputs of two
    successive replays and the state they leave behind (spectral-norm u / v after 2 x T power iterations, BatchNorm running statistics,
    num_batches_tracked) equal two eager forwards from the same start; the capture's warm-up forwards leave no trace; gan_step accepts it.
    Bounds: run-to-run noise of the 1xTF32 path (fp32 atomics of the tap-split convolutions) through train-mode BatchNorm."""
    from skillful_nowcasting_b200.inference import GraphedGenerator

    gen, disc, g0, d0 = c1_state
    _set_mode(cuda_backend, "tf32")
    gen.load_state_dict(g0)
    gen.cuda().train()
    x, y = c1_inputs()
    xc = x.cuda()
    runner = GraphedGenerator(gen, xc, train_mode=True)
    assert runner.launches > 100
    for k, v in gen.state_dict().items():
        assert torch.equal(v.cpu(), g0[k]), f"{k} changed during the capture"
    n0 = cuda_backend.launches
    outs_g = []
    for seed in (2, 3):
        torch.manual_seed(seed)
        outs_g.append(runner(xc).clone())
    assert cuda_backend.launches == n0
    st_g = {k: v.detach().cpu().clone() for k, v in gen.state_dict().items()}
    gen.load_state_dict(g0)
    outs_e = []
    with torch.no_grad():
        for seed in (2, 3):
            torch.manual_seed(seed)
            outs_e.append(gen(xc))
    st_e = {k: v.detach().cpu().clone() for k, v in gen.state_dict().items()}
    for a, b in zip(outs_g, outs_e):
        assert rel_err(a, b) < 2e-2
    assert rel_err(outs_g[0], outs_g[1]) > 1e-3      # a fresh latent per replay
    moved = 0
    for k in st_e:
        if st_e[k].dtype.is_floating_point:
            assert rel_err(st_g[k], st_e[k]) < 2e-2, k
            moved += int(not torch.equal(st_e[k], g0[k]))
        else:
            assert torch.equal(st_g[k], st_e[k]), k
    assert moved > 10                                # the buffers did advance (u, v, running statistics)
    gen.cpu()
