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
