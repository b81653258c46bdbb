"""N > 1 path on CPU: world_size-2 gloo run of the batch-sharded GAN step (host logic through the ABI emulator).
Checks the one collective of the path: after a step the replicas hold identical parameters, equal to a single-process
step whose flat gradients are the mean of the two ranks' gradients."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build():
    import skillful_nowcasting_b200 as B

    torch.manual_seed(0)
    gen = B.Generator(B.ContextConditioningStack(input_channels=1, output_channels=32),
                      B.LatentConditioningStack(shape=(8, 4, 4), output_channels=288),
                      B.Sampler(forecast_steps=2, latent_channels=288, context_channels=32))
    disc = B.SpatialDiscriminator(input_channels=1, num_timesteps=2)  # the full D works the same; this keeps the test fast
    return gen, disc


def _one_step(rank, world, port, out_path):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from emu_backend import EmuBackend
    from skillful_nowcasting_b200 import _lib, losses
    from skillful_nowcasting_b200.training import Adam

    _lib.set_backend(EmuBackend())
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    gen, disc = _build()
    gen.train(); disc.train()
    g_opt = Adam(gen.parameters(), lr=5e-5, betas=(0.0, 0.999))
    d_opt = Adam(disc.parameters(), lr=2e-4, betas=(0.0, 0.999))
    torch.manual_seed(100 + rank)  # each rank its own shard of the batch
    x, y = torch.rand(1, 4, 1, 128, 128), torch.rand(1, 2, 1, 128, 128)
    g_opt.zero_grad(); d_opt.zero_grad()
    pred = gen(x)
    scores = disc(torch.cat([torch.cat([x, y], 1), torch.cat([x, pred], 1)], 0))
    b = scores.shape[0] // 2
    loss = losses.loss_hinge_disc(scores[b:], scores[:b]) + pred.mean()
    loss.backward()
    local = {"g": g_opt.flat_g.clone(), "d": d_opt.flat_g.clone()}
    p_before = {"g": g_opt.flat_p.clone(), "d": d_opt.flat_p.clone()}
    g_opt.step(); d_opt.step()
    torch.save({"local": local, "summed": {"g": g_opt.flat_g.clone(), "d": d_opt.flat_g.clone()},
                "p_before": p_before, "p_after": {"g": g_opt.flat_p.clone(), "d": d_opt.flat_p.clone()}},
               out_path.format(rank=rank))
    dist.destroy_process_group()


@pytest.mark.timeout(900)
def test_two_rank_step_all_reduces_flat_gradients(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    pat = str(tmp_path / "rank{rank}.pt")
    # BatchNorm statistics stay per replica (as in the reference: no SyncBN); only gradients are exchanged
    mp.spawn(_one_step, args=(2, port, pat), nprocs=2, join=True)
    r0, r1 = torch.load(pat.format(rank=0)), torch.load(pat.format(rank=1))
    for k, lr in (("g", 5e-5), ("d", 2e-4)):
        assert torch.equal(r0["p_before"][k], r1["p_before"][k])                 # identical replicas going in
        assert not torch.equal(r0["local"][k], r1["local"][k])                   # different shards -> different gradients
        assert torch.equal(r0["summed"][k], r0["local"][k] + r1["local"][k])     # the one collective: sum over ranks
        assert torch.equal(r0["summed"][k], r1["summed"][k])
        assert torch.equal(r0["p_after"][k], r1["p_after"][k])                   # identical replicas coming out
        # first Adam step with beta1 = 0: every parameter with a non-zero mean gradient moves by exactly lr
        moved = (r0["p_after"][k] - r0["p_before"][k]).abs()
        nz = r0["summed"][k] != 0
        # (|g| / (|g| + eps) < 1 only for gradients comparable to eps = 1e-8)
        assert float(moved[nz].max()) <= lr * 1.02 and abs(float(moved[nz].median()) - lr) < lr * 0.02
        assert float(moved[~nz].abs().max() if (~nz).any() else 0.0) == 0.0
