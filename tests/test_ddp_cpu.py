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


def _one_step(rank, world, port, out_path, seeds):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from emu_backend import EmuBackend
    from skillful_nowcasting_b200 import _lib, losses
    from skillful_nowcasting_b200.training import Adam

    _lib.set_backend(EmuBackend())
    if world > 1:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    gen, disc = _build()
    gen.train(); disc.train()
    g_opt = Adam(gen.parameters(), lr=5e-5, betas=(0.0, 0.999))
    d_opt = Adam(disc.parameters(), lr=2e-4, betas=(0.0, 0.999))
    flat_grads = []
    buffers0 = [(m, {k: v.clone() for k, v in m.named_buffers()}) for m in (gen, disc)]
    for seed in seeds:  # single-process mode emulates both ranks' data and averages the gradients by hand
        for m, snap in buffers0:  # every rank starts the step from the same u/v and BN running statistics
            for k, v in m.named_buffers():
                v.copy_(snap[k])
        torch.manual_seed(seed)
        x, y = torch.rand(1, 4, 1, 128, 128), torch.rand(1, 2, 1, 128, 128)
        g_opt.zero_grad(); d_opt.zero_grad()
        torch.manual_seed(77)  # same latent / frame draws on every rank for an exact comparison
        pred = gen(x)
        scores = disc(torch.cat([torch.cat([x, y], 1), torch.cat([x, pred], 1)], 0))
        b = scores.shape[0] // 2
        loss = losses.loss_hinge_disc(scores[b:], scores[:b]) + pred.mean()
        loss.backward()
        flat_grads.append((g_opt.flat_g.clone(), d_opt.flat_g.clone()))
    if len(seeds) > 1:
        g_opt.flat_g.copy_(sum(g for g, _ in flat_grads) / len(seeds))
        d_opt.flat_g.copy_(sum(d for _, d in flat_grads) / len(seeds))
    g_opt.step(); d_opt.step()
    torch.save({"g": g_opt.flat_p.clone(), "d": d_opt.flat_p.clone()}, out_path.format(rank=rank))
    if world > 1:
        dist.destroy_process_group()


@pytest.mark.timeout(900)
def test_two_rank_step_matches_mean_gradient_step(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    pat = str(tmp_path / "rank{rank}.pt")
    # NOTE: BatchNorm statistics are per replica (as in the reference: no SyncBN), so each rank sees its own batch
    mp.spawn(_spawn_entry, args=(2, port, pat), nprocs=2, join=True)
    _one_step(0, 1, port, str(tmp_path / "single.pt"), seeds=(100, 101))
    r0, r1, single = (torch.load(pat.format(rank=0)), torch.load(pat.format(rank=1)), torch.load(str(tmp_path / "single.pt")))
    for k in ("g", "d"):
        assert torch.equal(r0[k], r1[k]), f"replicas diverged after the all-reduced step: {(r0[k]-r1[k]).abs().max().item()} n={(r0[k]!=r1[k]).sum().item()}"
        # first Adam step with beta1 = 0 is exactly lr*sign(g): a near-zero gradient whose sign flips with the summation order
        # moves by 2*lr; everything else must agree to rounding
        lr = 5e-5 if k == "g" else 2e-4
        diff = (r0[k] - single[k]).abs()
        assert diff.max().item() <= 2 * lr * 1.01, diff.max().item()
        assert (diff > 1e-6).float().mean().item() < 1e-3, "more than 0.1% of the parameters moved differently"


def _spawn_entry(rank, world, port, pat):
    _one_step(rank, world, port, pat, seeds=(100 + rank,))
