"""Data side (ref: train/run.py:114-158): the host mirror of the reference's row -> sample transform and the batch staging,
checked against the reference's own expressions restated with numpy (train/run.py imports wandb / datasets and cannot be imported here)."""
import numpy as np
import pytest
import torch

from skillful_nowcasting_b200 import data as D


def _rows(b, t_all, h, w, c, seed=0):
    rng = np.random.default_rng(seed)
    return [rng.random((t_all, h, w, c), dtype=np.float32) for _ in range(b)]


@pytest.mark.parametrize("t_all,c", [(24, 1), (22, 1), (30, 3)])
def test_row_to_sample_is_the_reference_transform(t_all, c):
    row = _rows(1, t_all, 8, 6, c)[0]
    # ref: train/run.py:118-123 (slices) and :152-158 (np.moveaxis(frames, [0, 1, 2, 3], [0, 2, 3, 1]))
    exp_in = np.moveaxis(row[-18 - 4:-18], [0, 1, 2, 3], [0, 2, 3, 1])
    exp_tg = np.moveaxis(row[-18:], [0, 1, 2, 3], [0, 2, 3, 1])
    got_in, got_tg = D.row_to_sample(row)
    assert got_in.shape == (4, c, 8, 6) and got_tg.shape == (18, c, 8, 6)
    assert np.array_equal(got_in, exp_in) and np.array_equal(got_tg, exp_tg)
    a, b = D.extract_input_and_target_frames(torch.from_numpy(row))     # tensors slice the same way
    assert torch.equal(a, torch.from_numpy(row[-22:-18])) and torch.equal(b, torch.from_numpy(row[-18:]))


@pytest.mark.parametrize("c", [1, 2])
def test_device_batcher_equals_per_sample_collate(c):
    """Window staging (one [B, 22, C, H, W] buffer, views handed out) == default-collating the reference's per-sample pairs; slots rotate."""
    b, h, w = 3, 8, 6
    batcher = D.DeviceBatcher(b, h, w, channels=c, device="cpu", depth=2)
    batches = [_rows(b, 22 + i, h, w, c, seed=i) for i in range(4)]
    batcher.put(batches[0])
    batcher.put(batches[1])
    with pytest.raises(RuntimeError):
        batcher.put(batches[2])                      # both slots staged
    for i in range(4):
        x, y = batcher.get(contiguous=bool(i % 2))
        assert x.is_contiguous() == bool(i % 2)
        ex, ey = D.collate_samples([D.row_to_sample(r) for r in batches[i]])
        assert x.shape == (b, 4, c, h, w) and y.shape == (b, 18, c, h, w)
        assert torch.equal(x, ex) and torch.equal(y, ey)
        assert torch.equal(batcher.window, torch.cat([ex, ey], dim=1))     # the sequence the discriminator's real half consumes
        if i + 2 < 4:
            batcher.put(batches[i + 2])
    with pytest.raises(RuntimeError):
        batcher.get()
    with pytest.raises(ValueError):
        batcher.put(_rows(b, 21, h, w, c))           # a row shorter than the 22-frame window
    with pytest.raises(ValueError):
        batcher.put(_rows(b - 1, 22, h, w, c))
