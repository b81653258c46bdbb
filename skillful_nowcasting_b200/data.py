"""Data side of the training step (ref: train/run.py:114-158).

The reference's dataset row is `radar_frames [T_all, H, W, C]` (nimrod-uk-1km, C = 1); `extract_input_and_target_frames` (:118-123)
aligns the 18 targets to the END of the window with the 4 inputs right before them, and `TFDataset.__getitem__` (:152-158) moves the
channel axis in front of H, W.  So a sample is the LAST 22 frames of a row, and the step consumes them as one sequence
(`training.gan_step` concatenates inputs and targets again for the discriminator's real half).

`DeviceBatcher` therefore stages exactly that window: B rows -> ONE pinned `[B, 22, C, H, W]` buffer (only the 22 frames that are
used cross the host link), one asynchronous host->device copy on a side stream, double-buffered so that the copy of batch i+1
overlaps step i, and `(images, future)` handed out as the two parts of the device window (dense copies by default, views on request).  No arithmetic happens here (the reference does
none: frames are fed as stored), so there is no kernel on this path -- torch copies on CUDA streams only.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch

NUM_INPUT_FRAMES = 4      # ref: train/run.py:114
NUM_TARGET_FRAMES = 18    # ref: train/run.py:115
WINDOW = NUM_INPUT_FRAMES + NUM_TARGET_FRAMES


def extract_input_and_target_frames(radar_frames):
    """Inputs and targets of one dataset row (ref: train/run.py:118-123): targets = the last 18 frames, inputs = the 4 before them.
    Works on anything sliceable along the first axis (numpy arrays, tensors, lists)."""
    input_frames = radar_frames[-NUM_TARGET_FRAMES - NUM_INPUT_FRAMES: -NUM_TARGET_FRAMES]
    target_frames = radar_frames[-NUM_TARGET_FRAMES:]
    return input_frames, target_frames


def row_to_sample(radar_frames) -> Tuple[np.ndarray, np.ndarray]:
    """What the reference's `TFDataset.__getitem__` returns for a row (ref: train/run.py:152-158): ([4, C, H, W], [18, C, H, W])."""
    input_frames, target_frames = extract_input_and_target_frames(np.asarray(radar_frames))
    return np.moveaxis(input_frames, 3, 1), np.moveaxis(target_frames, 3, 1)


class DeviceBatcher:
    """`put(rows)` stages a batch of dataset rows and starts its host->device copy; `get()` returns `(images [B,4,C,H,W],
    future [B,18,C,H,W])` of the oldest staged batch, the two parts of one device window `[B, 22, C, H, W]` (kept as `.window`).

    depth = number of batches in flight (2: the copy of the next batch overlaps the current step).  A slot is reused `depth` puts
    later; on CUDA the refill first waits for everything queued so far on the stream that was current at the slot's get() -- so call
    put() for a later batch AFTER the step that consumes the previous get() has been enqueued (the natural loop order:
    `x, y = b.get(); step(x, y); b.put(next_rows)`)."""

    def __init__(self, batch: int, height: int, width: int, channels: int = 1, device="cuda", depth: int = 2, dtype=torch.float32):
        self.device = torch.device(device)
        self.cuda = self.device.type == "cuda"
        if self.cuda and not torch.cuda.is_available():
            raise RuntimeError("DeviceBatcher(device='cuda') needs a GPU (there is no CPU path of the training step)")
        shape = (batch, WINDOW, channels, height, width)
        self.depth = max(1, int(depth))
        self._host = [torch.empty(shape, dtype=dtype, pin_memory=self.cuda) for _ in range(self.depth)]
        self._dev = [torch.empty(shape, dtype=dtype, device=self.device) for _ in range(self.depth)]
        self._copied = [torch.cuda.Event() if self.cuda else None for _ in range(self.depth)]
        self._consumer = [None] * self.depth        # the stream a slot was handed out on
        self._stream = torch.cuda.Stream(device=self.device) if self.cuda else None
        self._put = self._got = 0
        self.window = None

    def put(self, rows: Sequence) -> None:
        """rows: B arrays `[T_all >= 22, H, W, C]` (the dataset's `radar_frames`)."""
        if self._put - self._got >= self.depth:
            raise RuntimeError("DeviceBatcher.put: all slots are staged; call get() first")
        slot = self._put % self.depth
        host = self._host[slot]
        if len(rows) != host.shape[0]:
            raise ValueError(f"DeviceBatcher.put: expected {host.shape[0]} rows, got {len(rows)}")
        if self.cuda and self._copied[slot] is not None and self._put >= self.depth:
            self._copied[slot].synchronize()          # the previous copy out of this pinned slot has left the host buffer
        for i, row in enumerate(rows):
            frames = torch.as_tensor(np.asarray(row))
            if frames.dim() != 4 or frames.shape[0] < WINDOW or tuple(frames.shape[1:]) != (host.shape[3], host.shape[4], host.shape[2]):
                raise ValueError(f"DeviceBatcher.put: row {i} has shape {tuple(frames.shape)}, expected [>= {WINDOW}, {host.shape[3]}, "
                                 f"{host.shape[4]}, {host.shape[2]}]")
            host[i].copy_(frames[-WINDOW:].permute(0, 3, 1, 2))      # [22, H, W, C] -> [22, C, H, W]; a reshape when C = 1
        if self.cuda:
            if self._consumer[slot] is not None:
                self._stream.wait_stream(self._consumer[slot])         # the step that read this device slot has finished with it
            with torch.cuda.stream(self._stream):
                self._dev[slot].copy_(host, non_blocking=True)
                self._copied[slot].record(self._stream)
        else:
            self._dev[slot].copy_(host)
        self._put += 1

    def get(self, contiguous: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
        """contiguous (default): `images` / `future` are dense copies of the two parts of the window (two device-side copies of 92 MB at
        the benchmark's batch, ~0.03 ms); False: strided views of `.window` (batch stride = 22 frames)."""
        if self._got >= self._put:
            raise RuntimeError("DeviceBatcher.get: nothing staged; call put() first")
        slot = self._got % self.depth
        if self.cuda:
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(self._copied[slot])
            self._consumer[slot] = cur
        self._got += 1
        self.window = self._dev[slot]
        images, future = self.window[:, :NUM_INPUT_FRAMES], self.window[:, NUM_INPUT_FRAMES:]
        return (images.contiguous(), future.contiguous()) if contiguous else (images, future)


def collate_samples(samples: List[Tuple[np.ndarray, np.ndarray]]) -> Tuple[torch.Tensor, torch.Tensor]:
    """Default-collate equivalent for the reference's per-sample `(input, target)` pairs: ([B,4,C,H,W], [B,18,C,H,W]) tensors."""
    xs = torch.stack([torch.as_tensor(np.ascontiguousarray(s[0])) for s in samples])
    ys = torch.stack([torch.as_tensor(np.ascontiguousarray(s[1])) for s in samples])
    return xs, ys
