"""The three loss functions the reference's training step uses (dgmr/losses.py:158-192, 307-319), on the C ABI."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


def loss_hinge_disc(score_generated: torch.Tensor, score_real: torch.Tensor) -> torch.Tensor:
    """Discriminator hinge loss (ref: dgmr/losses.py:307-313) for ONE score column.
    Inputs [B,1,1] (any matching shape): mean(relu(1 - real)) + mean(relu(1 + generated))."""
    b = score_real.numel()
    both = ops.concat_channels(score_real.reshape(1, b), score_generated.reshape(1, b))  # real rows, then generated
    return ops.hinge_disc(both.reshape(2 * b, 1, 1))


def loss_hinge_disc_both(scores: torch.Tensor) -> torch.Tensor:
    """Sum of the spatial and temporal discriminator hinge losses (ref: dgmr/dgmr.py:166-168).
    scores: [2B,2,1], real rows first."""
    return ops.hinge_disc(scores)


def loss_hinge_gen(score_generated: torch.Tensor) -> torch.Tensor:
    """Generator hinge loss: -mean(score_generated) (ref: dgmr/losses.py:316-319)."""
    return ops.hinge_gen(score_generated)


class GridCellLoss(nn.Module):
    """Grid-cell regulariser (ref: dgmr/losses.py:158-192 with weight_fn of dgmr/dgmr.py:20-33:
    w(y) = max(y + 1, cap); result = ||(gen - y) * w||_1 / T * H * W, precedence as written there)."""

    def __init__(self, weight_fn=None, precip_weight_cap=24.0):
        super().__init__()
        self.precip_weight_cap = precip_weight_cap

    def forward(self, generated_images: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
        return ops.grid_cell(generated_images, targets, float(self.precip_weight_cap))
