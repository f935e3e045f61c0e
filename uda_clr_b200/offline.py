"""Offline prototype extraction -- the loop body of the reference's ``cal_prototype.py`` (:128-195) over this
package's pooling kernels.

The reference script runs a trained model over the target set and, per batch, thresholds the predictions
(``sigmoid(o_before)[:,1] > 0.5`` disc, ``[:,0] > 0.1`` cup, ``sigmoid(boundary_before) > 0.5`` boundary; :145-151),
pools the decoder features with one cuBLAS ``bmm`` per mask (:156-175: ``mean_b(m . X / (sum m + 1))`` on the 304-channel
``x_bu_feature`` for the boundary and the 305-channel ``x_feature`` for cup and disc), folds the result into a "running
mean" (:177-190) and finally ``torch.save``s ``{'bu', 'cup', 'disc'}`` (:192-195).  The model, the dataset and argparse
stay with the caller; this module is the per-batch arithmetic and the saved dict.

Quirk reproduced by default: the reference's running mean is ``p = (p*n + p) / (n+1)`` with ``p`` the CURRENT batch's
prototype on both sides, i.e. the identity -- the saved vectors are the LAST batch's prototypes.
``running_mean=True`` gives the mean the code evidently intended (count capped at 3000 as in :180, :185, :190).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from .ops import bmm_prototypes

DISC_THRESHOLD = 0.5       # cal_prototype.py:145
CUP_THRESHOLD = 0.1        # cal_prototype.py:146
BOUNDARY_THRESHOLD = 0.5   # cal_prototype.py:150-151
COUNT_CAP = 3000           # cal_prototype.py:180


def offline_masks(o_before: torch.Tensor, boundary_before: torch.Tensor):
    """The three binary masks of one batch: ``(cup_disc [B,2,H,W], boundary [B,1,H,W])`` as float {0,1}; plane 0 = cup
    (``sigmoid > 0.1``), plane 1 = disc (``sigmoid > 0.5``).  Thresholded with ATen's own sigmoid, like the reference."""
    p = torch.sigmoid(o_before.detach())
    cup_disc = torch.stack([p[:, 0] > CUP_THRESHOLD, p[:, 1] > DISC_THRESHOLD], 1).to(torch.float32)
    bu = (torch.sigmoid(boundary_before.detach()) > BOUNDARY_THRESHOLD).to(torch.float32)
    return cup_disc, bu


class OfflinePrototypes:
    """Accumulates ``{'bu', 'cup', 'disc'}`` over batches; ``objective_vectors()`` is what the reference saves."""

    def __init__(self, running_mean: bool = False):
        self.running_mean = running_mean
        self.vec: Dict[str, Optional[torch.Tensor]] = {"bu": None, "cup": None, "disc": None}
        self.num = {"bu": 0, "cup": 0, "disc": 0}

    def _fold(self, key: str, p: torch.Tensor) -> None:
        n = self.num[key]
        if self.running_mean and self.vec[key] is not None:
            p = (self.vec[key] * n + p) / (n + 1)
        # reference (:177-190): (p*n + p)/(n+1) == p -- the current batch's prototype replaces the stored one
        self.vec[key] = p
        self.num[key] = min(n + 1, COUNT_CAP)

    @torch.no_grad()
    def update(self, o_before: torch.Tensor, boundary_before: torch.Tensor, x_bu_feature: torch.Tensor,
               x_feature: torch.Tensor) -> Dict[str, torch.Tensor]:
        """One batch (cal_prototype.py:139-190).  ``o_before [B,2,h,w]``, ``boundary_before [B,1,h,w]`` logits,
        ``x_bu_feature [B,304,h,w]``, ``x_feature [B,305,h,w]``.  cup and disc are pooled in ONE read of ``x_feature``."""
        cup_disc, bu = offline_masks(o_before, boundary_before)
        p_bu = bmm_prototypes(bu, x_bu_feature.detach())[0]
        p_cd = bmm_prototypes(cup_disc, x_feature.detach())
        self._fold("bu", p_bu)
        self._fold("cup", p_cd[0])
        self._fold("disc", p_cd[1])
        return self.objective_vectors()

    def objective_vectors(self) -> Dict[str, torch.Tensor]:
        """``{'bu': [304], 'cup': [305], 'disc': [305]}`` (cal_prototype.py:192)."""
        return {k: v for k, v in self.vec.items() if v is not None}

    def save(self, path: str) -> None:
        torch.save(self.objective_vectors(), path)      # cal_prototype.py:194-195
