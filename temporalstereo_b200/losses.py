"""The reference's two training-loss terms evaluated on the device, FORWARD ONLY (the engine has no
backward, so these serve validation / logging of the per-level losses; they are not differentiable).

Same constructor arguments, call signatures and result dicts as the reference classes
(architecture/modeling/losses/smooth_l1_loss.py:9-94, warsserstein_distance_loss.py:9-113); one kernel launch per level, the
ground truth scaled and pooled onto each level's grid inside the kernel, results as 0-dim device tensors (no read-back).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Union

import torch

from . import ops


class _LevelLoss:
    def __init__(self, max_disp: int = 192, start_disp: int = 0, global_weight: float = 1.0,
                 weights: Optional[Sequence[float]] = None, sparse: bool = False):
        self.max_disp, self.start_disp, self.global_weight = max_disp, start_disp, global_weight
        self.weights, self.sparse = weights, sparse

    @classmethod
    def from_config(cls, cfg):
        """Same keys and defaults as the reference `from_config`s."""
        get = cfg.get if hasattr(cfg, "get") else (lambda k, d: getattr(cfg, k, d))
        return cls(get("MAX_DISP", 192), get("START_DISP", 0), get("GLOBAL_WEIGHT", 1.0), get("WEIGHTS", None), get("SPARSE", False))

    def _weighted(self, name: str, per_level: List[torch.Tensor]) -> dict:
        w = self.weights if self.weights is not None else [1.0] * len(per_level)
        return {f"{name}_lvl{i}": w[i] * v * self.global_weight for i, v in enumerate(per_level)}


class DispSmoothL1Loss(_LevelLoss):
    """`loss(estDisp, gtDisp)` -> {"l1_loss_lvl{i}": ...} (reference smooth_l1_loss.py:76-94)."""

    def __call__(self, estDisp: Union[torch.Tensor, Sequence[torch.Tensor]], gtDisp: torch.Tensor) -> dict:
        est = list(estDisp) if isinstance(estDisp, (list, tuple)) else [estDisp]
        gt = gtDisp.contiguous()
        return self._weighted("l1_loss", [ops.loss_smooth_l1(e.contiguous(), gt, self.max_disp, self.start_disp, self.sparse) for e in est])


class WarssersteinDistanceLoss(_LevelLoss):
    """`loss(estCosts, estOffsets, dispSamples, gtDisp)` -> {"wars_loss_lvl{i}": ...} (reference warsserstein_distance_loss.py:83-113)."""

    def __call__(self, estCosts, estOffsets, dispSamples, gtDisp: torch.Tensor) -> dict:
        costs = list(estCosts) if isinstance(estCosts, (list, tuple)) else [estCosts]
        offs = list(estOffsets) if isinstance(estOffsets, (list, tuple)) else [estOffsets]
        smps = list(dispSamples) if isinstance(dispSamples, (list, tuple)) else [dispSamples] * len(costs)
        assert len(costs) == len(offs) == len(smps), (len(costs), len(offs), len(smps))
        gt = gtDisp.contiguous()
        return self._weighted("wars_loss", [ops.loss_wasserstein(c.contiguous(), o.contiguous(), s.contiguous(), gt, self.max_disp,
                                                                 self.start_disp, self.sparse) for c, o, s in zip(costs, offs, smps)])
