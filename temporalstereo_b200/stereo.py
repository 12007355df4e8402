"""`StereoEngine` — the `forward(left, right, prev_state) -> disparity` surface over the B200 hot path, and the
full-resolution post-processing of the reference meta-architecture.

The reference has no module with that signature; its per-frame forward is
`TemporalStereo.forward(batch, outputs, is_train, timestamp)` (projects/TemporalStereo/TemporalStereo.py:282-324):
backbone -> update_map -> aggregation -> bilinear up-sampling of every disparity to full resolution (:305-309).  This
wrapper is that sequence with the backbone left pluggable (it is outside the hot path): pass any callable
with the reference backbone's contract `backbone(l_img, r_img, prev_info) -> (l_fms, r_fms, prev_info)`
(architecture/modeling/backbone/TemporalStereo.py:142-162), or hand the feature pyramids in directly.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch
import torch.nn as nn

from . import ops, temporal
from .aggregation import TEMPORALSTEREO


def upsample_disps(disps: Sequence[torch.Tensor], full_h: int, full_w: int) -> List[torch.Tensor]:
    """Every disparity map to full resolution, values rescaled by the width ratio
    (projects/TemporalStereo/TemporalStereo.py:305-309: F.interpolate(d * full_w / dw, (full_h, full_w), 'bilinear',
    align_corners=True)).  Maps already at full resolution are returned as they are."""
    out = []
    for d in disps:
        dh, dw = d.shape[-2:]
        if (dh, dw) == (full_h, full_w):
            out.append(d)
        else:
            out.append(ops.bilinear_resize(d.contiguous(), (full_h, full_w), mul=float(full_w), div=float(dw)))
    return out


class StereoEngine(nn.Module):
    """forward(left, right, prev_state) -> full-resolution disparity [B,1,H,W]; `prev_state` (a dict, the reference's
    `prev_info`) is updated in place and carries prev_disp / cost_memory / local_map to the next frame, like
    `video_inference.py:253-301` does.  The whole reference 6-tuple of the frame stays in `self.last`."""

    def __init__(self, aggregation: Optional[TEMPORALSTEREO] = None, backbone: Optional[Callable] = None,
                 with_previous: bool = True, use_past_cost: bool = True, local_map_size: int = 3):
        super().__init__()
        self.aggregation = aggregation if aggregation is not None else TEMPORALSTEREO()
        self.backbone = backbone
        self.with_previous, self.use_past_cost, self.local_map_size = with_previous, use_past_cost, local_map_size
        self.last: dict = {}

    @torch.no_grad()
    def forward(self, left, right, prev_state: Optional[dict] = None, *, feats=None, pose: Optional[dict] = None):
        """left / right: images [B,3,H,W] (ImageNet-normalised fp32).  feats = (left_feats, right_feats), each
        [1/4 (64ch), 1/8 (128ch), 1/16 (256ch)], when no backbone is installed.  pose = dict(K [B,4,4], T [B,4,4] of this
        frame, inv_T_prev [B,4,4] of the previous one, baseline [B,...]) switches the temporal warp on for frames that
        have a previous state (projects/TemporalStereo/TemporalStereo.py:292-294)."""
        if prev_state is None:
            prev_state = {}
        if feats is None:
            if self.backbone is None:
                raise ValueError("StereoEngine needs either a backbone or feats=(left_feats, right_feats)")
            lf, rf, prev_state = self.backbone(left, right, prev_state)
        else:
            lf, rf = feats
        H, W = left.shape[-2:]
        if self.with_previous and pose is not None and "prev_disp" in prev_state:
            temporal.update_map(prev_state, pose["K"], pose["T"], pose["inv_T_prev"], pose["baseline"], H, W,
                                use_past_cost=self.use_past_cost, local_map_size=self.local_map_size)
        disps, costs, samples, offs, ranges, prev_state = self.aggregation(lf, rf, left, right, prev_state)
        full = upsample_disps(disps, H, W)
        self.last = {"disps": full, "costs": costs, "disp_samples": samples, "offsets": offs, "search_ranges": ranges,
                     "prev_info": prev_state}
        return full[0]
