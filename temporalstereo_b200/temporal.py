"""Pose-conditioned temporal warp of the recurrent state, backed by libtstereo.so.

Three drop-in levels (SURVEY.md §8b "Temporal-warp boundary"):
  * `FunctionSoftsplat(tenInput, tenFlow, tenMetric, strType)` — reference
    architecture/modeling/layers/softsplat.py:334-360 ('softmax' mode, the only one the model uses),
  * `project_to_3d(depth, K, inv_K, T_target_to_source)` — reference
    architecture/modeling/layers/inverse_warp.py:92-178 (returns the keys `update_map` consumes),
  * `update_map(prev_info, K, T_now, inv_T_prev, baseline, full_h, full_w, ...)` — the fused form of
    reference projects/TemporalStereo/TemporalStereo.py:326-461, and `TemporalWarpMixin`, which
    overrides the LightningModule method with the same `(batch, prev_info, timestamp)` signature.

Unlike the reference's CuPy launch, every kernel runs on torch's current CUDA stream.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops

EXPMAX = 50.0


def FunctionSoftsplat(tenInput: torch.Tensor, tenFlow: torch.Tensor, tenMetric: Optional[torch.Tensor],
                      strType: str) -> torch.Tensor:
    assert tenMetric is None or tenMetric.shape[1] == 1
    assert strType in ["summation", "average", "linear", "softmax"]
    if strType != "softmax":
        raise NotImplementedError("libtstereo implements the 'softmax' splat (the only mode TemporalStereo uses)")
    return ops.softsplat(tenInput.contiguous(), tenFlow.contiguous(), tenMetric.contiguous())


def project_to_3d(depth: torch.Tensor, K: torch.Tensor, inv_K: Optional[torch.Tensor] = None,
                  T_target_to_source: Optional[torch.Tensor] = None, eps: float = 1e-7) -> dict:
    """Returns {'triangular_depth' [B,C,H,W], 'optical_flow' [B,2C,H,W]} — the keys `update_map` reads
    (TemporalStereo.py:358-360, 373-375, 406-412).  `inv_K` is recomputed on the device from `K`."""
    if T_target_to_source is None:
        raise NotImplementedError("project_to_3d without a transform is not on the hot path")
    if eps != 1e-7:
        raise NotImplementedError("libtstereo project_to_3d uses the reference default eps = 1e-7")
    B = depth.shape[0]
    dev = depth.device
    K4 = K
    if K.shape[-1] == 3:
        K4 = torch.eye(4, device=dev).repeat(B, 1, 1)
        K4[:, :3, :3] = K
    eye = torch.eye(4, device=dev).repeat(B, 1, 1).contiguous()
    one = torch.ones((B,), device=dev)
    params = ops.pose_prep(K4.contiguous().float(), T_target_to_source.contiguous().float(), eye, one, 1.0)
    flow, tri = ops.project_depth(depth.contiguous().float(), params)
    return {"triangular_depth": tri, "optical_flow": flow}


def update_map(prev_info: dict, K: torch.Tensor, T_now: torch.Tensor, inv_T_prev: torch.Tensor,
               baseline: torch.Tensor, full_h: int, full_w: int, use_past_cost: bool = True,
               local_map_size: int = 3, with_previous: bool = True, fused: bool = True) -> dict:
    """Temporal warp (reference projects/TemporalStereo/TemporalStereo.py:326-461): re-projects the previous
    disparity, the stored top-2 samples/costs and the local map into the current camera and softmax-splats them.
    Mutates and returns `prev_info`.  `fused` (default): three launches for the whole warp (ops.update_map_fused);
    otherwise — or when the cost memory and the local map live at different resolutions — the per-stage operators
    (pose_prep / reproject_disp / splat_metric / softsplat: the drop-ins of project_to_3d and FunctionSoftsplat)."""
    if not with_previous:
        return prev_info
    prev_disp = prev_info["prev_disp"].detach().contiguous().float()
    K = K.contiguous().float()
    T_now, inv_T_prev = T_now.contiguous().float(), inv_T_prev.contiguous().float()
    baseline = baseline.float().reshape(-1).contiguous()
    memory = prev_info.get("cost_memory", None) if use_past_cost else None
    lm = prev_info.get("local_map", None) if local_map_size > 0 else None
    hw_mem = tuple(memory["disp_sample"].shape[-2:]) if memory is not None else None
    hw_lm = (tuple(lm.shape[-2:]) if lm is not None else (full_h // 8, full_w // 8)) if local_map_size > 0 else None
    if fused and (hw_mem is None or hw_lm is None or hw_mem == hw_lm) and (hw_mem or hw_lm):
        n_out = min((lm.shape[1] if lm is not None else 0) + 1, local_map_size) if local_map_size > 0 else 0
        ds = memory["disp_sample"].detach().contiguous() if memory is not None else None
        cv = memory["cost_volume"].detach().contiguous() if memory is not None else None
        ws, wc, wl = ops.update_map_fused(prev_disp, K, T_now, inv_T_prev, baseline, ds, cv,
                                          lm.contiguous() if lm is not None else None, n_out, hw_mem or hw_lm)
        prev_info["cost_memory"] = {"disp_sample": ws, "cost_volume": wc} if memory is not None else None
        prev_info["use_past_cost"] = use_past_cost
        if local_map_size > 0:
            prev_info["local_map"] = wl
            prev_info["local_map_size"] = local_map_size
        return prev_info
    return _update_map_staged(prev_info, prev_disp, K, T_now, inv_T_prev, baseline, full_h, full_w, use_past_cost, local_map_size)


def _update_map_staged(prev_info, prev_disp, K, T_now, inv_T_prev, baseline, full_h, full_w, use_past_cost, local_map_size):
    def state_at(h, w):
        params = ops.pose_prep(K, T_now, inv_T_prev, baseline, full_w / w)
        pd = ops.bilinear_resize(prev_disp, (h, w), mul=w, div=prev_disp.shape[-1])
        return params, pd, ops.splat_metric(pd)

    memory = prev_info.get("cost_memory", None)
    if use_past_cost and memory is not None:                                   # update_past_cost (:386-426)
        ds = memory["disp_sample"].detach().contiguous()
        cv = memory["cost_volume"].detach().contiguous()
        B, c, h, w = ds.shape
        params, pd, metric = state_at(h, w)
        flow, _ = ops.reproject_disp(pd, params, True, False)
        packed = torch.empty((B, 2 * c, h, w), device=ds.device, dtype=torch.float32)
        ops.reproject_disp(ds, params, False, True, out=packed, c_off=0)
        packed[:, c:].copy_(cv)
        warped = ops.softsplat(packed, flow, metric)
        memory = {"disp_sample": warped[:, :c], "cost_volume": warped[:, c:]}
    elif not use_past_cost:
        memory = None
    prev_info["cost_memory"] = memory
    prev_info["use_past_cost"] = use_past_cost

    if local_map_size > 0:                                                     # update_local_map (:340-384)
        lm = prev_info.get("local_map", None)
        h, w = (lm.shape[-2:] if lm is not None else (full_h // 8, full_w // 8))
        params, pd, metric = state_at(h, w)
        if lm is None:
            flow, nd = ops.reproject_disp(pd, params, True, True)
            lm = ops.softsplat(nd, flow, metric)
        else:
            stack = torch.cat([pd, lm.contiguous()], 1)[:, :local_map_size].contiguous()
            flow, nd = ops.reproject_disp(stack, params, True, True)
            lm = ops.softsplat(nd, flow, metric)
        prev_info["local_map"] = lm
        prev_info["local_map_size"] = local_map_size
    return prev_info


class TemporalWarpMixin:
    """Mix into (or monkey-patch onto) the reference LightningModule to replace its `update_map`
    (projects/TemporalStereo/TemporalStereo.py:326) without touching the trainer:

        class FastTemporalStereo(TemporalWarpMixin, TemporalStereo): pass
    """

    def update_map(self, batch, prev_info, timestamp):
        outputs = {}
        full_h, full_w = batch[("color_aug", timestamp, "l")].shape[-2:]
        T = prev_info.get("T_past_to_now", None)
        if T is None:
            T_now, inv_prev = batch[("T", timestamp, "l")], batch[("inv_T", timestamp - 1, "l")]
            outputs[("T_past_to_now_gt", timestamp, "l")] = torch.bmm(T_now, inv_prev)
        else:
            T_now = T
            inv_prev = torch.eye(4, device=T.device).repeat(T.shape[0], 1, 1)
        prev_info = update_map(prev_info, batch[("K", 0)], T_now, inv_prev, batch["baseline"], full_h, full_w,
                               use_past_cost=self.use_past_cost, local_map_size=self.local_map_size,
                               with_previous=self.with_previous)
        return outputs, prev_info
