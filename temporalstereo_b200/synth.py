"""Deterministic synthetic weights / inputs for the aggregation hot path.

The reference ships no checkpoints that are reachable offline, and its own
`weight_init()` (He-normal, reference `coarse.py:52-67`) is so ill-conditioned
that the reference disagrees with itself between fp32 and fp64.
This module is the single, numpy-seeded recipe used by the golden generator
(oracle/make_golden.py, which loads the tensors into the *real* reference
modules with strict=True), by the tests and by bench.py, so that the same
bits exist on the build container and on the GPU box without shipping weights.

numpy's RandomState (MT19937) stream is stable across numpy versions, unlike
torch's CPU generator, which is why it is used here.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Tuple

import numpy as np
import torch

# Level hyper-parameters of every shipped YAML (reference configs/sceneflow.yaml:38-60).
DEFAULT_LEVELS = {
    "coarse": dict(in_planes=256, C=32, num_sample=12),
    "fine": dict(in_planes=128, C=16, num_sample=5),
    "precise": dict(in_planes=64, C=8, num_sample=5),
}


def _sep(prefix: str, cin: int, cout: int, k: int, bias: bool, transposed: bool = False):
    """Key/shape list of one separable (1,k,k)+(k,1,1) pair (reference module.py:111-184)."""
    out = []
    w0 = (cin, cout, 1, k, k) if transposed else (cout, cin, 1, k, k)
    out.append((f"{prefix}.conv.0.weight", w0))
    if bias:
        out.append((f"{prefix}.conv.0.bias", (cout,)))
    out += _bn(f"{prefix}.conv.0.norm", cout)
    out.append((f"{prefix}.conv.1.weight", (cout, cout, k, 1, 1)))
    if bias:
        out.append((f"{prefix}.conv.1.bias", (cout,)))
    out += _bn(f"{prefix}.conv.1.norm", cout)
    return out


def _bn(prefix: str, c: int):
    return [
        (f"{prefix}.weight", (c,)),
        (f"{prefix}.bias", (c,)),
        (f"{prefix}.running_mean", (c,)),
        (f"{prefix}.running_var", (c,)),
        (f"{prefix}.num_batches_tracked", ()),
    ]


def _init3d(prefix: str, cost_planes: int, C: int):
    out = _sep(f"{prefix}.0", cost_planes, C, 3, bias=True)
    h = f"{prefix}.1"
    out += _sep(f"{h}.conv1", C, 2 * C, 3, False)
    out += _sep(f"{h}.conv2", 2 * C, 2 * C, 3, False)
    out += _sep(f"{h}.conv3", 2 * C, 2 * C, 3, False)
    out += _sep(f"{h}.conv4", 2 * C, 2 * C, 3, False)
    out += _sep(f"{h}.conv5", 2 * C, 2 * C, 3, False, transposed=True)
    out += _sep(f"{h}.conv6", 2 * C, C, 3, False, transposed=True)
    out += _sep(f"{h}.shortcut5", 2 * C, 2 * C, 3, False)
    out += _sep(f"{h}.shortcut6", C, C, 3, False)
    out += _sep(f"{prefix}.2", C, C, 3, False)
    return out


def _heads(prefix: str, C: int):
    out = []
    for head in ("cost_head", "off_head"):
        out.append((f"{prefix}.{head}.0.weight", (C, C, 3, 1, 1)))
        out += _bn(f"{prefix}.{head}.0.norm", C)
        out.append((f"{prefix}.{head}.1.weight", (1, C, 1, 3, 3)))
    return out


def _memory_level(prefix: str, in_planes: int, C: int, cost_planes: int, with_phi: bool):
    out = []
    if with_phi:
        out.append((f"{prefix}.phi", (1,)))
    out += _init3d(f"{prefix}.init3d", cost_planes, C)
    out.append((f"{prefix}.past_conv.weight", (C, 1, 1, 1, 1)))
    out += _bn(f"{prefix}.past_conv.norm", C)
    out.append((f"{prefix}.fuse.conv_5x5.weight", (C, C, 5, 1, 1)))
    out += _bn(f"{prefix}.fuse.conv_5x5.norm", C)
    out += _sep(f"{prefix}.fuse.conv_fuse", 4 * C, C, 3, False)
    out += _heads(f"{prefix}.pred_heads", C)
    m = f"{prefix}.convex_upsample.mask"
    out.append((f"{m}.0.weight", (64, in_planes, 3, 3)))
    out.append((f"{m}.0.bias", (64,)))
    out += _bn(f"{m}.1", 64)
    out.append((f"{m}.3.weight", (36, 64, 1, 1)))
    out.append((f"{m}.3.bias", (36,)))
    return out


def _unet(prefix: str, out_planes: int):
    C = 32
    out = []

    def cbn(name, cin, cout, k=3, bias=False, transposed=False):
        w = (cin, cout, k, k) if transposed else (cout, cin, k, k)
        o = [(f"{prefix}.{name}.weight", w)]
        if bias:
            o.append((f"{prefix}.{name}.bias", (cout,)))
        o += _bn(f"{prefix}.{name}.norm", cout)
        return o

    out += cbn("conv2.0", 3, C) + cbn("conv2.1", C, C)
    out += cbn("conv4.0", C, out_planes) + cbn("conv4.1", out_planes, out_planes)
    out += cbn("fuse.0", 2 * out_planes, C) + cbn("fuse.1", C, C)
    out += cbn("deconv4", C, C, k=4, bias=True, transposed=True)
    out += cbn("concat", 2 * C, C)
    out.append((f"{prefix}.deconv2.weight", (C, 9, 4, 4)))
    out.append((f"{prefix}.deconv2.bias", (9,)))
    return out


def state_dict_spec(levels: Dict[str, dict] = None, block_cost_scale: int = 3):
    """Ordered (key, shape) list of the aggregation module's state dict.

    Same 526 names and shapes as the reference TEMPORALSTEREO aggregation
    (reference aggregation/TemporalStereo/TemporalStereo.py:15-78).
    """
    lv = levels or DEFAULT_LEVELS
    c, f, p = lv["coarse"], lv["fine"], lv["precise"]
    s = block_cost_scale
    spec = []
    spec += _memory_level("coarse", c["in_planes"], c["C"], c["in_planes"] + s * c["in_planes"] // 8, False)
    spec += _memory_level("fine", f["in_planes"], f["C"], 2 * f["in_planes"] + s * f["in_planes"] // 8, True)
    pp = p["in_planes"]
    spec += _init3d("precise.init3d", 4 * pp + s * 2 * pp // 8, p["C"])
    spec += _heads("precise.pred_heads", p["C"])
    spec += _unet("precise.refinement", pp)
    return spec


def synthetic_state_dict(seed: int = 0, gain: float = 0.6, bn_jitter: float = 0.1,
                         levels: Dict[str, dict] = None) -> "OrderedDict[str, torch.Tensor]":
    """He-normal * gain conv weights, mildly randomised BN statistics.

    gain=0.6 is the conditioning SURVEY.md §8c found necessary for the reference
    to agree with itself between fp32 and fp64 to ~1e-5 px.
    """
    rng = np.random.RandomState(seed)
    sd = OrderedDict()
    for key, shape in state_dict_spec(levels):
        leaf = key.rsplit(".", 1)[1]
        if leaf == "num_batches_tracked":
            t = torch.zeros((), dtype=torch.long)
        elif key.endswith("phi"):
            t = torch.zeros(1)
        elif leaf == "running_mean":
            t = torch.from_numpy((bn_jitter * rng.standard_normal(shape)).astype(np.float32))
        elif leaf == "running_var":
            t = torch.from_numpy(rng.uniform(1.0 - bn_jitter, 1.0 + 2 * bn_jitter, shape).astype(np.float32))
        elif leaf == "weight" and len(shape) == 1:      # BN gamma
            t = torch.from_numpy(rng.uniform(1.0 - bn_jitter, 1.0 + bn_jitter, shape).astype(np.float32))
        elif leaf == "bias":                            # BN beta or conv bias
            t = torch.from_numpy((bn_jitter * rng.standard_normal(shape)).astype(np.float32))
        else:                                           # conv / deconv weight
            is_t = ("conv5.conv" in key or "conv6.conv" in key or "deconv" in key)
            cout = shape[1] if is_t else shape[0]
            n = int(np.prod(shape[2:])) * cout
            std = gain * math.sqrt(2.0 / n)
            t = torch.from_numpy((std * rng.standard_normal(shape)).astype(np.float32))
        sd[key] = t
    return sd


def synthetic_frame(H: int, W: int, B: int = 1, seed: int = 1, shift_px: int = 12,
                    chans: Tuple[int, int, int] = (64, 128, 256)):
    """Feature pyramids + images with real matches: right = left rolled by the
    disparity `shift_px` (full-res pixels) + 0.1*noise."""
    assert H % 16 == 0 and W % 16 == 0, "sizes must be multiples of 16 (SURVEY.md fact 8)"
    rng = np.random.RandomState(seed)
    lf, rf = [], []
    for c, s in zip(chans, (4, 8, 16)):
        h, w = H // s, W // s
        l = rng.standard_normal((B, c, h, w)).astype(np.float32)
        sh = max(1, int(round(shift_px / s)))
        r = np.roll(l, -sh, axis=3) + 0.1 * rng.standard_normal((B, c, h, w)).astype(np.float32)
        lf.append(torch.from_numpy(l))
        rf.append(torch.from_numpy(r.astype(np.float32)))
    li = rng.standard_normal((B, 3, H, W)).astype(np.float32)
    ri = np.roll(li, -shift_px, axis=3) + 0.1 * rng.standard_normal((B, 3, H, W)).astype(np.float32)
    return lf, rf, torch.from_numpy(li), torch.from_numpy(ri.astype(np.float32))


def synthetic_temporal_state(H: int, W: int, B: int = 1, seed: int = 2, local_map_size: int = 3):
    """prev_info + pose/intrinsics batch entries for the temporal configs."""
    rng = np.random.RandomState(seed)
    h8, w8 = H // 8, W // 8
    # disparities are tied to the image width so that depth = b*f/disp stays in a
    # KITTI-like 5..31 m band at every test size and the 0.8 m ego-motion is physical
    prev_disp = rng.uniform(0.01 * W, 0.05 * W, (B, 1, H, W)).astype(np.float32)
    mem_sample = rng.uniform(0.01 * w8, 0.05 * w8, (B, 2, h8, w8)).astype(np.float32)
    mem_cost = rng.standard_normal((B, 2, h8, w8)).astype(np.float32)
    local_map = rng.uniform(0.01 * w8, 0.05 * w8, (B, local_map_size, h8, w8)).astype(np.float32)
    # KITTI-like intrinsics scaled to (H, W) (reference data/datasets/kitti/base.py:14-21)
    K = np.eye(4, dtype=np.float32)
    K[0, 0], K[0, 2], K[1, 1], K[1, 2] = 0.58 * W, 0.5 * W, 1.92 * H, 0.5 * H
    K = np.tile(K[None], (B, 1, 1))
    T = np.tile(np.eye(4, dtype=np.float32)[None], (B, 1, 1))
    T[:, 2, 3] = -0.8
    T[:, 0, 3] = 0.05
    eye = np.tile(np.eye(4, dtype=np.float32)[None], (B, 1, 1))
    return dict(
        prev_disp=torch.from_numpy(prev_disp),
        cost_memory=dict(disp_sample=torch.from_numpy(mem_sample), cost_volume=torch.from_numpy(mem_cost)),
        local_map=torch.from_numpy(local_map),
        K=torch.from_numpy(K), T_now=torch.from_numpy(T), inv_T_prev=torch.from_numpy(eye),
        baseline=torch.full((B, 1, 1, 1), 0.54),
    )
