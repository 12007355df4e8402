"""CPU ORACLE — test infrastructure, not product code.

A functional, state-dict-driven restatement (plain PyTorch CPU fp32 ops) of the
reference hot path: cost-volume construction -> 3-level separable 3-D aggregation
-> top-2 soft-argmin regression -> temporal warp.  Every function cites the
reference file:line it follows (paths relative to /root/reference).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this file, and only as the checker.  The product
package (temporalstereo_b200/) never imports it and has no CPU fallback.

Parity pinning: the reference has no tests or golden vectors of its own
(SURVEY.md §4), so this oracle is pinned against outputs of the *real* reference
code imported in the build container: oracle/make_golden.py writes
tests/golden/*.npz and tests/test_oracle_golden.py replays them here.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]
BN_EPS = 1e-5


# --------------------------------------------------------------------------- cost volume
def gwc_neg_sq(a: torch.Tensor, b: torch.Tensor, group: int = 8) -> torch.Tensor:
    """-(a-b)^2 summed over groups of 8 consecutive channels (block_cost.py:6-13)."""
    B, C, D, H, W = a.shape
    d = (a - b) ** 2
    return -d.view(B, C // group, group, D, H, W).sum(2)


def shift_right_volume(right: torch.Tensor, D: int) -> torch.Tensor:
    """R_d[..., x] = R[..., x-d], zero where x<d, d=0..D-1 (block_cost.py:34-41)."""
    B, C, H, W = right.shape
    vol = right.new_zeros(B, C, D, H, W)
    for d in range(min(D, W)):
        vol[:, :, d, :, d:] = right[:, :, :, : W - d]
    return vol


def warp_right_volume(right: torch.Tensor, samples: torch.Tensor) -> torch.Tensor:
    """Horizontal warp of R to x - sample per pixel/candidate (block_cost.py:50-56 calling
    inverse_warp_3d.py:4-58 with disp=-samples): normalise to [-1,1], 5-D grid_sample
    (trilinear, zeros padding, align_corners=True)."""
    B, S, H, W = samples.shape
    C = right.shape[1]
    vol = right.unsqueeze(2).expand(B, C, S, H, W)
    dev = samples.device
    gd = torch.linspace(0, S - 1, S).view(1, S, 1, 1).expand(B, S, H, W).to(dev)
    gh = torch.linspace(0, H - 1, H).view(1, 1, H, 1).expand(B, S, H, W).to(dev)
    gw = torch.linspace(0, W - 1, W).view(1, 1, 1, W).expand(B, S, H, W).to(dev)
    gw = gw + (-samples)
    gd = (gd / (S - 1) * 2) - 1
    gh = (gh / (H - 1) * 2) - 1
    gw = (gw / (W - 1) * 2) - 1
    grid = torch.stack((gw, gh, gd), dim=4)
    return F.grid_sample(vol, grid, padding_mode="zeros", align_corners=True)


def multiscale_group_terms(lvol: torch.Tensor, rvol: torch.Tensor, scales: int):
    """Three pooled group-wise terms, resized back trilinearly (block_cost.py:64-78)."""
    B, C, D, H, W = lvol.shape
    out = []
    for s in range(int(scales)):
        kh, kw = min(2 ** s, H), min(2 ** s, W)
        lp = F.avg_pool3d(lvol, (1, kh, kw), (1, kh, kw))
        rp = F.avg_pool3d(rvol, (1, kh, kw), (1, kh, kw))
        g = gwc_neg_sq(lp, rp)
        g = F.interpolate(g, size=(D, H, W), mode="trilinear", align_corners=True)
        out.append(g)
    return out


def block_cost(left: torch.Tensor, right: torch.Tensor, disp_sample, block_cost_scale: int = 3):
    """Drop-in restatement of block_cost (block_cost.py:16-83).
    int disp_sample  -> [-(L-R_d)^2, g0, g1, g2]      (coarse)
    tensor           -> [L bcast, R_warp, g0, g1, g2] (fine / precise)"""
    B, C, H, W = left.shape
    if isinstance(disp_sample, int):
        D = disp_sample
        rvol = shift_right_volume(right, D)
        lvol = left.unsqueeze(2).expand(B, C, D, H, W)
        first = -((lvol - rvol) ** 2)
    else:
        D = disp_sample.shape[1]
        rvol = warp_right_volume(right, disp_sample)
        lvol = left.unsqueeze(2).expand(B, C, D, H, W)
        first = torch.cat([lvol, rvol], 1)
    return torch.cat([first] + multiscale_group_terms(lvol, rvol, block_cost_scale), 1)


# --------------------------------------------------------------------------- conv blocks
def _bn(x: torch.Tensor, sd: SD, p: str) -> torch.Tensor:
    """Eval-mode batch norm with running statistics (basic_layers.py:10-39, 230-231)."""
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"],
                        sd[p + ".weight"], sd[p + ".bias"], False, 0.0, BN_EPS)


def _act(x: torch.Tensor, act: Optional[str]) -> torch.Tensor:
    if act is None:
        return x
    return {"SiLU": F.silu, "ReLU": F.relu}[act](x)


def conv3d_bn_act(x, sd: SD, p: str, stride=(1, 1, 1), padding=(0, 0, 0), dilation=(1, 1, 1),
                  bn: bool = True, act: Optional[str] = "SiLU"):
    """Conv3d wrapper: conv -> norm -> activation (basic_layers.py:194-235)."""
    y = F.conv3d(x, sd[p + ".weight"], sd.get(p + ".bias"), stride, padding, dilation)
    if bn:
        y = _bn(y, sd, p + ".norm")
    return _act(y, act)


def deconv3d_bn_act(x, sd: SD, p: str, stride, padding, output_padding, act: Optional[str]):
    """ConvTranspose3d wrapper (basic_layers.py:340-388)."""
    y = F.conv_transpose3d(x, sd[p + ".weight"], sd.get(p + ".bias"), stride, padding, output_padding)
    return _act(_bn(y, sd, p + ".norm"), act)


def sep_conv3d(x, sd: SD, p: str, stride: int = 1, padding: int = 1, dilation: int = 1,
               act: Optional[str] = "SiLU"):
    """'DepthwiseConv3D' = (1,k,k) then (k,1,1), both full channel mixing (module.py:111-147)."""
    y = conv3d_bn_act(x, sd, p + ".conv.0", (1, stride, stride), (0, padding, padding), (1, dilation, dilation), act=act)
    return conv3d_bn_act(y, sd, p + ".conv.1", (stride, 1, 1), (padding, 0, 0), (dilation, 1, 1), act=act)


def sep_deconv3d(x, sd: SD, p: str, act: Optional[str] = None):
    """'DepthwiseConvTranspose3D' k=3 s=2 p=1 op=1 (module.py:149-184, 239-249)."""
    y = deconv3d_bn_act(x, sd, p + ".conv.0", (1, 2, 2), (0, 1, 1), (0, 1, 1), act)
    return deconv3d_bn_act(y, sd, p + ".conv.1", (2, 1, 1), (1, 0, 0), (1, 0, 0), act)


def hourglass3d(x, sd: SD, p: str):
    """ResidualBlock3D forward (module.py:271-297)."""
    o = sep_conv3d(x, sd, p + ".conv1", stride=2)
    pre = sep_conv3d(o, sd, p + ".conv2")
    o = sep_conv3d(pre, sd, p + ".conv3", stride=2)
    o = F.silu(sep_conv3d(o, sd, p + ".conv4", act=None))
    o = sep_deconv3d(o, sd, p + ".conv5")
    o = F.interpolate(o, size=pre.shape[-3:], mode="trilinear", align_corners=True)
    o = F.silu(o + sep_conv3d(pre, sd, p + ".shortcut5", act=None))
    o = sep_deconv3d(o, sd, p + ".conv6")
    o = F.interpolate(o, size=x.shape[-3:], mode="trilinear", align_corners=True)
    return F.silu(o + sep_conv3d(x, sd, p + ".shortcut6", act=None))


def init3d(raw_cost, sd: SD, p: str):
    """sep-conv(bias) -> hourglass -> dilated sep-conv (coarse.py:36-40, fine.py:36-40, precise.py:32-36)."""
    y = sep_conv3d(raw_cost, sd, p + ".0")
    y = hourglass3d(y, sd, p + ".1")
    return sep_conv3d(y, sd, p + ".2", padding=2, dilation=2)


def pyramid_fusion(cost, sd: SD, p: str):
    """cat[x, conv(5,1,1), avgpool 5^3, maxpool 5^3] -> sep-conv 4C->C no act (module.py:401-421)."""
    parts = [
        cost,
        conv3d_bn_act(cost, sd, p + ".conv_5x5", padding=(2, 0, 0)),
        F.avg_pool3d(cost, 5, 1, 2),
        F.max_pool3d(cost, 5, 1, 2),
    ]
    return sep_conv3d(torch.cat(parts, 1), sd, p + ".conv_fuse", act=None)


def prediction_heads(vol, sd: SD, p: str, delta: float):
    """cost / offset heads; off = tanh(x/100).clamp(-1,1)*delta (module.py:356-398)."""
    def head(name):
        h = conv3d_bn_act(vol, sd, f"{p}.{name}.0", padding=(1, 0, 0))
        return conv3d_bn_act(h, sd, f"{p}.{name}.1", padding=(0, 1, 1), bn=False, act=None).squeeze(1)
    off = torch.tanh(head("off_head") / 100).clamp(-1, 1) * delta
    return head("cost_head"), off


def predict_disp(cost, disp_sample, off, k: int = 2):
    """top-k -> softmax over k -> gather(sample+off) -> weighted sum (coarse.py:69-75)."""
    top_cost, idx = torch.topk(cost, k=k, dim=1)
    prob = torch.softmax(top_cost, 1)
    top_disp = torch.gather(disp_sample + off, 1, idx)
    return (prob * top_disp).sum(1, keepdim=True), top_disp, top_cost


def convex_upsample(feat, disp, sd: SD, p: str, up: int = 2, win: int = 3):
    """Learned 3x3 convex combination + pixel shuffle x2 (module.py:300-353)."""
    B, C, H, W = disp.shape
    m = F.conv2d(feat, sd[p + ".mask.0.weight"], sd[p + ".mask.0.bias"], 1, 1)
    m = F.silu(_bn(m, sd, p + ".mask.1"))
    m = F.conv2d(m, sd[p + ".mask.3.weight"], sd[p + ".mask.3.bias"])
    m = torch.softmax(m.view(B, 1, win * win, up, up, H, W), 2)
    u = F.unfold(disp * up, (win, win), padding=win // 2).view(B, C, win * win, 1, 1, H, W)
    u = (m * u).sum(2)
    return u.permute(0, 1, 4, 2, 5, 3).reshape(B, C, H * up, W * up)


def merge_memory(vol, disp_sample, sd: SD, p: str, prev_info: dict, topk: int, coarse: bool):
    """Temporal memory merge: past_conv(memory|zeros) appended on D, samples sorted, volume
    planes gathered by the sort permutation (coarse.py:84-105, fine.py:104-122)."""
    B, C, D, H, W = vol.shape
    memory = prev_info.get("cost_memory", None)
    if memory is None or not prev_info.get("use_past_cost", False):
        ms = torch.zeros_like(disp_sample[:, :topk])
        mv = torch.zeros_like(ms).unsqueeze(1)
    else:
        ms, mv = memory["disp_sample"], memory["cost_volume"]
        if coarse:
            mw = ms.shape[-1]
            ms = F.interpolate(ms * W / mw, size=(H, W), mode="bilinear", align_corners=True)
            mv = F.interpolate(mv, size=(H, W), mode="bilinear", align_corners=True)
        mv = mv.unsqueeze(1)
    mv = conv3d_bn_act(mv, sd, p + ".past_conv")
    samples = torch.cat([disp_sample, ms], 1)
    vol = torch.cat([vol, mv], 2)
    samples, order = torch.sort(samples, dim=1)
    vol = torch.gather(vol, 2, order.unsqueeze(1).repeat(1, C, 1, 1, 1)).contiguous()
    return vol, samples


def range_samples(low, high):
    """min(low,high) + |high-low| * {0,3,4,5,8}/8 (fine.py:78-86, precise.py:69-79)."""
    frac = (torch.tensor([0.0, 3.0, 4.0, 5.0, 8.0]) / 8.0).view(1, 5, 1, 1).to(low.device)
    return torch.abs(high - low) * frac + torch.min(low, high)


# --------------------------------------------------------------------------- levels
def coarse_level(left, right, sd: SD, prev_info: dict, num_sample=12, delta=1.0, scales=3, topk=2):
    """CoarseAggregation.forward (coarse.py:77-116)."""
    B, _, H, W = left.shape
    raw = block_cost(left, right, int(num_sample), scales)
    samples = torch.linspace(0, num_sample - 1, num_sample).view(1, num_sample, 1, 1).expand(B, num_sample, H, W).to(left.device)
    vol = init3d(raw, sd, "coarse.init3d")
    vol, samples = merge_memory(vol, samples, sd, "coarse", prev_info, topk, coarse=True)
    vol = pyramid_fusion(vol, sd, "coarse.fuse")
    cost, off = prediction_heads(vol, sd, "coarse.pred_heads", delta)
    disp, _, _ = predict_disp(cost, samples, off, topk)
    return convex_upsample(left, disp, sd, "coarse.convex_upsample"), cost, off, samples


def fine_level(left, right, low, high, sd: SD, prev_info: dict, delta=1.0, scales=3, topk=2):
    """FineAggregation.forward (fine.py:97-132), local-map candidates first (fine.py:88-93)."""
    B, _, H, W = left.shape
    samples = range_samples(low, high)
    lm = prev_info.get("local_map", None)
    if lm is not None and prev_info.get("local_map_size", 0) > 0:
        lm = F.interpolate(lm * W / lm.shape[-1], size=(H, W), mode="bilinear", align_corners=True)
        samples = torch.cat([lm, samples], 1)
    raw = block_cost(left, right, samples, scales)
    vol = init3d(raw, sd, "fine.init3d")
    vol, samples = merge_memory(vol, samples, sd, "fine", prev_info, topk, coarse=False)
    vol = pyramid_fusion(vol, sd, "fine.fuse")
    cost, off = prediction_heads(vol, sd, "fine.pred_heads", delta)
    disp, _, _ = predict_disp(cost, samples, off, topk)
    return convex_upsample(left, disp, sd, "fine.convex_upsample"), cost, off, samples


def _c2d(x, sd: SD, p: str, stride=1, act="ReLU"):
    y = F.conv2d(x, sd[p + ".weight"], sd.get(p + ".bias"), stride, 1)
    return _act(_bn(y, sd, p + ".norm"), act)


def unet_encoder(img, sd: SD, p: str):
    """UNet.encoder for one image (module.py:459-466)."""
    s2 = _c2d(_c2d(img, sd, p + ".conv2.0", 2), sd, p + ".conv2.1")
    s4 = _c2d(_c2d(s2, sd, p + ".conv4.0", 2), sd, p + ".conv4.1")
    return s2, s4


def unet_upsample(mask, disp):
    """softmax-9 x bilinear-x4 of the 3x3 unfolded, width-rescaled disparity (module.py:468-483)."""
    mask = F.softmax(mask, 1)
    b, _, h, w = mask.shape
    dh, dw = disp.shape[-2:]
    u = F.unfold(disp, (3, 3), padding=1).reshape(b, 9, dh, dw)
    full = F.interpolate(u * w / dw, size=(h, w), mode="bilinear", align_corners=True)
    return (full * mask).sum(1, keepdim=True)


def unet_decoder(disp, feat, feat2x, sd: SD, p: str):
    """UNet.decoder (module.py:485-492)."""
    f = _c2d(_c2d(feat, sd, p + ".fuse.0"), sd, p + ".fuse.1")
    f = F.conv_transpose2d(f, sd[p + ".deconv4.weight"], sd[p + ".deconv4.bias"], 2, 1)
    f = F.relu(_bn(f, sd, p + ".deconv4.norm"))
    f = _c2d(torch.cat([f, feat2x], 1), sd, p + ".concat")
    mask = F.conv_transpose2d(f, sd[p + ".deconv2.weight"], sd[p + ".deconv2.bias"], 2, 1)
    return unet_upsample(mask, disp)


def precise_level(left, right, low, high, limg, rimg, sd: SD, prev_info: dict, delta=1.0, scales=3, topk=2):
    """PreciseAggregation.forward incl. state write-back (precise.py:81-105)."""
    p = "precise.refinement"
    s2l, s4l = unet_encoder(limg, sd, p)
    _, s4r = unet_encoder(rimg, sd, p)
    left, right = torch.cat([left, s4l], 1), torch.cat([right, s4r], 1)
    samples = range_samples(low, high)
    raw = block_cost(left, right, samples, scales)
    vol = init3d(raw, sd, "precise.init3d")
    cost, off = prediction_heads(vol, sd, "precise.pred_heads", delta)
    disp, top_disp, top_cost = predict_disp(cost, samples, off, topk)
    full = unet_decoder(disp, left, s2l, sd, p)
    prev_info["prev_disp"] = full.detach()
    prev_info["cost_memory"] = {
        "disp_sample": F.interpolate(top_disp / 2, scale_factor=1 / 2, mode="bilinear", align_corners=True),
        "cost_volume": F.interpolate(top_cost, scale_factor=1 / 2, mode="bilinear", align_corners=True),
    }
    return full, disp, cost, off, samples


def aggregation_forward(sd: SD, left_feats, right_feats, left_image, right_image, prev_info: dict,
                        num_sample: int = 12, disp_range: float = 4.0):
    """TEMPORALSTEREO.forward: the 6-tuple of the drop-in boundary
    (aggregation/TemporalStereo/TemporalStereo.py:97-135)."""
    l4, l8, l16 = left_feats
    r4, r8, r16 = right_feats
    d_c, c_c, o_c, s_c = coarse_level(l16, r16, sd, prev_info, num_sample)
    low, high = d_c - disp_range, d_c + disp_range
    rng_c = {"low": low, "high": high}
    d_f, c_f, o_f, s_f = fine_level(l8, r8, low, high, sd, prev_info)
    low, high = d_f - disp_range, d_f + disp_range
    rng_f = {"low": low, "high": high}
    full, d_p, c_p, o_p, s_p = precise_level(l4, r4, low, high, left_image, right_image, sd, prev_info)
    return ([full, d_p, d_f, d_c], [c_p, c_f, c_c], [s_p, s_f, s_c], [o_p, o_f, o_c], [rng_f, rng_c], prev_info)


def upsample_disps(disps, full_h: int, full_w: int):
    """Meta-arch post-processing: bilinear to full-res, values x full_w/dw
    (projects/TemporalStereo/TemporalStereo.py:305-309)."""
    return [F.interpolate(d * full_w / d.shape[-1], size=(full_h, full_w), mode="bilinear", align_corners=True) for d in disps]


# --------------------------------------------------------------------------- temporal warp
def softsplat_softmax(x, flow, metric):
    """CPU restatement of FunctionSoftsplat(..., 'softmax') + kernel_Softsplat_updateOutput
    (softsplat.py:334-360 packing/normalisation, :14-52 corner indices, weights, bounds)."""
    B, C, H, W = x.shape
    e = metric.exp()
    src = torch.cat([x * e, e], 1)                                   # softsplat.py:344-345
    ys, xs = torch.meshgrid(torch.arange(H, dtype=x.dtype), torch.arange(W, dtype=x.dtype), indexing="ij")
    fx = xs.unsqueeze(0) + flow[:, 0]
    fy = ys.unsqueeze(0) + flow[:, 1]
    x0 = torch.floor(fx).long()
    y0 = torch.floor(fy).long()
    out = src.new_zeros(B, C + 1, H * W)
    flat = src.reshape(B, C + 1, H * W)
    for dx, dy in ((0, 0), (1, 0), (0, 1), (1, 1)):                  # NW, NE, SW, SE
        xi, yi = x0 + dx, y0 + dy
        # weight of a corner = area to the opposite corner (softsplat.py:33-36)
        wx = (x0 + 1).to(x.dtype) - fx if dx == 0 else fx - x0.to(x.dtype)
        wy = (y0 + 1).to(x.dtype) - fy if dy == 0 else fy - y0.to(x.dtype)
        w = (wx * wy).reshape(B, 1, H * W)
        ok = ((xi >= 0) & (xi < W) & (yi >= 0) & (yi < H)).reshape(B, H * W)
        tgt = (yi.clamp(0, H - 1) * W + xi.clamp(0, W - 1)).reshape(B, H * W)
        for b in range(B):
            sel = ok[b]
            out[b].index_add_(1, tgt[b][sel], (flat[b] * w[b])[:, sel])
    out = out.view(B, C + 1, H, W)
    return out[:, :-1] / (out[:, -1:] + 1e-22)                       # softsplat.py:352-356


def project_to_3d(depth, K, inv_K, T, eps: float = 1e-7):
    """Back-project C depth maps, move by T, re-project with K (inverse_warp.py:92-178).
    Returns (optical_flow [B,2C,H,W], triangular_depth [B,C,H,W])."""
    B, C, H, W = depth.shape
    ys, xs = torch.meshgrid(torch.arange(H, dtype=depth.dtype), torch.arange(W, dtype=depth.dtype), indexing="ij")
    pix = torch.stack([xs, ys], 0).unsqueeze(0).expand(B, 2, H, W)
    homo = torch.cat([pix, torch.ones(B, 1, H, W, dtype=depth.dtype)], 1).reshape(B, 3, -1).repeat(1, 1, C)
    pts = torch.matmul(inv_K[:, :3, :3], homo) * depth.reshape(B, 1, -1)
    pts = torch.cat([pts, torch.ones(B, 1, C * H * W, dtype=depth.dtype)], 1)
    P = torch.matmul(K, T)[:, :3, :]
    src = torch.matmul(P, pts)
    tri = src[:, -1, :].reshape(B, C, H, W)
    uv = src[:, :2, :] / (src[:, 2:3, :] + eps)
    uv = uv.reshape(B, 2, C, H, W).permute(0, 2, 1, 3, 4).reshape(B, 2 * C, H, W)
    return uv - pix.repeat(1, C, 1, 1), tri


def _scaled_intrinsics(K, factor: float):
    """Rows 0 and 1 of the 4x4 K divided by the *width* ratio (TemporalStereo.py:347-353, 393-400)."""
    down = torch.cat([K[:, 0:1] / factor, K[:, 1:2] / factor, K[:, 2:]], 1)
    return down, torch.inverse(down), down[:, 0, 0].view(-1, 1, 1, 1)


EXPMAX = 50.0


def update_map(prev_info: dict, K, T_now, inv_T_prev, baseline, full_h: int, full_w: int,
               use_past_cost: bool = True, local_map_size: int = 3):
    """Pose-conditioned temporal warp of the recurrent state
    (projects/TemporalStereo/TemporalStereo.py:326-461)."""
    T = torch.bmm(T_now, inv_T_prev)                                          # :333-338
    prev_disp_full = prev_info["prev_disp"].detach()
    baseline = baseline.float()

    def flow_of(prev_disp, h, w):
        dK, dKi, f = _scaled_intrinsics(K, full_w / w)
        pd = F.interpolate(prev_disp * w / prev_disp.shape[-1], size=(h, w), mode="bilinear", align_corners=True)
        flow, tri = project_to_3d(baseline * f / (pd + 1e-5), dK, dKi, T)
        metric = (pd[:, :1] - pd[:, :1].mean()).clamp(-EXPMAX, EXPMAX)
        return dK, dKi, f, pd, flow[:, :2], tri, metric

    memory = prev_info.get("cost_memory", None)
    if use_past_cost and memory is not None:                                  # :386-426
        ds, cv = memory["disp_sample"].detach(), memory["cost_volume"].detach()
        c, h, w = ds.shape[1:]
        dK, dKi, f, pd, flow, _, metric = flow_of(prev_disp_full, h, w)
        _, tri = project_to_3d(baseline * f / (ds + 1e-5), dK, dKi, T)
        new_ds = baseline * f / (tri + 1e-5)
        warped = softsplat_softmax(torch.cat([new_ds, cv], 1), flow, metric)
        memory = {"disp_sample": warped[:, :c], "cost_volume": warped[:, c:]}
    elif not use_past_cost:
        memory = None
    prev_info["cost_memory"] = memory
    prev_info["use_past_cost"] = use_past_cost

    if local_map_size > 0:                                                    # :340-384
        lm = prev_info.get("local_map", None)
        h, w = (lm.shape[-2:] if lm is not None else (full_h // 8, full_w // 8))
        dK, dKi, f, pd, flow, tri, metric = flow_of(prev_disp_full, h, w)
        warp_disp = softsplat_softmax(baseline * f / (tri + 1e-5), flow, metric)
        if lm is None:
            lm = warp_disp
        else:
            lm = torch.cat([pd, lm], 1)[:, :local_map_size]
            lflow, ltri = project_to_3d(baseline * f / (lm + 1e-5), dK, dKi, T)
            lm = softsplat_softmax(baseline * f / (ltri + 1e-5), lflow[:, :2], metric)
        prev_info["local_map"] = lm.detach()
        prev_info["local_map_size"] = local_map_size
    return prev_info


# --------------------------------------------------------------------------- training losses, forward
def _scaled_gt(gt, H: int, W: int, max_disp: float, start_disp: float, sparse: bool):
    """gt / scale, adaptive pooling onto the level's grid, validity mask (smooth_l1_loss.py:50-63 ==
    warsserstein_distance_loss.py:57-69)."""
    scale = 1.0
    g = gt
    if gt.shape[-2] != H or gt.shape[-1] != W:
        scale = gt.shape[-1] / (W * 1.0)
        g = (F.adaptive_max_pool2d if sparse else F.adaptive_avg_pool2d)(gt / scale, (H, W))
    mask = (g > start_disp) & (g < (max_disp / scale))
    return g, mask


def smooth_l1_loss_level(est, gt, max_disp=192, start_disp=0, sparse=False):
    """DispSmoothL1Loss.loss_per_level (losses/smooth_l1_loss.py:49-74)."""
    g, mask = _scaled_gt(gt, est.shape[-2], est.shape[-1], max_disp, start_disp, sparse)
    if mask.sum() < 1.0:
        return (torch.abs(est - g) * mask.float()).mean()
    return F.smooth_l1_loss(est[mask], g[mask], reduction="mean")


def wasserstein_loss_level(cost, off, sample, gt, max_disp=192, start_disp=0, sparse=False):
    """WarssersteinDistanceLoss.loss_per_level (losses/warsserstein_distance_loss.py:53-81)."""
    prob = torch.softmax(cost, dim=1)
    g, mask = _scaled_gt(gt, cost.shape[-2], cost.shape[-1], max_disp, start_disp, sparse)
    if mask.sum() < 1.0:
        return (prob * torch.abs(off + sample - g) * mask.float()).sum(dim=1).mean()
    return ((prob * 1.0 + 0.25) * torch.abs(off + sample - g) * mask.float()).sum(dim=1).mean()

