"""Drop-in `TEMPORALSTEREO` aggregation module backed by libtstereo.so.

Mirrors the reference operator boundary:
  * registered as `TEMPORALSTEREO` in `AGGREGATION_REGISTRY`
    (reference architecture/modeling/aggregation/builder.py:3-20),
  * built either from explicit level dicts or from a config through `from_config`, reading the same
    `MODEL.AGGREGATION.{COARSE,FINE,PRECISE}.*` keys
    (reference aggregation/TemporalStereo/TemporalStereo.py:38-78),
  * exposes the same 526 state-dict entries (names + shapes), so reference checkpoints load with
    strict=True (reference projects/TemporalStereo/demo.py:250-251),
  * `forward(left_feats, right_feats, left_image, right_image, prev_info)` returns the same 6-tuple
    and updates `prev_info` in place (reference aggregation/TemporalStereo/TemporalStereo.py:97-135).

All arithmetic runs in hand-written sm_100a kernels through the C ABI; this file only folds
eval-mode BatchNorm into packed weights and sequences the launches on torch's current stream.
Inference only (the reference's history frames also run under no_grad, TemporalStereo.py:268-274).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import ops
from .registry import AGGREGATION_REGISTRY
from .synth import DEFAULT_LEVELS, state_dict_spec

BN_EPS = 1e-5
DISP_RANGE = 4.0          # reference aggregation/TemporalStereo/TemporalStereo.py:103


class _Node(nn.Module):
    """Anonymous container used to reproduce the reference's parameter tree."""


class _Packed:
    __slots__ = ("w", "b", "cout", "wtc", "wtc2", "lazy")

    def __init__(self, w, b, cout, wtc=None, wtc2=None):
        self.w, self.b, self.cout = w, b, cout
        self.lazy = {}          # operand images built on first use (stride-2 / transposed tensor-core forms)
        self.wtc = wtc          # tcgen05 (3xTF32) operand image of a 3x3 / (k,1,1) conv, or None
        self.wtc2 = wtc2        # operand image of the kx-folded tcgen05 3x3 kernel (Cout <= 32), or None


def _level_cfg(node, defaults: dict) -> dict:
    get = node.get if hasattr(node, "get") else (lambda k, d: getattr(node, k, d))
    return dict(
        in_planes=get("IN_PLANES", defaults["in_planes"]),
        C=get("C", defaults["C"]),
        num_sample=get("NUM_SAMPLE", defaults["num_sample"]),
        delta=get("DELTA", 1),
        block_cost_scale=get("BLOCK_COST_SCALE", 3),
        topk=get("TOPK", 2),
        spatial_fusion=get("SPATIAL_FUSION", True),
        norm=get("NORM", "BN3d"),
        activation=get("ACTIVATION", "SiLU"),
    )


@AGGREGATION_REGISTRY.register()
class TEMPORALSTEREO(nn.Module):
    """B200 engine behind the reference's `TEMPORALSTEREO` aggregation interface."""

    def __init__(self, cfg=None, *, coarse: Optional[dict] = None, fine: Optional[dict] = None,
                 precise: Optional[dict] = None, norm: str = "BN", activation: str = "SiLU"):
        super().__init__()
        if cfg is not None:
            kw = self.from_config(cfg)
            coarse, fine, precise = kw["coarse"], kw["fine"], kw["precise"]
            norm, activation = kw["norm"], kw["activation"]
        self.levels = {}
        for name, given in (("coarse", coarse), ("fine", fine), ("precise", precise)):
            lv = dict(DEFAULT_LEVELS[name], delta=1, block_cost_scale=3, topk=2, spatial_fusion=True,
                      norm="BN3d", activation="SiLU")
            lv.update(given or {})
            if lv["block_cost_scale"] != 3 or lv["topk"] != 2:
                raise NotImplementedError("libtstereo implements BLOCK_COST_SCALE=3, TOPK=2 (every shipped config)")
            if lv["norm"] not in ("BN3d", "BN") or lv["activation"] != "SiLU":
                raise NotImplementedError("libtstereo implements NORM=BN3d, ACTIVATION=SiLU (every shipped config)")
            if name != "precise" and not lv["spatial_fusion"]:
                raise NotImplementedError("libtstereo implements SPATIAL_FUSION=True (every shipped config)")
            self.levels[name] = lv
        self.norm, self.activation = norm, activation
        self._build_tree()
        self.weight_init()
        self._pk: Optional[Dict[str, _Packed]] = None
        self._const: Dict[tuple, torch.Tensor] = {}
        # every contraction (3x3 stride 1 / 2, (k,1,1) along D, stride-2 transposed) on tcgen05 with hi+lo split operands
        # (fp32-equivalent results); False keeps them all on the fp32 FMA pipe
        self.tensor_cores = True
        # per-(layer shape) choice between the tensor-core and the fp32-FMA kernel: "auto" times both on the
        # first call of a shape (outside CUDA-graph capture) and keeps the faster; "tc2" / "simt" force one
        # (or a dict per operator kind: "hw3", "hw3s2", "d", "dc")
        self.plan_mode = "auto"
        self._plan: Dict[tuple, str] = {}
        # tensor-core operand split: fp16 hi + lo (kind::f16, 16 channels per MMA; activations < 65504) or tf32 hi + lo
        self.half_split = True
        # run the UNet encoder on a side stream, concurrently with the coarse and fine levels
        self.overlap_encoder = True
        self._side: Dict[str, torch.cuda.Stream] = {}
        self.register_load_state_dict_post_hook(lambda m, _k: m.invalidate())
        super().train(False)

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_config(cls, cfg) -> dict:
        """Same keys and defaults as the reference `from_config`
        (aggregation/TemporalStereo/TemporalStereo.py:38-78)."""
        agg = cfg["MODEL"]["AGGREGATION"] if isinstance(cfg, dict) else cfg.MODEL.AGGREGATION
        sub = (lambda k: agg[k]) if isinstance(agg, dict) else (lambda k: getattr(agg, k))
        get = agg.get if hasattr(agg, "get") else (lambda k, d: getattr(agg, k, d))
        return {
            "coarse": _level_cfg(sub("COARSE"), dict(in_planes=192, C=32, num_sample=12)),
            "fine": _level_cfg(sub("FINE"), dict(in_planes=64, C=16, num_sample=5)),
            "precise": _level_cfg(sub("PRECISE"), dict(in_planes=48, C=8, num_sample=5)),
            "norm": get("NORM", "BN"),
            "activation": get("ACTIVATION", "SiLU"),
        }

    def _build_tree(self) -> None:
        spec = state_dict_spec({k: dict(in_planes=v["in_planes"], C=v["C"], num_sample=v["num_sample"])
                                for k, v in self.levels.items()})
        for key, shape in spec:
            *path, leaf = key.split(".")
            node = self
            for part in path:
                if part not in node._modules:
                    node.add_module(part, _Node())
                node = node._modules[part]
            if leaf in ("running_mean", "running_var"):
                node.register_buffer(leaf, torch.zeros(shape) if leaf == "running_mean" else torch.ones(shape))
            elif leaf == "num_batches_tracked":
                node.register_buffer(leaf, torch.zeros((), dtype=torch.long))
            else:
                node.register_parameter(leaf, nn.Parameter(torch.zeros(shape), requires_grad=(leaf != "phi")))

    def weight_init(self) -> None:
        """He-normal over k*Cout for conv weights, BN at identity (reference coarse.py:52-67)."""
        with torch.no_grad():
            for key, p in self.named_parameters():
                leaf = key.rsplit(".", 1)[1]
                if leaf == "weight" and p.dim() > 1:
                    # the reference computes n from out_channels, which for ConvTranspose is dim 1
                    transposed = any(t in key for t in ("conv5.conv", "conv6.conv", "deconv"))
                    cout = p.shape[1] if transposed else p.shape[0]
                    n = int(math.prod(p.shape[2:])) * cout
                    p.normal_(0, math.sqrt(2.0 / n))
                elif leaf == "weight":
                    p.fill_(1.0)
                else:
                    p.zero_()
        self.invalidate()

    def invalidate(self) -> None:
        """Drop the packed (BN-folded) weights; they are rebuilt on the next forward."""
        self._pk = None

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._pk = None
        self._const = {}
        return out

    def train(self, mode: bool = True):
        if mode:
            raise NotImplementedError("libtstereo is an inference engine: BatchNorm is folded with running statistics; "
                                      "call .eval() (history frames in the reference run the same way, "
                                      "projects/TemporalStereo/TemporalStereo.py:268-274)")
        return super().train(False)

    # ------------------------------------------------------------------ weight packing
    def _fold(self, sd, conv: str, bn: Optional[str], transposed: bool = False) -> _Packed:
        """[Cout,Cin,*k] (or [Cin,Cout,*k]) -> [Cin][taps][CoutP] with eval-mode BN folded in."""
        w = sd[conv + ".weight"].detach().float()
        if transposed:
            w = w.transpose(0, 1)
        cout, cin = w.shape[:2]
        w = w.reshape(cout, cin, -1)
        bias = sd.get(conv + ".bias")
        bias = bias.detach().float() if bias is not None else None
        if bn is not None:
            s = sd[bn + ".weight"].detach().float() / torch.sqrt(sd[bn + ".running_var"].detach().float() + BN_EPS)
            w = w * s.view(-1, 1, 1)
            b0 = bias if bias is not None else torch.zeros_like(s)
            bias = (b0 - sd[bn + ".running_mean"].detach().float()) * s + sd[bn + ".bias"].detach().float()
        coutp = (cout + 3) // 4 * 4
        packed = torch.zeros((cin, w.shape[2], coutp), device=w.device, dtype=torch.float32)
        packed[:, :, :cout] = w.permute(1, 2, 0)
        wtc = None
        is_hw = conv.endswith(".conv.0") or ".refinement." in conv or ".mask." in conv     # (1,k,k) / 2-D kernels
        tc_ok = (w.shape[2] == 9 and not transposed) if is_hw else w.shape[2] in (3, 5)
        wtc2 = None
        if tc_ok and cout <= 64 and cin >= 8 and w.is_cuda and self.tensor_cores:
            wtc = ops.pack_conv_tc(w)
            if is_hw:
                wtc2 = ops.pack_conv_hw3_tc2(w, self.half_split)
        return _Packed(packed.contiguous(), None if bias is None else bias.contiguous(), cout, wtc, wtc2)

    def _pack(self) -> Dict[str, _Packed]:
        sd = dict(self.state_dict(keep_vars=True))
        pk: Dict[str, _Packed] = {}

        def sep(p, transposed=False):
            for i in ("0", "1"):
                pk[f"{p}.conv.{i}"] = self._fold(sd, f"{p}.conv.{i}", f"{p}.conv.{i}.norm", transposed)

        def init3d(p):
            sep(p + ".0")
            for n in ("conv1", "conv2", "conv3", "conv4", "shortcut5", "shortcut6"):
                sep(f"{p}.1.{n}")
            sep(f"{p}.1.conv5", True)
            sep(f"{p}.1.conv6", True)
            sep(p + ".2")

        def heads(p):
            # both (3,1,1) head convs fused into one Cout = 2C conv: [cost-head feats | offset-head feats]
            a = self._fold(sd, f"{p}.cost_head.0", f"{p}.cost_head.0.norm")
            b = self._fold(sd, f"{p}.off_head.0", f"{p}.off_head.0.norm")
            c = a.cout
            w = torch.cat([a.w[:, :, :c], b.w[:, :, :c]], 2)
            cp = (2 * c + 3) // 4 * 4
            wp = torch.zeros((w.shape[0], w.shape[1], cp), device=w.device)
            wp[:, :, :2 * c] = w
            wtc = None
            if self.tensor_cores and w.is_cuda and 2 * c <= 64 and w.shape[0] >= 8:
                wtc = ops.pack_conv_tc(w.permute(2, 0, 1).contiguous())            # [Cin][k][2C] -> [2C][Cin][k]
            pk[p + ".stem"] = _Packed(wp.contiguous(), torch.cat([a.b, b.b]).contiguous(), 2 * c, wtc)
            w1 = torch.stack([sd[f"{p}.cost_head.1.weight"].detach().float().reshape(c, 9),
                              sd[f"{p}.off_head.1.weight"].detach().float().reshape(c, 9)])
            pk[p + ".final"] = _Packed(w1.contiguous(), None, 2)

        for lvl in ("coarse", "fine"):
            init3d(f"{lvl}.init3d")
            pc = self._fold(sd, f"{lvl}.past_conv", f"{lvl}.past_conv.norm")
            c = pc.cout
            pk[f"{lvl}.past_conv"] = _Packed(pc.w[0, 0, :c].contiguous(), pc.b, c)
            pk[f"{lvl}.fuse.conv_5x5"] = self._fold(sd, f"{lvl}.fuse.conv_5x5", f"{lvl}.fuse.conv_5x5.norm")
            sep(f"{lvl}.fuse.conv_fuse")
            heads(f"{lvl}.pred_heads")
            m = f"{lvl}.convex_upsample.mask"
            pk[m + ".0"] = self._fold(sd, m + ".0", m + ".1")
            pk[m + ".3"] = _Packed(sd[m + ".3.weight"].detach().float().reshape(36, 64).contiguous(),
                                   sd[m + ".3.bias"].detach().float().contiguous(), 36)
        init3d("precise.init3d")
        heads("precise.pred_heads")
        r = "precise.refinement"
        for n in ("conv2.0", "conv2.1", "conv4.0", "conv4.1", "fuse.0", "fuse.1", "concat"):
            pk[f"{r}.{n}"] = self._fold(sd, f"{r}.{n}", f"{r}.{n}.norm")
        pk[f"{r}.deconv4"] = self._fold(sd, f"{r}.deconv4", f"{r}.deconv4.norm", True)
        pk[f"{r}.deconv2"] = self._fold(sd, f"{r}.deconv2", None, True)
        return pk

    # ------------------------------------------------------------------ building blocks
    def _pick(self, key, cands):
        """Plan cache (per-shape plan cache): which kernel runs this layer shape.
        `cands` maps a kernel name ("tc2", "simt") to a thunk; plan_mode "auto" times each on the first
        call of a shape (outside CUDA-graph capture) and keeps the fastest; any other plan_mode forces that
        kernel where it exists."""
        names = list(cands)
        mode = self.plan_mode.get(key[0], "auto") if isinstance(self.plan_mode, dict) else self.plan_mode
        if mode != "auto":
            choice = mode if mode in cands else names[0]
        else:
            choice = self._plan.get(key)
        if choice is None:
            if torch.cuda.is_current_stream_capturing() or len(names) == 1:
                choice = names[0]
            else:
                times = []
                for nme in names:
                    fn = cands[nme]
                    fn()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(3):
                        fn()
                    e1.record()
                    e1.synchronize()
                    times.append(e0.elapsed_time(e1))
                choice = names[times.index(min(times))]
                self._plan[key] = choice
        return cands[choice]()

    def _hw3(self, x, k: _Packed, stride=1, dil=1, act=None, out=None):
        """3x3 conv over (H,W): tensor cores (stride 1) or the fp32 FMA kernel, per the plan."""
        simt = lambda: ops.conv_hw3(x, k.w, k.b, k.cout, stride, dil, act, out=out)
        if not self.tensor_cores or not k.w.is_cuda:
            return simt()
        if stride == 2 and dil == 1 and k.w.shape[1] == 9:
            if "s2" not in k.lazy:      # [Cin][9][CoutP] -> [Cout][Cin][9]
                k.lazy["s2"] = ops.pack_conv_hw3s2_tc2(k.w[:, :, :k.cout].permute(2, 0, 1).contiguous(), self.half_split)
            return self._pick(("hw3s2", tuple(x.shape), k.cout),
                              {"tc2": lambda: ops.conv_hw3s2_tc2(x, k.lazy["s2"], k.b, k.cout, act, out=out, half=self.half_split),
                               "simt": simt})
        if k.wtc2 is None or stride != 1:
            return simt()
        # the first-generation kernel (ops.conv_hw3_tc) stays an operator of the library but is no plan candidate:
        # conv_hw3_tc2 is faster on every layer shape of the model
        cands = {"tc2": lambda: ops.conv_hw3_tc2(x, k.wtc2, k.b, k.cout, dil, act, out=out, half=self.half_split), "simt": simt}
        return self._pick(("hw3", tuple(x.shape), k.cout, dil), cands)

    def _d(self, x, k: _Packed, ksz=3, stride=1, dil=1, transposed=False, act=None, out=None):
        """(k,1,1) conv along D: tensor cores when a tcgen05 operand image exists."""
        simt = lambda: ops.conv_d(x, k.w, k.b, k.cout, ksz, stride, dil, transposed, act, out=out)
        if k.wtc is None or not self.tensor_cores:
            return simt()
        key = ("d", tuple(x.shape), k.cout, ksz, stride, dil, transposed)
        if "d2" not in k.lazy:          # [Cin][k][CoutP] -> [Cout][Cin][k]
            k.lazy["d2"] = ops.pack_conv_d_tc2(k.w[:, :, :k.cout].permute(2, 0, 1).contiguous(), self.half_split)
        return self._pick(key, {"tc2": lambda: ops.conv_d_tc2(x, k.lazy["d2"], k.b, k.cout, ksz, stride, dil, transposed, act, out=out,
                                                              half=self.half_split),
                                "simt": simt})

    def _deconv_hw(self, x, k: _Packed, ksz, act=None, out=None):
        """Stride-2 transposed (1,k,k) / kxk conv: four tensor-core phase launches or the fp32 FMA kernel."""
        simt = lambda: ops.deconv_hw(x, k.w, k.b, k.cout, ksz, act, out=out)
        if not self.tensor_cores or not k.w.is_cuda or k.w.shape[0] < 8:
            return simt()
        if "dc" not in k.lazy:          # [Cin][k*k][CoutP] (transposed-conv tap order) -> [Cout][Cin][k*k]
            k.lazy["dc"] = ops.pack_deconv_hw_tc2(k.w[:, :, :k.cout].permute(2, 0, 1).contiguous(), ksz, self.half_split)
        return self._pick(("dc", tuple(x.shape), k.cout, ksz),
                          {"tc2": lambda: ops.deconv_hw_tc2(x, k.lazy["dc"], k.b, k.cout, act, out=out, half=self.half_split),
                           "simt": simt})

    def _sep(self, x, p, stride=1, dil=1, act0="SiLU", act1="SiLU", out=None):
        """'DepthwiseConv3D': (1,3,3) conv then (3,1,1) conv, BN folded (reference module.py:111-147)."""
        a, b = self._pk[p + ".conv.0"], self._pk[p + ".conv.1"]
        y = self._hw3(x, a, stride, dil, act0)
        return self._d(y, b, 3, stride, dil, False, act1, out=out)

    def _sep_t(self, x, p):
        """'DepthwiseConvTranspose3D' k3 s2 p1 op1, no activation (reference module.py:149-184)."""
        a, b = self._pk[p + ".conv.0"], self._pk[p + ".conv.1"]
        y = self._deconv_hw(x, a, 3)
        return self._d(y, b, 3, 2, 1, True, None)

    def _hourglass(self, x, p):
        """ResidualBlock3D (reference module.py:271-297)."""
        o = self._sep(x, p + ".conv1", stride=2)
        pre = self._sep(o, p + ".conv2")
        o = self._sep(pre, p + ".conv3", stride=2)
        o = self._sep(o, p + ".conv4", act0=None, act1="SiLU")
        o = self._sep_t(o, p + ".conv5")
        sc = self._sep(pre, p + ".shortcut5", act0=None, act1=None)
        o = ops.resize_add_act(o, pre.shape[-3:], sc, "SiLU")
        o = self._sep_t(o, p + ".conv6")
        sc = self._sep(x, p + ".shortcut6", act0=None, act1=None)
        return ops.resize_add_act(o, x.shape[-3:], sc, "SiLU")

    def _init3d(self, raw, p):
        y = self._sep(raw, p + ".0")
        y = self._hourglass(y, p + ".1")
        return self._sep(y, p + ".2", dil=2)

    def _heads_predict(self, vol, samples, p, delta, want_top=False):
        st, fin = self._pk[p + ".stem"], self._pk[p + ".final"]
        feat = self._d(vol, st, 3, 1, 1, False, "SiLU")
        cost, off = ops.heads(feat, fin.w, delta)
        disp, td, tc = ops.predict_disp(cost, samples, off, want_top)
        return disp, cost, off, td, tc

    def _linspace_samples(self, B, n, H, W, device):
        key = ("lin", B, n, H, W, str(device))
        if key not in self._const:
            s = torch.linspace(0, n - 1, n, device=device).view(1, n, 1, 1).expand(B, n, H, W).contiguous()
            self._const[key] = s
        return self._const[key]

    def _memory_level(self, lvl, left, right, samples, prev_info, coarse):
        """Shared body of CoarseAggregation / FineAggregation.forward after candidate generation
        (reference coarse.py:77-116, fine.py:97-132)."""
        cfg = self.levels[lvl]
        C = cfg["C"]
        B, _, H, W = left.shape
        raw = ops.block_cost(left, right, cfg["num_sample"] if coarse else samples)
        if coarse:
            samples = self._linspace_samples(B, cfg["num_sample"], H, W, left.device)
        vol = self._init3d(raw, f"{lvl}.init3d")
        D = vol.shape[2]
        ms = mv = None
        memory = prev_info.get("cost_memory", None)
        if memory is not None and prev_info.get("use_past_cost", False):
            ms, mv = memory["disp_sample"].contiguous(), memory["cost_volume"].contiguous()
            if coarse:
                mw = ms.shape[-1]
                ms = ops.bilinear_resize(ms, (H, W), mul=W, div=mw)
                mv = ops.bilinear_resize(mv, (H, W))
            assert ms.shape == (B, 2, H, W) and mv.shape == (B, 2, H, W), "cost memory / level resolution mismatch"
        pc = self._pk[f"{lvl}.past_conv"]
        cat = torch.empty((B, 4 * C, D + 2, H, W), device=left.device, dtype=torch.float32)
        _, samples = ops.merge_memory(vol, samples, ms, mv, pc.w, pc.b, 2, out_vol=cat[:, :C])
        c5 = self._pk[f"{lvl}.fuse.conv_5x5"]
        self._d(cat[:, :C], c5, 5, 1, 1, False, "SiLU", out=cat[:, C:2 * C])
        ops.pool5(cat[:, :C], cat[:, 2 * C:3 * C], cat[:, 3 * C:])
        vol = self._sep(cat, f"{lvl}.fuse.conv_fuse", act0=None, act1=None)
        disp, cost, off, _, _ = self._heads_predict(vol, samples, f"{lvl}.pred_heads", float(cfg["delta"]))
        m0, m3 = self._pk[f"{lvl}.convex_upsample.mask.0"], self._pk[f"{lvl}.convex_upsample.mask.3"]
        mfeat = self._hw3(left, m0, 1, 1, "SiLU")
        up = ops.convex_upsample(mfeat, m3.w, m3.b, disp)
        return up, cost, off, samples

    def _side_stream(self, dev) -> "torch.cuda.Stream":
        key = str(dev)
        if key not in self._side:
            self._side[key] = torch.cuda.Stream(device=dev)
        return self._side[key]

    def _conv2d(self, x, p, stride=1, act="ReLU", out=None):
        k = self._pk[p]
        return self._hw3(x, k, stride, 1, act, out=out)

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, left_feats: List[torch.Tensor], right_feats: List[torch.Tensor], left_image: torch.Tensor,
                right_image: torch.Tensor, prev_info: Optional[dict] = None):
        if prev_info is None:
            prev_info = {}
        if self._pk is None:
            self._pk = self._pack()
        l4, l8, l16 = [t.contiguous() for t in left_feats]
        r4, r8, r16 = [t.contiguous() for t in right_feats]
        left_image, right_image = left_image.contiguous(), right_image.contiguous()
        dev = l4.device
        B = l4.shape[0]
        if not all(t.is_cuda for t in (l4, l8, l16, r4, r8, r16, left_image, right_image)):
            raise TypeError("libtstereo ops need fp32 CUDA tensors (there is no CPU fallback for the hot path)")

        # ---- UNet encoder (1/2- and 1/4-scale image features, reference module.py:459-466): independent of the
        #      coarse and fine levels, so it runs on a side stream while their small, latency-bound launches
        #      leave most SMs idle.  Left and right images go through as one batch of 2B.
        H4, W4 = l4.shape[-2:]
        H, W = left_image.shape[-2:]
        r = "precise.refinement"
        cf = l4.shape[1]
        c4 = self._pk[r + ".conv4.1"].cout
        c2 = self._pk[r + ".conv2.1"].cout
        lrcat = torch.empty((2 * B, cf + c4, H4, W4), device=dev, dtype=torch.float32)
        lcat, rcat = lrcat[:B], lrcat[B:]
        cat2lr = torch.empty((2 * B, 2 * c2, (H - 1) // 2 + 1, (W - 1) // 2 + 1), device=dev, dtype=torch.float32)
        cat2 = cat2lr[:B]                 # [deconv4 output | left 1/2-scale features]; the right half's first c2 planes stay unused
        images = torch.cat([left_image, right_image], 0)
        main = torch.cuda.current_stream(dev)
        side = self._side_stream(dev) if self.overlap_encoder else main
        if side is not main:
            side.wait_stream(main)
        with torch.cuda.stream(side):
            lcat[:, :cf].copy_(l4)
            rcat[:, :cf].copy_(r4)
            self._conv2d(self._conv2d(images, r + ".conv2.0", 2), r + ".conv2.1", out=cat2lr[:, c2:])
            self._conv2d(self._conv2d(cat2lr[:, c2:], r + ".conv4.0", 2), r + ".conv4.1", out=lrcat[:, cf:])
            enc_done = torch.cuda.Event()
            enc_done.record(side)

        # ---- coarse (1/16): integer-shift volume over num_sample candidates
        d_c, c_c, o_c, s_c = self._memory_level("coarse", l16, r16, None, prev_info, True)

        # ---- fine (1/8): [local-map candidates | 5 range candidates]
        H8, W8 = l8.shape[-2:]
        lm = prev_info.get("local_map", None)
        n_lm = lm.shape[1] if (lm is not None and prev_info.get("local_map_size", 0) > 0) else 0
        samples = torch.empty((B, n_lm + 5, H8, W8), device=dev, dtype=torch.float32)
        if n_lm:
            lm = lm.contiguous()
            ops.bilinear_resize(lm, (H8, W8), mul=W8, div=lm.shape[-1], out=samples, c_off=0)
        low_c, high_c = ops.range_samples(d_c, DISP_RANGE, samples, n_lm)
        d_f, c_f, o_f, s_f = self._memory_level("fine", l8, r8, samples, prev_info, False)

        # ---- precise (1/4): UNet encoder features concatenated to the backbone features
        if side is not main:
            main.wait_event(enc_done)
            for t in (images, l4, r4):            # read on the side stream: keep the allocator from reusing them early
                t.record_stream(side)

        samples_p = torch.empty((B, 5, H4, W4), device=dev, dtype=torch.float32)
        low_f, high_f = ops.range_samples(d_f, DISP_RANGE, samples_p, 0)
        raw = ops.block_cost(lcat, rcat, samples_p)
        vol = self._init3d(raw, "precise.init3d")
        d_p, c_p, o_p, top_disp, top_cost = self._heads_predict(vol, samples_p, "precise.pred_heads",
                                                                float(self.levels["precise"]["delta"]), True)
        f = self._conv2d(self._conv2d(lcat, r + ".fuse.0"), r + ".fuse.1")
        self._deconv_hw(f, self._pk[r + ".deconv4"], 4, "ReLU", out=cat2[:, :c2])
        f = self._conv2d(cat2, r + ".concat")
        logits = self._deconv_hw(f, self._pk[r + ".deconv2"], 4)
        full = ops.unet_upsample(logits, d_p)

        # ---- recurrent state write-back (reference precise.py:98-103)
        prev_info["prev_disp"] = full
        half = (int(H4 * 0.5), int(W4 * 0.5))
        prev_info["cost_memory"] = {
            "disp_sample": ops.bilinear_resize(top_disp, half, mul=1.0, div=2.0),
            "cost_volume": ops.bilinear_resize(top_cost, half),
        }
        return ([full, d_p, d_f, d_c], [c_p, c_f, c_c], [samples_p, s_f, s_c], [o_p, o_f, o_c],
                [{"low": low_f, "high": high_f}, {"low": low_c, "high": high_c}], prev_info)


def build_aggregation(cfg) -> nn.Module:
    """Reference `build_aggregation` (architecture/modeling/aggregation/builder.py:12-20)."""
    name = cfg["MODEL"]["AGGREGATION"]["NAME"] if isinstance(cfg, dict) else cfg.MODEL.AGGREGATION.NAME
    return AGGREGATION_REGISTRY.get(name)(cfg)
