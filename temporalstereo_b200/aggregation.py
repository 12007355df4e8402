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
import warnings
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
    """One layer's weights as the kernels read them (BatchNorm folded): `w` [Cin][taps][CoutP] + `b` for the fp32 FMA
    kernels, `tc[kind]` the tensor-core operand images ("hw3" stride-1 3x3, "s2" stride-2 3x3, "dc" stride-2
    transposed, "d" (k,1,1) along D, "cost" / "left" the fused cost -> first-conv pair).  Built on the CPU and
    uploaded once per checkpoint load."""
    __slots__ = ("w", "b", "cout", "tc", "osc")

    def __init__(self, w, b, cout, tc=None, osc=None):
        self.w, self.b, self.cout = w, b, cout
        self.tc = tc or {}
        self.osc = osc          # fp16 split: 1 / the per-output-channel weight pre-scale (ops.fp16_prescale), else None


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
        # per-(layer shape) choice between the tensor-core and the fp32-FMA kernel: "auto" decides from the shape
        # alone (`_rule`: deterministic, the same on every run and rank); "timed" times both on the first call of a
        # shape and keeps the faster (experiments); "tc2" / "simt" force one (or a dict per operator kind:
        # "hw3", "hw3s2", "d", "dc")
        self.plan_mode = "auto"
        self._plan: Dict[tuple, str] = {}
        self._plan_times: Dict[tuple, dict] = {}
        # the raw cost volume of a level never touches HBM: its first (1,3,3) conv rebuilds it in the producer
        # (ops.cost_conv_*).  True / False / a collection of level names.  Default: the two warp levels (fine, precise:
        # the left half of the volume is hoisted out of the candidate loop, -23 % / -20 % measured); the coarse shift
        # volume (34 MB per frame, nothing to hoist) is cheaper materialised with ops.block_cost (B200: 406 vs 581 us at B=8)
        self.fuse_cost = ("fine", "precise")
        # how a fused warp level's first conv runs: "taps" — the channel contraction of the right half commutes with the
        # warp (a per-position lerp shared by all channels), so it is done ONCE per frame as a 1x1 conv (T = 9 taps x Cout
        # channels) and each candidate only gathers / lerps T (ops.cost_taps) next to a 3x3 conv over the group-wise
        # channels; "producer" — the tensor-core conv's producer rebuilds the warped right features per candidate
        # (ops.cost_conv_warp; B200, precise level B = 8: 545 us against ~150 us for projection + gather + group conv)
        self.cost_form = "taps"
        # the UNet decoder (fuse -> deconv4 -> concat -> deconv2: the mask logits of the final convex up-sampling, after
        # the last top-2 selection) runs single-term fp16 MMAs: measured on the oracle, rounding its operands to fp16 moves
        # the full-resolution disparity by 7.6e-5 px EPE (max 8e-4) and nothing else (tests/tools/precision_probe_decoder.py);
        # every layer upstream of a top-2 selection keeps the 3-term hi+lo split
        self.decoder_single_term = True
        self._warned_train = False
        # per-level parity tests only: {"coarse_disp": t, "fine_disp": t} replace the engine's own up-sampled coarse /
        # fine disparity as the centre of the next level's candidates ("teacher forcing"), so that a level can be compared
        # with the oracle on bit-identical candidates; the outputs still carry the engine's own disparities
        self._inject: Optional[dict] = None
        # tensor-core operand split: fp16 hi + lo (kind::f16, 16 channels per MMA; activations < 65504) or tf32 hi + lo
        self.half_split = True
        # S-format activations between tensor-core convolutions (include/tstereo.h `tstereo_split`): the epilogue of a
        # layer writes the fp16 hi / lo halves its consumer's MMAs need, laid out so that the consumer stages a K-chunk
        # with one TMA box (no producer warps, no per-tile conversion).  Bit-identical results (tests/test_gpu_split.py);
        # needs the fp16 operand split.  False: every convolution reads and writes fp32 NC(D)HW
        self.split_format = True
        # run the UNet encoder on a side stream, concurrently with the coarse and fine levels
        self.overlap_encoder = True
        # independent branches inside a level (hourglass shortcuts, mask conv, left / projection convs) on a second side stream
        self.overlap_branches = True
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
        """The engine stays in folded-BatchNorm inference whatever the caller toggles: the reference trainer calls
        `self.train()` after every history frame (projects/TemporalStereo/TemporalStereo.py:268-274) and Lightning
        calls it around fit / validate / test, so `train(True)` must not raise.  It warns once; gradients are what
        is not available (`forward` raises if an input requires grad while grad mode is on)."""
        if mode and not self._warned_train:
            self._warned_train = True
            warnings.warn("libtstereo TEMPORALSTEREO is an inference engine (BatchNorm folded with running statistics): "
                          "train(True) keeps it in eval mode", stacklevel=2)
        return super().train(False)

    # ------------------------------------------------------------------ weight packing (CPU; one upload)
    @staticmethod
    def _fold(sd, conv: str, bn: Optional[str], transposed: bool = False):
        """[Cout,Cin,*k] (or [Cin,Cout,*k]) -> (w [Cout, Cin, taps], bias [Cout] | None) with eval-mode BN folded in."""
        w = sd[conv + ".weight"]
        if transposed:
            w = w.transpose(0, 1)
        cout, cin = w.shape[:2]
        w = w.reshape(cout, cin, -1)
        bias = sd.get(conv + ".bias")
        if bn is not None:
            s = sd[bn + ".weight"] / torch.sqrt(sd[bn + ".running_var"] + BN_EPS)
            w = w * s.view(-1, 1, 1)
            b0 = bias if bias is not None else torch.zeros_like(s)
            bias = (b0 - sd[bn + ".running_mean"]) * s + sd[bn + ".bias"]
        return w.contiguous(), bias

    def _mk(self, w: torch.Tensor, bias: Optional[torch.Tensor], kinds=(), ksz: int = 3) -> _Packed:
        """w [Cout, Cin, taps] (CPU) -> the FMA layout [Cin][taps][CoutP] plus the tensor-core operand images in `kinds`."""
        cout, cin, T = w.shape
        packed = torch.zeros((cin, T, (cout + 3) // 4 * 4), dtype=torch.float32)
        packed[:, :, :cout] = w.permute(1, 2, 0)
        tc, osc = {}, None
        if self.tensor_cores:
            h = self.half_split
            ws = w
            if h:
                ws, osc = ops.fp16_prescale(w)
            for kind in kinds:
                if kind == "hw3" and cin >= 8 and T == 9:
                    tc[kind] = ops.pack_conv_hw3_tc2(ws, h)
                elif kind == "s2" and T == 9:
                    tc[kind] = ops.pack_conv_hw3s2_tc2(ws, h)
                elif kind == "dc" and cin >= 8:
                    tc[kind] = ops.pack_deconv_hw_tc2(ws, ksz, h)
                elif kind == "d" and cin >= 8 and T in (3, 5):
                    tc[kind] = ops.pack_conv_d_tc2(ws, h)
                elif kind == "cost_warp":          # [L (C) | warp(R) (C) | g (3C/8)]: the L half is hoisted out of the D loop
                    C = cin * 8 // 19
                    tc["left"] = ops.pack_conv_hw3_tc2(ws[:, :C].contiguous(), h)
                    tc["cost"] = ops.pack_conv_hw3_tc2(ws[:, C:].contiguous(), h)
                    if h and cout in (8, 16, 32):
                        # tap-projection form (ops.cost_taps): T = 1x1 conv of R with the right-half weights, one output
                        # channel per (tap, co), its own per-channel fp16 pre-scale; the group-wise channels keep a 3x3 conv
                        wt, osc_t = ops.fp16_prescale(ops.tap_projection_weights(w[:, C:2 * C].contiguous()))
                        tc["taps"] = ops.pack_conv_d_tc2(wt, True)
                        tc["taps_osc"] = osc_t
                        tc["gconv"] = ops.pack_conv_hw3_tc2(ws[:, 2 * C:].contiguous(), True)
                elif kind == "cost_shift":         # [-(L - R_d)^2 (C) | g (3C/8)]
                    tc["cost"] = ops.pack_conv_hw3_tc2(ws, h)
        return _Packed(packed, bias, cout, tc, osc if tc else None)

    def _pack(self, dev) -> Dict[str, _Packed]:
        """BatchNorm folding and operand packing run on the CPU; every packed tensor lands in ONE device buffer with
        a single host-to-device copy (no device-side torch kernels: the first forward launches only libtstereo's)."""
        sd = {k: v.detach().to("cpu", torch.float32) for k, v in self.state_dict(keep_vars=True).items()
              if v.dtype.is_floating_point}
        pk: Dict[str, _Packed] = {}
        fold, mk = self._fold, self._mk

        def sep(p, transposed=False, stride=1):
            w0, b0 = fold(sd, f"{p}.conv.0", f"{p}.conv.0.norm", transposed)
            pk[f"{p}.conv.0"] = mk(w0, b0, ("dc",) if transposed else (("s2",) if stride == 2 else ("hw3",)), 3)
            w1, b1 = fold(sd, f"{p}.conv.1", f"{p}.conv.1.norm", transposed)
            pk[f"{p}.conv.1"] = mk(w1, b1, ("d",))

        def init3d(p, warp):
            # first conv: also the operand images of the fused cost -> first-conv path (virtual channels
            # [feature half | group terms])
            w0, b0 = fold(sd, f"{p}.0.conv.0", f"{p}.0.conv.0.norm")
            pk[f"{p}.0.conv.0"] = mk(w0, b0, ("hw3", "cost_warp" if warp else "cost_shift"))
            w1, b1 = fold(sd, f"{p}.0.conv.1", f"{p}.0.conv.1.norm")
            pk[f"{p}.0.conv.1"] = mk(w1, b1, ("d",))
            sep(f"{p}.1.conv1", stride=2)
            sep(f"{p}.1.conv2")
            sep(f"{p}.1.conv3", stride=2)
            for n in ("conv4", "shortcut5", "shortcut6"):
                sep(f"{p}.1.{n}")
            sep(f"{p}.1.conv5", True)
            sep(f"{p}.1.conv6", True)
            sep(p + ".2")

        def heads(p):
            # both (3,1,1) head convs fused into one Cout = 2C conv: [cost-head feats | offset-head feats]
            wa, ba = fold(sd, f"{p}.cost_head.0", f"{p}.cost_head.0.norm")
            wb, bb = fold(sd, f"{p}.off_head.0", f"{p}.off_head.0.norm")
            c = wa.shape[0]
            pk[p + ".stem"] = mk(torch.cat([wa, wb], 0), torch.cat([ba, bb]), ("d",) if 2 * c <= 64 else ())
            w1 = torch.stack([sd[f"{p}.cost_head.1.weight"].reshape(c, 9), sd[f"{p}.off_head.1.weight"].reshape(c, 9)])
            pk[p + ".final"] = _Packed(w1.contiguous(), None, 2)

        for lvl in ("coarse", "fine"):
            cfg = self.levels[lvl]
            init3d(f"{lvl}.init3d", lvl == "fine")
            w, b = fold(sd, f"{lvl}.past_conv", f"{lvl}.past_conv.norm")
            pk[f"{lvl}.past_conv"] = _Packed(w[:, 0, 0].contiguous(), b, w.shape[0])
            w, b = fold(sd, f"{lvl}.fuse.conv_5x5", f"{lvl}.fuse.conv_5x5.norm")
            pk[f"{lvl}.fuse.conv_5x5"] = mk(w, b, ("d",))
            sep(f"{lvl}.fuse.conv_fuse")
            heads(f"{lvl}.pred_heads")
            m = f"{lvl}.convex_upsample.mask"
            w, b = fold(sd, m + ".0", m + ".1")
            pk[m + ".0"] = mk(w, b, ("hw3",))
            pk[m + ".3"] = _Packed(sd[m + ".3.weight"].reshape(36, 64).contiguous(), sd[m + ".3.bias"].contiguous(), 36)
        r = "precise.refinement"
        init3d("precise.init3d", True)
        heads("precise.pred_heads")
        for n, stride in (("conv2.0", 2), ("conv2.1", 1), ("conv4.0", 2), ("conv4.1", 1), ("fuse.0", 1), ("fuse.1", 1), ("concat", 1)):
            w, b = fold(sd, f"{r}.{n}", f"{r}.{n}.norm")
            pk[f"{r}.{n}"] = mk(w, b, ("s2",) if stride == 2 else ("hw3",))
        w, b = fold(sd, f"{r}.deconv4", f"{r}.deconv4.norm", True)
        pk[f"{r}.deconv4"] = mk(w, b, ("dc",), 4)
        w, b = fold(sd, f"{r}.deconv2", None, True)
        pk[f"{r}.deconv2"] = mk(w, b, ("dc",), 4)

        # one arena, one upload; every tensor starts 256-byte aligned (the operand images need 16)
        items = []
        for k in pk.values():
            items.append((k, "w", None))
            if k.b is not None:
                items.append((k, "b", None))
            if k.osc is not None:
                items.append((k, "osc", None))
            items += [(k, "tc", name) for name in k.tc]
        get = lambda k, f, n: (k.tc[n] if f == "tc" else getattr(k, f))
        offs, total = [], 0
        for it in items:
            offs.append(total)
            total += (get(*it).numel() + 63) // 64 * 64
        host = torch.zeros((total,), dtype=torch.float32)
        for it, o in zip(items, offs):
            t = get(*it).contiguous().view(-1)
            host[o:o + t.numel()] = t
        arena = host.to(dev)
        for (k, f, n), o in zip(items, offs):
            t = get(k, f, n)
            v = arena[o:o + t.numel()].view(t.shape)
            if f == "tc":
                k.tc[n] = v
            else:
                setattr(k, f, v)
        self._arena = arena
        return pk

    # ------------------------------------------------------------------ building blocks
    @staticmethod
    def _rule(kind: str, x_shape, cout: int, ksz: int = 3) -> str:
        """Deterministic kernel choice from the layer shape alone, fitted to device times measured on B200 at B = 1 and
        B = 8 (profiles/r02_plan_probe.md; `plan_mode = "timed"` re-measures).  3x3 forms: the tensor-core kernel always
        wins.  (k,1,1) convs along D: it wins for 32 input channels, loses for <= 16 (a few MACs per byte: the fp32 FMA
        kernel streams at HBM speed) and for 64 (two 32-channel output groups, each a 24-chunk latency chain).  Stride-2
        transposed convs are four phase launches on the tensor cores: worth it only with enough work per launch."""
        cin = x_shape[1]
        planes = x_shape[0] * (x_shape[2] if len(x_shape) == 5 else 1)
        hw = x_shape[-2] * x_shape[-1]
        if kind == "d":
            if cin >= 64 or cin < 16:
                return "simt"
            if cin >= 32:
                return "tc2"
            return "tc2" if (ksz == 5 or cout > 16) else "simt"
        if kind == "dc":
            return "tc2" if (cin >= 64 and planes * hw >= 20000) or hw >= 30000 else "simt"
        return "tc2"

    def _pick(self, key, cands):
        """Plan cache (per-shape plan cache): which kernel runs this layer shape.  `cands` maps a kernel
        name ("tc2", "simt") to a thunk.  plan_mode "auto": `_rule` (shape only, deterministic); "timed": time each on
        the first call of a shape (outside CUDA-graph capture) and keep the fastest; anything else forces that kernel
        where it exists."""
        names = list(cands)
        mode = self.plan_mode.get(key[0], "auto") if isinstance(self.plan_mode, dict) else self.plan_mode
        if len(names) == 1:
            choice = names[0]
        elif mode == "auto":
            choice = self._rule(key[0], key[1], key[2], key[3] if key[0] == "d" else 3)
            if choice not in cands:
                choice = names[0]
        elif mode != "timed":
            choice = mode if mode in cands else names[0]
        else:
            choice = self._plan.get(key)
            if choice is None:
                if torch.cuda.is_current_stream_capturing():
                    choice = names[0]
                else:
                    times = {}
                    for nme in names:
                        # device time of the kernel itself: 8 launches recorded into a CUDA graph and replayed (launching
                        # one by one from Python costs ~20 us per call, more than most of these kernels run for)
                        fn = cands[nme]
                        fn()
                        torch.cuda.synchronize()
                        g = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g):
                            for _ in range(8):
                                fn()
                        g.replay()
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                        g.replay()
                        e1.record()
                        e1.synchronize()
                        times[nme] = e0.elapsed_time(e1) / 8 * 1e3
                        del g
                    choice = min(times, key=times.get)
                    self._plan[key] = choice
                    self._plan_times[key] = times
        return cands[choice]()

    def _sfmt(self) -> bool:
        """S-format activations between the tensor-core convolutions (see `split_format`)."""
        return bool(self.split_format and self.half_split and self.tensor_cores)

    @staticmethod
    def _new_split(shape, like, parts=2) -> "ops.Split":
        five = len(shape) == 5
        B, C = shape[:2]
        D = shape[2] if five else 1
        dev = like.t.device if isinstance(like, ops.Split) else like.device
        return ops.Split(B, C, D, shape[-2], shape[-1], parts, device=dev, five=five)

    def _hw3(self, x, k: _Packed, stride=1, dil=1, act=None, out=None, single=False, fmt="f"):
        """3x3 conv over (H,W): tensor cores or the fp32 FMA kernel, per the plan.  single: one fp16 MMA term.
        `x` is an fp32 tensor or an S-format `ops.Split`; fmt "s" returns the result as a Split (S-format mode only)."""
        if isinstance(x, ops.Split) or fmt == "s":
            h = 2 if (single and self.decoder_single_term) else 1
            shp = tuple(x.shape)
            if stride == 2:
                oshp = shp[:1] + (k.cout,) + shp[2:-2] + ((shp[-2] - 1) // 2 + 1, (shp[-1] - 1) // 2 + 1)
                so = self._new_split(oshp, x) if fmt == "s" else None
                f, _ = ops.conv_hw3s2_s(x, k.tc["s2"], k.b, k.cout, act, out=out, half=h, oscale=k.osc, sout=so)
            else:
                so = self._new_split(shp[:1] + (k.cout,) + shp[2:], x) if fmt == "s" else None
                f, _ = ops.conv_hw3_s(x, k.tc["hw3"], k.b, k.cout, dil, act, out=out, half=h, oscale=k.osc, sout=so)
            return so if fmt == "s" else f
        simt = lambda: ops.conv_hw3(x, k.w, k.b, k.cout, stride, dil, act, out=out)
        h = 2 if (single and self.half_split and self.decoder_single_term) else self.half_split
        if stride == 2 and dil == 1 and "s2" in k.tc:
            return self._pick(("hw3s2", tuple(x.shape), k.cout),
                              {"tc2": lambda: ops.conv_hw3s2_tc2(x, k.tc["s2"], k.b, k.cout, act, out=out, half=h, oscale=k.osc), "simt": simt})
        if stride != 1 or "hw3" not in k.tc:
            return simt()
        return self._pick(("hw3", tuple(x.shape), k.cout, dil),
                          {"tc2": lambda: ops.conv_hw3_tc2(x, k.tc["hw3"], k.b, k.cout, dil, act, out=out, half=h, oscale=k.osc), "simt": simt})

    def _d(self, x, k: _Packed, ksz=3, stride=1, dil=1, transposed=False, act=None, out=None, fmt="f"):
        """(k,1,1) conv along D: tensor cores when a tcgen05 operand image exists.  S-format operands as in `_hw3`."""
        if isinstance(x, ops.Split) or fmt == "s":
            B, _, Din, H, W = x.shape
            Dout = 2 * Din if transposed else (Din - 1) // stride + 1
            so = self._new_split((B, k.cout, Dout, H, W), x) if fmt == "s" else None
            f, _ = ops.conv_d_s(x, k.tc["d"], k.b, k.cout, ksz, stride, dil, transposed, act, out=out, half=1, oscale=k.osc, sout=so)
            return so if fmt == "s" else f
        simt = lambda: ops.conv_d(x, k.w, k.b, k.cout, ksz, stride, dil, transposed, act, out=out)
        if "d" not in k.tc:
            return simt()
        key = ("d", tuple(x.shape), k.cout, ksz, stride, dil, transposed)
        return self._pick(key, {"tc2": lambda: ops.conv_d_tc2(x, k.tc["d"], k.b, k.cout, ksz, stride, dil, transposed, act, out=out,
                                                              half=self.half_split, oscale=k.osc),
                                "simt": simt})

    def _deconv_hw(self, x, k: _Packed, ksz, act=None, out=None, single=False, fmt="f"):
        """Stride-2 transposed (1,k,k) / kxk conv: four tensor-core phase launches or the fp32 FMA kernel.  S-format
        operands as in `_hw3`."""
        if isinstance(x, ops.Split) or fmt == "s":
            h = 2 if (single and self.decoder_single_term) else 1
            shp = tuple(x.shape)
            so = self._new_split(shp[:1] + (k.cout,) + shp[2:-2] + (2 * shp[-2], 2 * shp[-1]), x) if fmt == "s" else None
            f, _ = ops.deconv_hw_s(x, k.tc["dc"], k.b, k.cout, act, out=out, half=h, oscale=k.osc, sout=so)
            return so if fmt == "s" else f
        simt = lambda: ops.deconv_hw(x, k.w, k.b, k.cout, ksz, act, out=out)
        if "dc" not in k.tc:
            return simt()
        h = 2 if (single and self.half_split and self.decoder_single_term) else self.half_split
        return self._pick(("dc", tuple(x.shape), k.cout, ksz),
                          {"tc2": lambda: ops.deconv_hw_tc2(x, k.tc["dc"], k.b, k.cout, act, out=out, half=h, oscale=k.osc),
                           "simt": simt})

    def _sep(self, x, p, stride=1, dil=1, act0="SiLU", act1="SiLU", out=None, fmt="f"):
        """'DepthwiseConv3D': (1,3,3) conv then (3,1,1) conv, BN folded (reference module.py:111-147).  In S-format mode
        the intermediate never exists in fp32."""
        a, b = self._pk[p + ".conv.0"], self._pk[p + ".conv.1"]
        y = self._hw3(x, a, stride, dil, act0, fmt="s" if self._sfmt() else "f")
        return self._d(y, b, 3, stride, dil, False, act1, out=out, fmt=fmt)

    def _sep_t(self, x, p, fmt="f"):
        """'DepthwiseConvTranspose3D' k3 s2 p1 op1, no activation (reference module.py:149-184)."""
        a, b = self._pk[p + ".conv.0"], self._pk[p + ".conv.1"]
        y = self._deconv_hw(x, a, 3, fmt="s" if self._sfmt() else "f")
        return self._d(y, b, 3, 2, 1, True, None, fmt=fmt)

    def _hourglass(self, x, p):
        """ResidualBlock3D (reference module.py:271-297).  S-format mode: `x` and the result are `ops.Split`s; only the
        operands of the two resize + add + SiLU kernels exist in fp32."""
        sf = self._sfmt()
        m = "s" if sf else "f"
        dev = x.t.device if isinstance(x, ops.Split) else x.device
        # the two shortcuts depend on x / pre only: they run beside the down-up chain (outputs allocated before the fork)
        c6 = self._pk[p + ".shortcut6.conv.1"].cout
        sc6 = torch.empty((x.shape[0], c6) + tuple(x.shape[2:]), device=dev, dtype=torch.float32)
        with self._branch(dev) as br6:
            self._sep(x, p + ".shortcut6", act0=None, act1=None, out=sc6)
        o = self._sep(x, p + ".conv1", stride=2, fmt=m)
        pre = self._sep(o, p + ".conv2", fmt=m)
        c5 = self._pk[p + ".shortcut5.conv.1"].cout
        sc5 = torch.empty((pre.shape[0], c5) + tuple(pre.shape[2:]), device=dev, dtype=torch.float32)
        with self._branch(dev) as br5:            # same side stream: queued behind shortcut6
            self._sep(pre, p + ".shortcut5", act0=None, act1=None, out=sc5)
        o = self._sep(pre, p + ".conv3", stride=2, fmt=m)
        o = self._sep(o, p + ".conv4", act0=None, act1="SiLU", fmt=m)
        o = self._sep_t(o, p + ".conv5")
        raa = ops.resize_add_act_s if sf else ops.resize_add_act
        br5.join()
        o = raa(o, pre.shape[-3:], sc5, "SiLU")
        o = self._sep_t(o, p + ".conv6")
        br6.join()
        return raa(o, x.shape[-3:], sc6, "SiLU")

    def _first_conv(self, left, right, samples, p, s_left=None, s_right=None):
        """block_cost -> the first (1,3,3) conv + BN + SiLU of a level's init3d (reference block_cost.py:16-83 feeding
        module.py:111-147 through coarse.py:82-83, fine.py:102-103, precise.py:88-90).  `samples` is the candidate tensor
        [B,S,H,W] (warp volume) or an int (shift volume).  With `fuse_cost` the raw volume is never materialised (see
        `cost_form`).  Returns fp32 [B,C,D,H,W] or, in S-format mode, an `ops.Split`.  (bench.py times this function.)"""
        a = self._pk[p + ".0.conv.0"]
        fuse = self.fuse_cost if isinstance(self.fuse_cost, bool) else p.split(".")[0] in self.fuse_cost
        if fuse and "cost" in a.tc:
            if isinstance(samples, int):
                g = ops.group_cost(left, right, samples)
                y = ops.cost_conv_shift(left, right, g, a.tc["cost"], a.b, a.cout, "SiLU", half=self.half_split, oscale=a.osc)
            elif self.cost_form == "taps" and "taps" in a.tc:
                B_, C_, H_, W_ = right.shape
                sf = self._sfmt()
                addl = torch.empty((B_, a.cout, H_, W_), device=right.device, dtype=torch.float32)
                T5 = torch.empty((B_, 9 * a.cout, 1, H_, W_), device=right.device, dtype=torch.float32)
                T = T5.view(B_, 9 * a.cout, H_, W_)
                # the left-half conv and the projection depend on the features only: beside the group terms (the branch's
                # outputs are allocated above, before the fork)
                with self._branch(right.device) as br:
                    if sf:
                        # both feature maps in S-format (given by the caller when a producer already wrote them): the left-half
                        # conv and the 9*Cout-channel projection (3-5 output groups, each a pass over the input) are TMA-fed
                        sl = s_left if s_left is not None else ops.split_pack(left)
                        sr = s_right if s_right is not None else ops.split_pack(right)
                        ops.conv_hw3_s(sl, a.tc["left"], None, a.cout, 1, None, out=addl, half=1, oscale=a.osc)
                        sr5 = ops.Split(B_, C_, 1, H_, W_, sr.parts, t=sr.t, five=True)
                        # 9*Cout outputs = 3-5 groups of 32, each a pass over the input: sub-batches whose S-format features
                        # (4 bytes per element) stay L2-resident, so only the first pass of a sub-batch reads DRAM
                        nb = max(1, min(B_, int(64e6 // (C_ * H_ * W_ * 4))))
                        for b0 in range(0, B_, nb):
                            b1 = min(B_, b0 + nb)
                            ops.conv_d_s(sr5.batches(b0, b1), a.tc["taps"], None, 9 * a.cout, 1, 1, 1, False, None, out=T5[b0:b1], half=1,
                                         oscale=a.tc["taps_osc"])
                    else:
                        ops.conv_hw3_tc2(left, a.tc["left"], None, a.cout, 1, None, out=addl, half=True, oscale=a.osc)
                        ops.conv_d_tc2(right.unsqueeze(2), a.tc["taps"], None, 9 * a.cout, 1, 1, 1, False, None, out=T5, half=True,
                                       oscale=a.tc["taps_osc"])
                g = ops.group_cost(left, right, samples)
                gc = ops.conv_hw3_tc2(g, a.tc["gconv"], None, a.cout, 1, None, half=True, oscale=a.osc)
                br.join()
                so = self._new_split((B_, a.cout, samples.shape[1], H_, W_), right) if sf else None
                y, so = ops.cost_taps(T, samples, gc, addl, a.b, a.cout, "SiLU", sout=so)
                if sf:
                    y = so
            else:
                g = ops.group_cost(left, right, samples)
                addl = ops.conv_hw3_tc2(left, a.tc["left"], None, a.cout, 1, None, half=self.half_split, oscale=a.osc)
                y = ops.cost_conv_warp(right, samples, g, addl, a.tc["cost"], a.b, a.cout, "SiLU", half=self.half_split, oscale=a.osc)
        elif self._sfmt() and isinstance(samples, int) and left.shape[1] % 64 == 0:
            # materialised shift volume (coarse level), written in the S-format its only consumer stages by TMA
            y = self._hw3(ops.block_cost_shift_s(left, right, samples), a, 1, 1, "SiLU", fmt="s")
        else:
            y = self._hw3(ops.block_cost(left, right, samples), a, 1, 1, "SiLU")
        return y

    def _init3d(self, left, right, samples, p, out_fmt="f", s_left=None, s_right=None):
        """block_cost -> init3d stack (reference coarse.py:82-83, fine.py:102-103, precise.py:88-90)."""
        b = self._pk[p + ".0.conv.1"]
        y = self._first_conv(left, right, samples, p, s_left, s_right)
        y = self._d(y, b, 3, 1, 1, False, "SiLU", fmt="s" if self._sfmt() else "f")
        y = self._hourglass(y, p + ".1")
        return self._sep(y, p + ".2", dil=2, fmt=out_fmt if self._sfmt() else "f")

    def _heads_predict(self, vol, samples, p, delta, want_top=False):
        st, fin = self._pk[p + ".stem"], self._pk[p + ".final"]
        feat = self._d(vol, st, 3, 1, 1, False, "SiLU")
        cost, off = ops.heads(feat, fin.w, delta)
        disp, td, tc = ops.predict_disp(cost, samples, off, want_top)
        return disp, cost, off, td, tc

    def _linspace_samples(self, B, n, H, W, device):
        key = ("lin", B, n, H, W, str(device))
        if key not in self._const:
            s = torch.linspace(0, n - 1, n).view(1, n, 1, 1).expand(B, n, H, W).contiguous()
            self._const[key] = s.to(device)
        return self._const[key]

    def _memory_level(self, lvl, left, right, samples, prev_info, coarse):
        """Shared body of CoarseAggregation / FineAggregation.forward after candidate generation
        (reference coarse.py:77-116, fine.py:97-132)."""
        cfg = self.levels[lvl]
        C = cfg["C"]
        B, _, H, W = left.shape
        # the convex up-sampling's mask conv reads the left features only: beside the whole level
        m0, m3 = self._pk[f"{lvl}.convex_upsample.mask.0"], self._pk[f"{lvl}.convex_upsample.mask.3"]
        mfeat = torch.empty((B, m0.cout, H, W), device=left.device, dtype=torch.float32)
        with self._branch(left.device) as br_mask:
            self._hw3(ops.split_pack(left) if self._sfmt() else left, m0, 1, 1, "SiLU", out=mfeat)
        vol = self._init3d(left, right, cfg["num_sample"] if coarse else samples, f"{lvl}.init3d")
        if coarse:
            samples = self._linspace_samples(B, cfg["num_sample"], H, W, left.device)
        D = vol.shape[2]
        ms = mv = None
        memory = prev_info.get("cost_memory", None)
        if memory is not None and prev_info.get("use_past_cost", False):
            ms, mv = memory["disp_sample"].contiguous(), memory["cost_volume"].contiguous()
            if coarse:
                mw = ms.shape[-1]
                ms = ops.bilinear_resize(ms, (H, W), mul=W, div=mw)
                mv = ops.bilinear_resize(mv, (H, W))
            if ms.shape != (B, 2, H, W) or mv.shape != (B, 2, H, W):
                raise ValueError(f"cost memory {tuple(ms.shape)} does not match the {lvl} level {(B, 2, H, W)}")
        pc = self._pk[f"{lvl}.past_conv"]
        cat = torch.empty((B, 4 * C, D + 2, H, W), device=left.device, dtype=torch.float32)
        _, samples = ops.merge_memory(vol, samples, ms, mv, pc.w, pc.b, 2, out_vol=cat[:, :C])
        c5 = self._pk[f"{lvl}.fuse.conv_5x5"]
        with self._branch(left.device) as br_pool:          # both read the merged volume, write their own channel slices of `cat`
            ops.pool5(cat[:, :C], cat[:, 2 * C:3 * C], cat[:, 3 * C:])
        self._d(cat[:, :C], c5, 5, 1, 1, False, "SiLU", out=cat[:, C:2 * C])
        br_pool.join()
        vol = self._sep(cat, f"{lvl}.fuse.conv_fuse", act0=None, act1=None, fmt="s" if self._sfmt() else "f")
        disp, cost, off, _, _ = self._heads_predict(vol, samples, f"{lvl}.pred_heads", float(cfg["delta"]))
        br_mask.join()
        up = ops.convex_upsample(mfeat, m3.w, m3.b, disp)
        return up, cost, off, samples

    def _side_stream(self, dev, n: int = 0) -> "torch.cuda.Stream":
        key = f"{dev}/{n}"
        if key not in self._side:
            self._side[key] = torch.cuda.Stream(device=dev)
        return self._side[key]

    class _Branch:
        """`with self._branch(dev) as br: ...` runs the block on a second side stream, forked from the current stream; `br.join()`
        makes the current stream wait for it.  Most launches of the 3-D levels fill a fraction of the 148 SMs and sit on
        their fixed latency, so independent branches (the hourglass shortcuts, the mask conv, the left / projection convs
        of the first conv) overlap for free; captured into the CUDA graph they become parallel branches.  Contract: every
        tensor the branch hands back is allocated by the CALLER before the fork (`out=`), like the encoder's buffers, so
        the caching allocator never recycles it under the other stream; temporaries inside the block live and die on the
        side stream."""

        def __init__(self, eng, dev):
            self.main = torch.cuda.current_stream(dev)
            self.side = eng._side_stream(dev, 1) if eng.overlap_branches else self.main
            self.ctx = None
            self.done = None

        def __enter__(self):
            if self.side is not self.main:
                self.side.wait_stream(self.main)
                self.ctx = torch.cuda.stream(self.side)
                self.ctx.__enter__()
            return self

        def __exit__(self, *exc):
            if self.ctx is not None:
                self.done = torch.cuda.Event()
                self.done.record(self.side)
                self.ctx.__exit__(*exc)
            return False

        def join(self):
            if self.done is not None:
                self.main.wait_event(self.done)
                self.done = None

    def _branch(self, dev):
        return TEMPORALSTEREO._Branch(self, dev)

    def _conv2d(self, x, p, stride=1, act="ReLU", out=None, single=False):
        k = self._pk[p]
        return self._hw3(x, k, stride, 1, act, out=out, single=single)

    @staticmethod
    def _check_pyramid(l4, l8, l16, r4, r8, r16, left_image, right_image):
        """The reference fails with a shape error on sizes that are not multiples of 16 (cat of
        135- and 136-row tensors); the kernels write into pre-sized views, so the geometry is validated up front."""
        B = l4.shape[0]
        H16, W16 = l16.shape[-2:]
        want = {"left 1/8": (l8, (2 * H16, 2 * W16)), "left 1/4": (l4, (4 * H16, 4 * W16)),
                "left image": (left_image, (16 * H16, 16 * W16))}
        for name, (t, hw) in want.items():
            if tuple(t.shape[-2:]) != hw or t.shape[0] != B:
                raise ValueError(f"{name} has shape {tuple(t.shape)}; the 1/16 features {tuple(l16.shape)} need (H, W) = {hw} "
                                 "(input sizes must be multiples of 16: pad or resize like the reference's data pipeline)")
        for a, b_, name in ((l4, r4, "1/4"), (l8, r8, "1/8"), (l16, r16, "1/16"), (left_image, right_image, "image")):
            if a.shape != b_.shape:
                raise ValueError(f"left / right {name} shapes differ: {tuple(a.shape)} vs {tuple(b_.shape)}")

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def _forward(self, left_feats: List[torch.Tensor], right_feats: List[torch.Tensor], left_image: torch.Tensor,
                 right_image: torch.Tensor, prev_info: dict):
        l4, l8, l16 = [t.contiguous() for t in left_feats]
        r4, r8, r16 = [t.contiguous() for t in right_feats]
        left_image, right_image = left_image.contiguous(), right_image.contiguous()
        dev = l4.device
        B = l4.shape[0]
        if not all(t.is_cuda for t in (l4, l8, l16, r4, r8, r16, left_image, right_image)):
            raise TypeError("libtstereo ops need fp32 CUDA tensors (there is no CPU fallback for the hot path)")
        self._check_pyramid(l4, l8, l16, r4, r8, r16, left_image, right_image)
        if self._pk is None:
            self._pk = self._pack(dev)

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
        if not (self.split_format and self.half_split and self.tensor_cores):
            cat2lr = torch.empty((2 * B, 2 * c2, (H - 1) // 2 + 1, (W - 1) // 2 + 1), device=dev, dtype=torch.float32)
            cat2 = cat2lr[:B]             # [deconv4 output | left 1/2-scale features]; the right half's first c2 planes stay unused
            half2 = torch.empty((2 * B, c2, (H - 1) // 2 + 1, (W - 1) // 2 + 1), device=dev, dtype=torch.float32)
        main = torch.cuda.current_stream(dev)
        side = self._side_stream(dev) if self.overlap_encoder else main
        if side is not main:
            side.wait_stream(main)
        sfmt = self.split_format and self.half_split and self.tensor_cores
        single = self.decoder_single_term
        if sfmt:        # allocated on the main stream (like lrcat): the decoder reads them there after the encoder's event
            H2, W2 = (H - 1) // 2 + 1, (W - 1) // 2 + 1
            s_img = ops.Split(2 * B, left_image.shape[1], 1, H, W, 2, device=dev, five=False)
            s_half2 = ops.Split(2 * B, c2, 1, H2, W2, 2, device=dev, five=False)
            s_cat2 = ops.Split(2 * B, 2 * c2, 1, H2, W2, 2, device=dev, five=False)
            s_q = ops.Split(2 * B, c4, 1, H4, W4, 2, device=dev, five=False)
            s_lrcat = ops.Split(2 * B, cf + c4, 1, H4, W4, 2, device=dev, five=False)      # S-format of lrcat (both images)
            # the decoder (mask logits of the final up-sampling) depends on the encoder only, not on any level: it follows the
            # encoder on the side stream, beside the three levels (hi half only when it runs single-term MMAs)
            h_dec, np_ = (2, 1) if single else (1, 2)
            s_f0 = ops.Split(B, self._pk[r + ".fuse.0"].cout, 1, H4, W4, np_, device=dev, five=False)
            s_f1 = ops.Split(B, self._pk[r + ".fuse.1"].cout, 1, H4, W4, np_, device=dev, five=False)
            s_cc = ops.Split(B, self._pk[r + ".concat"].cout, 1, H2, W2, np_, device=dev, five=False)
            logits = torch.empty((B, self._pk[r + ".deconv2"].cout, H, W), device=dev, dtype=torch.float32)
        with torch.cuda.stream(side):
            ops.copy_planes(l4, lcat[:, :cf])
            ops.copy_planes(r4, rcat[:, :cf])
            if sfmt:
                # S-format chain: image -> conv2.0 -> conv2.1 -> conv4.0 -> conv4.1; only conv4.1 also writes fp32 (the
                # 1/4-scale features of the precise cost volume).  s_cat2 = [deconv4 output | conv2.1 output] is the
                # decoder's concat buffer, s_lrcat = [backbone features | conv4.1 output] of both images: the input of fuse.0 and of the precise cost path
                pk = self._pk
                # the two images meet in one 2B batch of S-format pixels (3 channels + 5 zeros per 16-byte vector), so
                # conv2.0 is ONE TMA-fed launch instead of two register-producer launches over strided fp32 loads
                ops.split_pack(left_image, out=s_img.batches(0, B))
                ops.split_pack(right_image, out=s_img.batches(B, 2 * B))
                k = pk[r + ".conv2.0"]
                ops.conv_hw3s2_s(s_img, k.tc["s2"], k.b, k.cout, "ReLU", oscale=k.osc, sout=s_half2)
                k = pk[r + ".conv2.1"]
                ops.conv_hw3_s(s_half2, k.tc["hw3"], k.b, k.cout, 1, "ReLU", oscale=k.osc, sout=s_cat2.channels(c2, 2 * c2))
                k = pk[r + ".conv4.0"]
                ops.conv_hw3s2_s(s_cat2.channels(c2, 2 * c2), k.tc["s2"], k.b, k.cout, "ReLU", oscale=k.osc, sout=s_q)
                k = pk[r + ".conv4.1"]
                ops.conv_hw3_s(s_q, k.tc["hw3"], k.b, k.cout, 1, "ReLU", out=lrcat[:, cf:], oscale=k.osc,
                               sout=s_lrcat.channels(cf, cf + c4))
                ops.split_pack(l4, out=s_lrcat.batches(0, B).channels(0, cf))
                ops.split_pack(r4, out=s_lrcat.batches(B, 2 * B).channels(0, cf))
            else:
                self._conv2d(left_image, r + ".conv2.0", 2, out=half2[:B])          # the two images meet in one 2B batch
                self._conv2d(right_image, r + ".conv2.0", 2, out=half2[B:])
                self._conv2d(half2, r + ".conv2.1", out=cat2lr[:, c2:])
                self._conv2d(self._conv2d(cat2lr[:, c2:], r + ".conv4.0", 2), r + ".conv4.1", out=lrcat[:, cf:])
            enc_done = torch.cuda.Event()
            enc_done.record(side)
            if sfmt:
                pk, h = self._pk, h_dec
                k = pk[r + ".fuse.0"]
                ops.conv_hw3_s(s_lrcat.batches(0, B), k.tc["hw3"], k.b, k.cout, 1, "ReLU", half=h, oscale=k.osc, sout=s_f0)
                k = pk[r + ".fuse.1"]
                ops.conv_hw3_s(s_f0, k.tc["hw3"], k.b, k.cout, 1, "ReLU", half=h, oscale=k.osc, sout=s_f1)
                k = pk[r + ".deconv4"]
                up = s_cat2.batches(0, B).channels(0, c2)
                ops.deconv_hw_s(s_f1, k.tc["dc"], k.b, k.cout, "ReLU", half=h, oscale=k.osc, sout=up.hi() if single else up)
                k = pk[r + ".concat"]
                ops.conv_hw3_s(s_cat2.batches(0, B), k.tc["hw3"], k.b, k.cout, 1, "ReLU", half=h, oscale=k.osc, sout=s_cc)
                k = pk[r + ".deconv2"]
                ops.deconv_hw_s(s_cc, k.tc["dc"], k.b, k.cout, None, out=logits, half=h, oscale=k.osc)
                dec_done = torch.cuda.Event()
                dec_done.record(side)

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
        centre = self._inject["coarse_disp"].contiguous() if self._inject and "coarse_disp" in self._inject else d_c
        low_c, high_c = ops.range_samples(centre, DISP_RANGE, samples, n_lm)
        d_f, c_f, o_f, s_f = self._memory_level("fine", l8, r8, samples, prev_info, False)

        # ---- precise (1/4): UNet encoder features concatenated to the backbone features
        if side is not main:
            main.wait_event(enc_done)
            for t in (left_image, right_image, l4, r4):     # read on the side stream: keep the allocator from reusing them early
                t.record_stream(side)

        samples_p = torch.empty((B, 5, H4, W4), device=dev, dtype=torch.float32)
        centre = self._inject["fine_disp"].contiguous() if self._inject and "fine_disp" in self._inject else d_f
        low_f, high_f = ops.range_samples(centre, DISP_RANGE, samples_p, 0)
        vol = self._init3d(lcat, rcat, samples_p, "precise.init3d", out_fmt="s",
                           s_left=s_lrcat.batches(0, B) if sfmt else None, s_right=s_lrcat.batches(B, 2 * B) if sfmt else None)
        d_p, c_p, o_p, top_disp, top_cost = self._heads_predict(vol, samples_p, "precise.pred_heads",
                                                                float(self.levels["precise"]["delta"]), True)
        if sfmt:
            if side is not main:
                main.wait_event(dec_done)
        else:
            f = self._conv2d(self._conv2d(lcat, r + ".fuse.0", single=True), r + ".fuse.1", single=True)
            self._deconv_hw(f, self._pk[r + ".deconv4"], 4, "ReLU", out=cat2[:, :c2], single=True)
            f = self._conv2d(cat2, r + ".concat", single=True)
            logits = self._deconv_hw(f, self._pk[r + ".deconv2"], 4, single=True)
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

    def forward(self, left_feats: List[torch.Tensor], right_feats: List[torch.Tensor], left_image: torch.Tensor,
                right_image: torch.Tensor, prev_info: Optional[dict] = None):
        if prev_info is None:
            prev_info = {}
        if torch.is_grad_enabled() and any(t.requires_grad for t in list(left_feats) + list(right_feats)):
            raise NotImplementedError("libtstereo TEMPORALSTEREO has no backward: call it under torch.no_grad() or with "
                                      "detached features (the reference runs its history frames the same way, "
                                      "projects/TemporalStereo/TemporalStereo.py:268-274)")
        return self._forward(left_feats, right_feats, left_image, right_image, prev_info)


def build_aggregation(cfg) -> nn.Module:
    """Reference `build_aggregation` (architecture/modeling/aggregation/builder.py:12-20)."""
    name = cfg["MODEL"]["AGGREGATION"]["NAME"] if isinstance(cfg, dict) else cfg.MODEL.AGGREGATION.NAME
    return AGGREGATION_REGISTRY.get(name)(cfg)
