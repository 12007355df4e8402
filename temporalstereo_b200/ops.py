"""Tensor-level wrappers of the C ABI (include/tstereo.h).

Each function checks device / dtype / contiguity, allocates the outputs, and launches the kernel on
torch's *current* CUDA stream.  PyTorch is only the allocator and stream provider here; all
arithmetic happens inside libtstereo.so.  Names and argument meaning follow the reference operators
they replace (cited per function, paths relative to the reference repository).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib

ACT = {None: 0, "none": 0, "SiLU": 1, "silu": 1, "ReLU": 2, "relu": 2}


def _stream() -> int:
    """The current stream of the current device.  libtstereo keeps per-process launch state for ONE device (function
    attributes, the zero-bias buffer), so every tensor must live on the current device (`_chk` / `_view5` enforce it)."""
    return torch.cuda.current_stream().cuda_stream


def _same_device(t: torch.Tensor) -> None:
    if t.device.index != torch.cuda.current_device():
        raise ValueError(f"libtstereo launches on the current device (cuda:{torch.cuda.current_device()}), got a tensor on "
                         f"{t.device}: wrap the call in torch.cuda.device(...)")


def _out(out: Optional[torch.Tensor], shape, like: torch.Tensor) -> torch.Tensor:
    """Allocate the output, or check that a caller-provided view has exactly the computed geometry (a wrong-sized
    `out` would be written out of bounds by the kernel)."""
    if out is None:
        return torch.empty(shape, device=like.device, dtype=torch.float32)
    if tuple(out.shape) != tuple(shape):
        raise ValueError(f"out has shape {tuple(out.shape)}, the operator produces {tuple(shape)}")
    return out


def _chk(*ts: Optional[torch.Tensor]) -> None:
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise TypeError("libtstereo ops need CUDA tensors (there is no CPU fallback)")
        _same_device(t)
        if t.dtype != torch.float32:
            raise TypeError(f"libtstereo ops are fp32-only, got {t.dtype}")
        if not t.is_contiguous():
            raise ValueError("libtstereo ops need contiguous tensors")


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _view5(t: torch.Tensor) -> Tuple[int, int, int]:
    """(sB, sC, sD) element strides of a [B,C,D,H,W] (or [B,C,H,W]) view whose (H,W) plane is dense."""
    if t.dtype != torch.float32 or not t.is_cuda:
        raise TypeError("libtstereo ops need fp32 CUDA tensors")
    _same_device(t)
    if t.dim() == 4:
        B, Cc, H, W = t.shape
        sB, sC, sH, sW = t.stride()
        sD = 0
    else:
        B, Cc, D, H, W = t.shape
        sB, sC, sD, sH, sW = t.stride()
    if W > 1 and sW != 1 or H > 1 and sH != W:
        raise ValueError("the (H, W) plane of a libtstereo activation view must be dense")
    return sB, sC, sD


# --------------------------------------------------------------------------- cost volume
def block_cost(left: torch.Tensor, right: torch.Tensor, disp_sample, block_cost_scale: int = 3) -> torch.Tensor:
    """Drop-in for `block_cost(left, right, disp_sample, block_cost_scale)`
    (architecture/modeling/aggregation/utils/block_cost.py:16-83): int `disp_sample` -> shifted
    difference volume, tensor `disp_sample` [B,S,H,W] -> [L, warp(R)] concat volume; both followed by
    the three pooled group-wise terms."""
    _chk(left, right)
    if block_cost_scale != 3:
        raise ValueError("libtstereo block_cost implements block_cost_scale = 3 (every shipped config)")
    B, Cc, H, W = left.shape
    assert right.shape == left.shape, "left / right feature shapes differ"
    G = Cc // 8
    if isinstance(disp_sample, int):
        D = disp_sample
        out = torch.empty((B, Cc + 3 * G, D, H, W), device=left.device, dtype=torch.float32)
        n = _lib.load().tstereo_block_cost_scratch_floats(B, Cc, H, W, D)
        scratch = torch.empty((max(int(n), 1),), device=left.device, dtype=torch.float32)
        _lib.call("tstereo_block_cost_shift", _p(left), _p(right), _p(out), _p(scratch), B, Cc, H, W, D, _stream())
    else:
        _chk(disp_sample)
        S = disp_sample.shape[1]
        assert disp_sample.shape == (B, S, H, W)
        out = torch.empty((B, 2 * Cc + 3 * G, S, H, W), device=left.device, dtype=torch.float32)
        n = _lib.load().tstereo_block_cost_scratch_floats(B, Cc, H, W, S)
        scratch = torch.empty((max(int(n), 1),), device=left.device, dtype=torch.float32)
        _lib.call("tstereo_block_cost_warp", _p(left), _p(right), _p(disp_sample), _p(out), _p(scratch),
                  B, Cc, H, W, S, _stream())
    return out


def group_cost(left: torch.Tensor, right: torch.Tensor, disp_sample) -> torch.Tensor:
    """Only the three pooled group-wise terms of `block_cost` (block_cost.py:6-13, 64-78): [B, 3C/8, D, H, W].
    Side input of the fused cost -> first-conv path (`cost_conv_warp` / `cost_conv_shift`)."""
    _chk(left, right)
    B, Cc, H, W = left.shape
    assert right.shape == left.shape, "left / right feature shapes differ"
    G = Cc // 8
    D = disp_sample if isinstance(disp_sample, int) else disp_sample.shape[1]
    out = torch.empty((B, 3 * G, D, H, W), device=left.device, dtype=torch.float32)
    n = _lib.load().tstereo_block_cost_scratch_floats(B, Cc, H, W, D)
    scratch = torch.empty((max(int(n), 1),), device=left.device, dtype=torch.float32)
    if isinstance(disp_sample, int):
        _lib.call("tstereo_group_cost_shift", _p(left), _p(right), _p(out), _p(scratch), B, Cc, H, W, D, _stream())
    else:
        _chk(disp_sample)
        assert disp_sample.shape == (B, D, H, W)
        _lib.call("tstereo_group_cost_warp", _p(left), _p(right), _p(disp_sample), _p(out), _p(scratch), B, Cc, H, W, D,
                  _stream())
    return out


def block_cost_shift_s(left: torch.Tensor, right: torch.Tensor, D: int) -> "Split":
    """The shift cost volume of `block_cost(left, right, D)` as an S-format tensor [B, C + 3C/8, D, H, W] (fp16 hi / lo split
    of the fp32 volume), the layout the TMA-fed first conv reads: the C cost planes come straight from the cost kernel, the
    3C/8 group-wise planes through `split_pack` of the compact group volume.  C/8 + 3C/64 chunks (C a multiple of 64)."""
    _chk(left, right)
    B, Cc, H, W = left.shape
    assert right.shape == left.shape, "left / right feature shapes differ"
    G = Cc // 8
    if (3 * G) % 8:
        raise ValueError(f"block_cost_shift_s needs C a multiple of 64 (3C/8 group planes must fill whole chunks), got C={Cc}")
    vol = Split(B, Cc + 3 * G, D, H, W, 2, device=left.device)
    gv = torch.empty((B, 3 * G, D, H, W), device=left.device, dtype=torch.float32)
    n = _lib.load().tstereo_block_cost_scratch_floats(B, Cc, H, W, D)
    scratch = torch.empty((max(int(n), 1),), device=left.device, dtype=torch.float32)
    keep, ref = _sref(vol.channels(0, Cc))
    _lib.call("tstereo_block_cost_shift_s", _p(left), _p(right), ref, _p(gv), _p(scratch), B, Cc, H, W, D, _stream())
    split_pack(gv, out=vol.channels(Cc, Cc + 3 * G))
    return vol


def cost_conv_warp(right: torch.Tensor, samples: torch.Tensor, gvol: torch.Tensor, addL: Optional[torch.Tensor],
                   wpack: torch.Tensor, bias: Optional[torch.Tensor], cout: int, act=None,
                   out: Optional[torch.Tensor] = None, half: bool = False, oscale: Optional[torch.Tensor] = None) -> torch.Tensor:
    """act(conv(1,3,3)(block_cost(left, right, samples)) + bias) without the volume (block_cost.py:47-81 ->
    module.py:111-147): `wpack` is the tensor-core image of W[:, C:] over [warp(R) | group terms] and `addL`
    [B, cout, H, W] = conv3x3(left, W[:, :C]) is the candidate-invariant left half."""
    _chk(right, samples, gvol, addL, wpack, bias, oscale)
    B, Cc, H, W = right.shape
    S = samples.shape[1]
    assert samples.shape == (B, S, H, W) and gvol.shape == (B, 3 * (Cc // 8), S, H, W)
    assert addL is None or addL.shape == (B, cout, H, W)
    out = _out(out, (B, cout, S, H, W), right)
    osB, osC, osD = _view5(out)
    assert wpack.numel() == _lib.load().tstereo_cost_conv_wpack_floats(Cc, cout, int(half))
    _lib.call("tstereo_cost_conv_warp", _p(right), _p(samples), _p(gvol), _p(addL), _p(out), osB, osC, osD, _p(wpack),
              _p(bias), _p(oscale), B, Cc, cout, S, H, W, ACT[act], int(half), _stream())
    return out


def cost_conv_shift(left: torch.Tensor, right: torch.Tensor, gvol: torch.Tensor, wpack: torch.Tensor,
                    bias: Optional[torch.Tensor], cout: int, act=None, out: Optional[torch.Tensor] = None,
                    half: bool = False, oscale: Optional[torch.Tensor] = None) -> torch.Tensor:
    """act(conv(1,3,3)(block_cost(left, right, D)) + bias) without the volume (block_cost.py:34-45, 64-81 ->
    module.py:111-147); `wpack` is the tensor-core image of the whole W over [-(L - R_d)^2 | group terms]."""
    _chk(left, right, gvol, wpack, bias, oscale)
    B, Cc, H, W = left.shape
    D = gvol.shape[2]
    assert right.shape == left.shape and gvol.shape == (B, 3 * (Cc // 8), D, H, W)
    out = _out(out, (B, cout, D, H, W), left)
    osB, osC, osD = _view5(out)
    assert wpack.numel() == _lib.load().tstereo_cost_conv_wpack_floats(Cc, cout, int(half))
    _lib.call("tstereo_cost_conv_shift", _p(left), _p(right), _p(gvol), _p(out), osB, osC, osD, _p(wpack), _p(bias),
              _p(oscale), B, Cc, cout, D, H, W, ACT[act], int(half), _stream())
    return out


# --------------------------------------------------------------------------- convolutions
def tap_projection_weights(w_r: torch.Tensor) -> torch.Tensor:
    """Right-feature part of a warp level's first-conv weights [Cout, C, 9] -> the [9*Cout, C, 1] weights of the 1x1 conv
    whose output is T[t*Cout + co] = sum_c w[co, c, t] * R[c] (see `cost_taps`)."""
    cout, C, T = w_r.shape
    assert T == 9
    return w_r.permute(2, 0, 1).reshape(9 * cout, C, 1).contiguous()


def cost_taps(T: torch.Tensor, samples: torch.Tensor, gconv: Optional[torch.Tensor], addL: Optional[torch.Tensor],
              bias: torch.Tensor, cout: int, act=None, out: Optional[torch.Tensor] = None, sout: Optional["Split"] = None,
              want_f32: bool = False):
    """Tap-projection form of the warp levels' first (1,3,3) conv: act(bias + addL + gconv + sum over the 3x3 neighbours of
    the x-lerp of T at the neighbour's warp taps) — T [B, 9*cout, H, W] is the 1x1 projection of the right features
    (`tap_projection_weights`), the same for every candidate.  Returns (fp32 [B, cout, S, H, W] or None, sout)."""
    B, S, H, W = samples.shape
    _chk(T, samples, gconv, addL, bias)
    if tuple(T.shape) != (B, 9 * cout, H, W):
        raise ValueError(f"T has shape {tuple(T.shape)}, expected {(B, 9 * cout, H, W)}")
    if gconv is not None and tuple(gconv.shape) != (B, cout, S, H, W):
        raise ValueError(f"gconv has shape {tuple(gconv.shape)}, expected {(B, cout, S, H, W)}")
    if addL is not None and tuple(addL.shape) != (B, cout, H, W):
        raise ValueError(f"addL has shape {tuple(addL.shape)}, expected {(B, cout, H, W)}")
    if sout is not None and sout.shape != (B, cout, S, H, W):
        raise ValueError(f"S-format output has shape {sout.shape}, the operator produces {(B, cout, S, H, W)}")
    if out is None and (want_f32 or sout is None):
        out = torch.empty((B, cout, S, H, W), device=T.device, dtype=torch.float32)
    osB = osC = osD = 0
    if out is not None:
        if tuple(out.shape) != (B, cout, S, H, W):
            raise ValueError(f"out has shape {tuple(out.shape)}, the operator produces {(B, cout, S, H, W)}")
        osB, osC, osD = _view5(out)
    keep, ref = _sref(sout)
    _lib.call("tstereo_cost_taps", _p(T), _p(samples), _p(gconv), _p(addL), _p(bias), _p(out), osB, osC, osD, ref,
              B, cout, S, H, W, ACT[act], _stream())
    return out, sout


def conv_hw3(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], cout: int, stride: int = 1,
             dilation: int = 1, act=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(1,3,3) / 3x3 conv, padding = dilation, packed weights w[Cin][9][CoutP]
    (layers/basic_layers.py:194-235 with eval-mode BN folded)."""
    five = x.dim() == 5
    B, Cin = x.shape[:2]
    D = x.shape[2] if five else 1
    Hin, Win = x.shape[-2:]
    Hout, Wout = (Hin - 1) // stride + 1, (Win - 1) // stride + 1
    out = _out(out, (B, cout, D, Hout, Wout) if five else (B, cout, Hout, Wout), x)
    isB, isC, isD = _view5(x)
    osB, osC, osD = _view5(out)
    _chk(w, bias)
    _lib.call("tstereo_conv_hw3", _p(x), isB, isC, isD, _p(out), osB, osC, osD, _p(w), _p(bias),
              B, Cin, cout, D, Hin, Win, Hout, Wout, stride, dilation, ACT[act], _stream())
    return out


def tf32_split(w: torch.Tensor):
    """w = hi + lo with both parts exactly representable in tf32 (10 mantissa bits), rounding to
    nearest / ties away like cvt.rna.tf32.f32."""
    def rna(t):
        i = t.contiguous().view(torch.int32)
        return ((i + 0x1000) & ~0x1FFF).view(torch.float32)
    hi = rna(w)
    return hi, rna(w - hi)


def fp16_prescale(w: torch.Tensor):
    """Per-output-channel power-of-two scaling of BN-folded weights [Cout, Cin, taps] for the fp16 hi+lo operand split:
    returns (w * s, 1 / s) with max|w[c] * s[c]| in (511, 1023].  The lo half of a small weight would otherwise fall into
    fp16's subnormal range (|w| = 1e-3 keeps 14 of 22 bits); scaled, hi + lo carries 22 bits for every weight within
    2^-13 of its channel's largest.  The kernel multiplies the fp32 accumulator by 1 / s (exact) before the bias
    (`oscale` of the tensor-core operators)."""
    m = w.abs().flatten(1).amax(1).clamp_min(1e-30)
    s = torch.exp2(torch.floor(torch.log2(1023.0 / m)))
    return w * s.view(-1, *([1] * (w.dim() - 1))), (1.0 / s).contiguous()


def _pack_tc2_group(w: torch.Tensor, nky: int = 3, half: bool = False, fold: int = 3) -> torch.Tensor:
    """One output-channel group (<= 32): [Cout, Cin, nky*fold] -> [chunk][ky nky][khalf 2][row 2N][16 B],
    row = part*N + kx*CP + co, N = fold*CP (fold = 3 kx taps, or 1), CP = 8|16|32.  tf32 form: chunk = 8 channels, a 16-byte row holds 4 floats,
    parts = tf32 hi / lo.  half form: chunk = 16 channels, a row holds 8 halves, parts = fp16 hi / lo (returned
    reinterpreted as float32 pairs so the ABI keeps one pointer type)."""
    cout, cin, T = w.shape
    assert T == fold * nky and cout <= 32
    CP = 8 if cout <= 8 else 16 if cout <= 16 else 32
    per = 16 if half else 8
    nch = (cin + per - 1) // per
    full = torch.zeros((CP, nch * per, nky, fold), device=w.device, dtype=torch.float32)
    full[:cout, :cin] = w.reshape(cout, cin, nky, fold)
    if half:
        hi = full.half()
        lo = (full - hi.float()).half()
        parts = torch.stack([hi, lo]).view(2, CP, nch, 2, 8, nky, fold)   # [part, co, chunk, khalf, i, ky, kx]
        return parts.permute(2, 5, 3, 0, 6, 1, 4).contiguous().view(-1).view(torch.float32)
    hi, lo = tf32_split(full)
    parts = torch.stack([hi, lo]).view(2, CP, nch, 2, 4, nky, fold)       # [part, co, chunk, khalf, i, ky, kx]
    return parts.permute(2, 5, 3, 0, 6, 1, 4).contiguous().view(-1)       # [chunk, ky, khalf, part, kx, co, i]


def pack_conv_d_tc2(w: torch.Tensor, half: bool = False) -> torch.Tensor:
    """(k,1,1) conv along D, w [Cout, Cin, k] -> operand image of tstereo_conv_d_tc2: the k input planes are stacked
    on the channel axis (virtual channel = tap*Cin8 + c) of a 1x1 conv (one ky tap).  fp16 form: single column block
    (N = CP); tf32 form: the kx-folded layout with the weights in the kx = 0 block."""
    cout, cin, k = w.shape
    cin8 = (cin + 7) // 8 * 8
    fold = 1 if half else 3
    virt = torch.zeros((cout, k, cin8, 1, fold), device=w.device, dtype=torch.float32)
    virt[:, :, :cin, 0, 0] = w.permute(0, 2, 1)
    virt = virt.reshape(cout, k * cin8, fold)
    return torch.cat([_pack_tc2_group(virt[c0:c0 + 32], 1, half, fold) for c0 in range(0, cout, 32)])


def pack_conv_hw3_tc2(w: torch.Tensor, half: bool = False) -> torch.Tensor:
    """[Cout, Cin, 9] (BN folded, taps ky*3+kx) -> the B-operand image of tstereo_conv_hw3_tc2, output channels
    in groups of 32."""
    return torch.cat([_pack_tc2_group(w[c0:c0 + 32], 3, half) for c0 in range(0, w.shape[0], 32)])


def virtual_weights_s2(w: torch.Tensor) -> torch.Tensor:
    """Stride-2 3x3 conv (padding 1), w [Cout, Cin, 9] -> the weights [Cout, 4*Cin8, 3, 3] of the equivalent stride-1
    3x3 conv (padding 1) over the four input parity phases stacked on the channel axis (virtual channel =
    (row parity*2 + col parity)*Cin8 + c, phase image P[pr][pc](m, n) = x(2m + pr, 2n + pc)):
    tap k reads input 2*o + k - 1 = phase (k+1)%2 at o + {-1, 0, 0}[k]."""
    cout, cin, T = w.shape
    assert T == 9
    cin8 = (cin + 7) // 8 * 8
    w4 = w.reshape(cout, cin, 3, 3)
    virt = torch.zeros((cout, 4, cin8, 3, 3), device=w.device, dtype=torch.float32)
    tap = {0: (1, -1), 1: (0, 0), 2: (1, 0)}                      # k -> (parity, offset)
    for ky, (pr, dm) in tap.items():
        for kx, (pc, dn) in tap.items():
            virt[:, pr * 2 + pc, :cin, dm + 1, dn + 1] = w4[:, :, ky, kx]
    return virt.reshape(cout, 4 * cin8, 3, 3)


def pack_conv_hw3s2_tc2(w: torch.Tensor, half: bool = False) -> torch.Tensor:
    """Operand image of tstereo_conv_hw3s2_tc2 (see virtual_weights_s2)."""
    v = virtual_weights_s2(w)
    return pack_conv_hw3_tc2(v.reshape(v.shape[0], v.shape[1], 9), half)


def virtual_weights_deconv(w: torch.Tensor, k: int) -> torch.Tensor:
    """Transposed conv (stride 2, padding 1; k=3 with output_padding 1, or k=4), w [Cout, Cin, k*k] in the
    transposed-conv tap order (out[2i - 1 + t] += in[i] * w[t]) -> [4 (py*2+px), Cout, Cin, 3, 3]: the 3x3 shift
    kernel (stride-1 conv, padding 1) of each output parity phase: out[2m + p] = sum_d in[m + d] * w[t(p, d)]."""
    cout, cin, T = w.shape
    assert T == k * k and k in (3, 4)
    w4 = w.reshape(cout, cin, k, k)
    shifts = ({0: {0: 1}, 1: {1: 0, 0: 2}} if k == 3 else {0: {0: 1, -1: 3}, 1: {1: 0, 0: 2}})   # parity -> {shift: tap}
    out = torch.zeros((4, cout, cin, 3, 3), device=w.device, dtype=torch.float32)
    for py in (0, 1):
        for px in (0, 1):
            for dy, ky in shifts[py].items():
                for dx, kx in shifts[px].items():
                    out[py * 2 + px, :, :, dy + 1, dx + 1] = w4[:, :, ky, kx]
    return out


def pack_deconv_hw_tc2(w: torch.Tensor, k: int, half: bool = False) -> torch.Tensor:
    """Operand image of tstereo_deconv_hw_tc2: the four phase kernels of virtual_weights_deconv, packed one after
    the other."""
    v = virtual_weights_deconv(w, k)
    return torch.cat([pack_conv_hw3_tc2(v[ph].reshape(v.shape[1], v.shape[2], 9), half) for ph in range(4)])


def conv_hw3_tc2(x: torch.Tensor, wpack: torch.Tensor, bias: Optional[torch.Tensor], cout: int, dilation: int = 1,
                 act=None, out: Optional[torch.Tensor] = None, half: bool = False,
                 oscale: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Stride-1 (1,3,3) / 3x3 conv on the tensor cores (kx-folded tcgen05 kernel, hi+lo split operands).  `oscale` [Cout]:
    per-output-channel multiplier applied to the accumulator before the bias (1 / the weight pre-scale of `fp16_prescale`)."""
    five = x.dim() == 5
    B, Cin = x.shape[:2]
    D = x.shape[2] if five else 1
    H, W = x.shape[-2:]
    out = _out(out, (B, cout, D, H, W) if five else (B, cout, H, W), x)
    isB, isC, isD = _view5(x)
    osB, osC, osD = _view5(out)
    _chk(wpack, bias, oscale)
    assert wpack.numel() == _lib.load().tstereo_conv_hw3_tc2_wpack_floats(Cin, cout, int(bool(half)))
    _lib.call("tstereo_conv_hw3_tc2", _p(x), isB, isC, isD, _p(out), osB, osC, osD, _p(wpack), _p(bias), _p(oscale),
              B, Cin, cout, D, H, W, dilation, ACT[act], int(half), _stream())
    return out


def conv_hw3s2_tc2(x: torch.Tensor, wpack: torch.Tensor, bias: Optional[torch.Tensor], cout: int, act=None,
                   out: Optional[torch.Tensor] = None, half: bool = False, oscale: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Stride-2 (1,3,3) / 3x3 conv, padding 1, on the tensor cores (phase-decomposed input, 3xTF32)."""
    five = x.dim() == 5
    B, Cin = x.shape[:2]
    D = x.shape[2] if five else 1
    Hin, Win = x.shape[-2:]
    H, W = (Hin - 1) // 2 + 1, (Win - 1) // 2 + 1
    out = _out(out, (B, cout, D, H, W) if five else (B, cout, H, W), x)
    isB, isC, isD = _view5(x)
    osB, osC, osD = _view5(out)
    _chk(wpack, bias, oscale)
    assert wpack.numel() == _lib.load().tstereo_conv_hw3s2_tc2_wpack_floats(Cin, cout, int(bool(half)))
    _lib.call("tstereo_conv_hw3s2_tc2", _p(x), isB, isC, isD, _p(out), osB, osC, osD, _p(wpack), _p(bias), _p(oscale),
              B, Cin, cout, D, Hin, Win, ACT[act], int(half), _stream())
    return out


def deconv_hw_tc2(x: torch.Tensor, wpack: torch.Tensor, bias: Optional[torch.Tensor], cout: int, act=None,
                  out: Optional[torch.Tensor] = None, half: bool = False, oscale: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Transposed (1,k,k)/kxk conv, stride 2 (Hout = 2*Hin), on the tensor cores: one output parity phase per launch."""
    five = x.dim() == 5
    B, Cin = x.shape[:2]
    D = x.shape[2] if five else 1
    Hin, Win = x.shape[-2:]
    out = _out(out, (B, cout, D, 2 * Hin, 2 * Win) if five else (B, cout, 2 * Hin, 2 * Win), x)
    isB, isC, isD = _view5(x)
    osB, osC, osD = _view5(out)
    _chk(wpack, bias, oscale)
    assert wpack.numel() == _lib.load().tstereo_deconv_hw_tc2_wpack_floats(Cin, cout, int(bool(half)))
    _lib.call("tstereo_deconv_hw_tc2", _p(x), isB, isC, isD, _p(out), osB, osC, osD, _p(wpack), _p(bias), _p(oscale),
              B, Cin, cout, D, Hin, Win, ACT[act], int(half), _stream())
    return out


def conv_d_tc2(x: torch.Tensor, wpack: torch.Tensor, bias: Optional[torch.Tensor], cout: int, k: int = 3, stride: int = 1,
               dilation: int = 1, transposed: bool = False, act=None, out: Optional[torch.Tensor] = None,
               half: bool = False, oscale: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(k,1,1) conv along D (or its stride-2 transposed form) through the tensor-core kernel."""
    B, Cin, Din, H, W = x.shape
    Dout = 2 * Din if transposed else (Din - 1) // stride + 1
    out = _out(out, (B, cout, Dout, H, W), x)
    isB, isC, isD = _view5(x)
    osB, osC, osD = _view5(out)
    _chk(wpack, bias, oscale)
    assert wpack.numel() == _lib.load().tstereo_conv_d_tc2_wpack_floats(Cin, cout, k, int(half))
    _lib.call("tstereo_conv_d_tc2", _p(x), isB, isC, isD, _p(out), osB, osC, osD, _p(wpack), _p(bias), _p(oscale),
              B, Cin, cout, Din, Dout, H, W, k, stride, dilation, int(transposed), ACT[act], int(half), _stream())
    return out


# --------------------------------------------------------------------------- S-format (TMA-fed) convolutions
class Split:
    """S-format activation (include/tstereo.h `tstereo_split`): the fp16 hi / lo halves of an fp32 tensor [B, C, D, H, W],
    stored [B][D][part][C8][H][W][8] so that the tensor-core convolutions stage a K-chunk with one TMA box.  `parts` = 1
    keeps the hi half only (operands of the single-term form).  `t` is the backing fp16 tensor (possibly a chunk slice
    of a larger one: channel concatenation is a slice of the C8 axis)."""

    def __init__(self, B: int, C: int, D: int, H: int, W: int, parts: int = 2, device=None, t: Optional[torch.Tensor] = None,
                 five: bool = True):
        self.C, self.five = C, five
        c8 = (C + 7) // 8
        if t is None:
            t = torch.empty((B, D, parts, c8, H, W, 8), device=device, dtype=torch.float16)
        if tuple(t.shape) != (B, D, parts, c8, H, W, 8) or t.dtype != torch.float16:
            raise ValueError(f"S-format backing tensor has shape {tuple(t.shape)}, expected {(B, D, parts, c8, H, W, 8)}")
        if not t.is_cuda:
            raise TypeError("libtstereo ops need CUDA tensors (there is no CPU fallback)")
        _same_device(t)
        if t.stride()[4:] != (W * 8, 8, 1):
            raise ValueError("the [H][W][8] block of an S-format tensor must be dense")
        self.t = t

    @property
    def shape(self):
        B, D, _, _, H, W, _ = self.t.shape
        return (B, self.C, D, H, W) if self.five else (B, self.C, H, W)

    @property
    def parts(self) -> int:
        return self.t.shape[2]

    def channels(self, c0: int, c1: int) -> "Split":
        """Channels [c0, c1) (multiples of 8) as an S-format view: the operand of a concatenation."""
        assert c0 % 8 == 0 and (c1 % 8 == 0 or c1 == self.C) and 0 <= c0 < c1 <= self.C
        B, D, P_, _, H, W, _ = self.t.shape
        return Split(B, c1 - c0, D, H, W, P_, t=self.t[:, :, :, c0 // 8:(c1 + 7) // 8], five=self.five)

    def batches(self, b0: int, b1: int) -> "Split":
        B, D, P_, _, H, W, _ = self.t.shape
        return Split(b1 - b0, self.C, D, H, W, P_, t=self.t[b0:b1], five=self.five)

    def hi(self) -> "Split":
        """The hi half alone (what a single-term consumer reads, what a single-term producer needs to write)."""
        B, D, _, _, H, W, _ = self.t.shape
        return Split(B, self.C, D, H, W, 1, t=self.t[:, :, :1], five=self.five)

    def struct(self, nb: int = 0) -> "_lib.SplitStruct":
        sB, sD, sP, sC8 = self.t.stride()[:4]
        return _lib.SplitStruct(self.t.data_ptr(), sB, sD, sP if self.parts == 2 else 0, sC8, self.t.shape[3], self.parts, nb)

    def float(self) -> torch.Tensor:
        """hi (+ lo) as an fp32 [B, C, (D,) H, W] tensor (tests / debugging)."""
        v = self.t.float().sum(2)                                  # [B, D, C8, H, W, 8]
        B, D, C8, H, W, _ = v.shape
        v = v.permute(0, 2, 5, 1, 3, 4).reshape(B, C8 * 8, D, H, W)[:, :self.C]
        return v.contiguous() if self.five else v[:, :, 0].contiguous()


def _sref(s: Optional[Split], nb: int = 0):
    import ctypes
    if s is None:
        return None, None
    st = s.struct(nb)
    return st, ctypes.byref(st)


def split_pack(x: torch.Tensor, parts: int = 2, out: Optional[Split] = None) -> Split:
    """fp32 [B, C, (D,) H, W] view -> S-format."""
    five = x.dim() == 5
    B, Cc = x.shape[:2]
    D = x.shape[2] if five else 1
    H, W = x.shape[-2:]
    if out is None:
        out = Split(B, Cc, D, H, W, parts, device=x.device, five=five)
    if out.shape != tuple(x.shape):
        raise ValueError(f"S-format output has shape {out.shape}, the input {tuple(x.shape)}")
    isB, isC, isD = _view5(x)
    keep, ref = _sref(out)
    _lib.call("tstereo_split_pack", _p(x), isB, isC, isD, ref, B, Cc, D, H, W, _stream())
    return out


def _s_io(x, out, sout, out_shape, want_f32, nb=0):
    """Common argument handling of the `_s` operators: x is a torch tensor or a Split; returns the ABI pieces."""
    xs = x if isinstance(x, Split) else None
    if xs is None:
        isB, isC, isD = _view5(x)
        xp = _p(x)
        like = x
    else:
        isB = isC = isD = 0
        xp = None
        like = xs.t
    if sout is not None:
        # nb: only the leading nb batches are written, so the S-format output may hold fewer than B (but at least nb)
        if sout.shape[1:] != tuple(out_shape[1:]) or not (sout.shape[0] == out_shape[0] or (nb and sout.shape[0] >= nb)):
            raise ValueError(f"S-format output has shape {sout.shape}, the operator produces {tuple(out_shape)}"
                             + (f" (batches [0, {nb}))" if nb else ""))
    if out is None and (want_f32 or sout is None):
        out = torch.empty(out_shape, device=like.device, dtype=torch.float32)
    if out is not None:
        if tuple(out.shape) != tuple(out_shape):
            raise ValueError(f"out has shape {tuple(out.shape)}, the operator produces {tuple(out_shape)}")
        osB, osC, osD = _view5(out)
    else:
        osB = osC = osD = 0
    return xs, xp, (isB, isC, isD), out, (osB, osC, osD)


def conv_hw3_s(x, wpack: torch.Tensor, bias: Optional[torch.Tensor], cout: int, dilation: int = 1, act=None,
               out: Optional[torch.Tensor] = None, half: int = 1, oscale: Optional[torch.Tensor] = None,
               sout: Optional[Split] = None, want_f32: bool = False, nb: int = 0):
    """`conv_hw3_tc2` with S-format operands: `x` may be a Split (staged by TMA), `sout` a Split written by the epilogue
    (batches [0, nb) only when nb > 0); the fp32 output is produced when `out` is given, `want_f32`, or there is no `sout`.
    Returns (fp32 output or None, sout)."""
    shp = x.shape
    five = len(shp) == 5
    B, Cin = shp[:2]
    D = shp[2] if five else 1
    H, W = shp[-2:]
    xs, xp, (isB, isC, isD), out, (osB, osC, osD) = _s_io(x, out, sout, (B, cout, D, H, W) if five else (B, cout, H, W), want_f32, nb)
    _chk(wpack, bias, oscale)
    assert wpack.numel() == _lib.load().tstereo_conv_hw3_tc2_wpack_floats(Cin, cout, 1)
    k1, r1 = _sref(xs)
    k2, r2 = _sref(sout, nb)
    _lib.call("tstereo_conv_hw3_s", xp, isB, isC, isD, r1, _p(out), osB, osC, osD, r2, _p(wpack), _p(bias), _p(oscale),
              B, Cin, cout, D, H, W, dilation, ACT[act], int(half), _stream())
    return out, sout


def conv_hw3s2_s(x, wpack: torch.Tensor, bias: Optional[torch.Tensor], cout: int, act=None, out: Optional[torch.Tensor] = None,
                 half: int = 1, oscale: Optional[torch.Tensor] = None, sout: Optional[Split] = None, want_f32: bool = False,
                 nb: int = 0):
    """`conv_hw3s2_tc2` with S-format operands (see conv_hw3_s)."""
    shp = x.shape
    five = len(shp) == 5
    B, Cin = shp[:2]
    D = shp[2] if five else 1
    Hin, Win = shp[-2:]
    H, W = (Hin - 1) // 2 + 1, (Win - 1) // 2 + 1
    xs, xp, (isB, isC, isD), out, (osB, osC, osD) = _s_io(x, out, sout, (B, cout, D, H, W) if five else (B, cout, H, W), want_f32, nb)
    _chk(wpack, bias, oscale)
    assert wpack.numel() == _lib.load().tstereo_conv_hw3s2_tc2_wpack_floats(Cin, cout, 1)
    k1, r1 = _sref(xs)
    k2, r2 = _sref(sout, nb)
    _lib.call("tstereo_conv_hw3s2_s", xp, isB, isC, isD, r1, _p(out), osB, osC, osD, r2, _p(wpack), _p(bias), _p(oscale),
              B, Cin, cout, D, Hin, Win, ACT[act], int(half), _stream())
    return out, sout


def deconv_hw_s(x, wpack: torch.Tensor, bias: Optional[torch.Tensor], cout: int, act=None, out: Optional[torch.Tensor] = None,
                half: int = 1, oscale: Optional[torch.Tensor] = None, sout: Optional[Split] = None, want_f32: bool = False,
                nb: int = 0):
    """`deconv_hw_tc2` with S-format operands (see conv_hw3_s)."""
    shp = x.shape
    five = len(shp) == 5
    B, Cin = shp[:2]
    D = shp[2] if five else 1
    Hin, Win = shp[-2:]
    xs, xp, (isB, isC, isD), out, (osB, osC, osD) = _s_io(
        x, out, sout, (B, cout, D, 2 * Hin, 2 * Win) if five else (B, cout, 2 * Hin, 2 * Win), want_f32, nb)
    _chk(wpack, bias, oscale)
    assert wpack.numel() == _lib.load().tstereo_deconv_hw_tc2_wpack_floats(Cin, cout, 1)
    k1, r1 = _sref(xs)
    k2, r2 = _sref(sout, nb)
    _lib.call("tstereo_deconv_hw_s", xp, isB, isC, isD, r1, _p(out), osB, osC, osD, r2, _p(wpack), _p(bias), _p(oscale),
              B, Cin, cout, D, Hin, Win, ACT[act], int(half), _stream())
    return out, sout


def conv_d_s(x, wpack: torch.Tensor, bias: Optional[torch.Tensor], cout: int, k: int = 3, stride: int = 1, dilation: int = 1,
             transposed: bool = False, act=None, out: Optional[torch.Tensor] = None, half: int = 1,
             oscale: Optional[torch.Tensor] = None, sout: Optional[Split] = None, want_f32: bool = False, nb: int = 0):
    """`conv_d_tc2` with S-format operands (see conv_hw3_s)."""
    B, Cin, Din, H, W = x.shape
    Dout = 2 * Din if transposed else (Din - 1) // stride + 1
    xs, xp, (isB, isC, isD), out, (osB, osC, osD) = _s_io(x, out, sout, (B, cout, Dout, H, W), want_f32, nb)
    _chk(wpack, bias, oscale)
    assert wpack.numel() == _lib.load().tstereo_conv_d_tc2_wpack_floats(Cin, cout, k, 1)
    k1, r1 = _sref(xs)
    k2, r2 = _sref(sout, nb)
    _lib.call("tstereo_conv_d_s", xp, isB, isC, isD, r1, _p(out), osB, osC, osD, r2, _p(wpack), _p(bias), _p(oscale),
              B, Cin, cout, Din, Dout, H, W, k, stride, dilation, int(transposed), ACT[act], int(half), _stream())
    return out, sout


def conv_d(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], cout: int, k: int = 3, stride: int = 1,
           dilation: int = 1, transposed: bool = False, act=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(k,1,1) conv along D (or its stride-2 transposed form), packed weights w[Cin][k][CoutP]."""
    B, Cin, Din, H, W = x.shape
    Dout = 2 * Din if transposed else (Din - 1) // stride + 1
    out = _out(out, (B, cout, Dout, H, W), x)
    isB, isC, isD = _view5(x)
    osB, osC, osD = _view5(out)
    _chk(w, bias)
    _lib.call("tstereo_conv_d", _p(x), isB, isC, isD, _p(out), osB, osC, osD, _p(w), _p(bias),
              B, Cin, cout, Din, Dout, H * W, k, stride, dilation, int(transposed), ACT[act], _stream())
    return out


def deconv_hw(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], cout: int, k: int = 3, act=None,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Transposed (1,k,k)/kxk conv, stride 2, padding 1 (k=3: output_padding 1), w[Cin][k*k][CoutP]."""
    five = x.dim() == 5
    B, Cin = x.shape[:2]
    D = x.shape[2] if five else 1
    Hin, Win = x.shape[-2:]
    out = _out(out, (B, cout, D, 2 * Hin, 2 * Win) if five else (B, cout, 2 * Hin, 2 * Win), x)
    isB, isC, isD = _view5(x)
    osB, osC, osD = _view5(out)
    _chk(w, bias)
    _lib.call("tstereo_deconv_hw", _p(x), isB, isC, isD, _p(out), osB, osC, osD, _p(w), _p(bias),
              B, Cin, cout, D, Hin, Win, k, ACT[act], _stream())
    return out


def copy_planes(x: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """out[...] = x for a dense x [B,C,H,W] and a channel-slice view `out` of a concat buffer (the `torch.cat`s of
    precise.py:86 are never materialised by a library call)."""
    _chk(x)
    B, Cc, H, W = x.shape
    if tuple(out.shape) != tuple(x.shape):
        raise ValueError(f"copy_planes: out {tuple(out.shape)} vs in {tuple(x.shape)}")
    osB, osC, _ = _view5(out)
    _lib.call("tstereo_copy_planes", _p(x), _p(out), osB, osC, B, Cc, H * W, _stream())
    return out


def resize_add_act(a: torch.Tensor, size, skip: Optional[torch.Tensor] = None, act=None) -> torch.Tensor:
    """act(F.interpolate(a, size, 'trilinear', align_corners=True) + skip) (module.py:285-295)."""
    _chk(a, skip)
    B, Cc, Da, Ha, Wa = a.shape
    D, H, W = size
    out = torch.empty((B, Cc, D, H, W), device=a.device, dtype=torch.float32)
    if skip is not None:
        assert tuple(skip.shape) == tuple(out.shape)
    _lib.call("tstereo_resize_add_act", _p(a), _p(skip), _p(out), B, Cc, Da, Ha, Wa, D, H, W, ACT[act], _stream())
    return out


def resize_add_act_s(a: torch.Tensor, size, skip: Optional[torch.Tensor] = None, act=None, parts: int = 2) -> "Split":
    """`resize_add_act` with the result in S-format (the operand layout of the TMA-fed convolutions)."""
    _chk(a, skip)
    B, Cc, Da, Ha, Wa = a.shape
    D, H, W = size
    out = Split(B, Cc, D, H, W, parts, device=a.device)
    if skip is not None:
        assert tuple(skip.shape) == out.shape
    keep, ref = _sref(out)
    _lib.call("tstereo_resize_add_act_s", _p(a), _p(skip), ref, B, Cc, Da, Ha, Wa, D, H, W, ACT[act], _stream())
    return out


def pool5(x: torch.Tensor, avg: torch.Tensor, mx: torch.Tensor) -> None:
    """avg_pool3d / max_pool3d, kernel 5, stride 1, padding 2, written into two views (module.py:416-417)."""
    B, Cc, D, H, W = x.shape
    if tuple(avg.shape) != tuple(x.shape) or tuple(mx.shape) != tuple(x.shape):
        raise ValueError(f"pool5 outputs must have the input's shape {tuple(x.shape)}")
    xsB, xsC, xsD = _view5(x)
    osB, osC, osD = _view5(avg)
    assert (xsD == H * W or D == 1) and (osD == H * W or D == 1) and _view5(mx) == (osB, osC, osD)
    _lib.call("tstereo_pool5", _p(x), xsB, xsC, _p(avg), _p(mx), osB, osC, B, Cc, D, H, W, _stream())


def merge_memory(vol: torch.Tensor, samples: torch.Tensor, mem_sample: Optional[torch.Tensor],
                 mem_cost: Optional[torch.Tensor], past_w: torch.Tensor, past_b: torch.Tensor, M: int = 2,
                 out_vol: Optional[torch.Tensor] = None):
    """Temporal memory merge (coarse.py:84-105, fine.py:104-122): returns (volume [B,C,D+M,H,W], samples)."""
    _chk(vol, samples, mem_sample, mem_cost, past_w, past_b)
    B, Cc, D, H, W = vol.shape
    out_vol = _out(out_vol, (B, Cc, D + M, H, W), vol)
    if tuple(samples.shape) != (B, D, H, W):
        raise ValueError(f"samples has shape {tuple(samples.shape)}, expected {(B, D, H, W)}")
    osB, osC, osD = _view5(out_vol)
    assert osD == H * W
    out_s = torch.empty((B, D + M, H, W), device=vol.device, dtype=torch.float32)
    _lib.call("tstereo_merge_memory", _p(vol), _p(samples), _p(mem_sample), _p(mem_cost), _p(past_w), _p(past_b),
              _p(out_vol), osB, osC, _p(out_s), B, Cc, D, M, H, W, _stream())
    return out_vol, out_s


def heads(feat: torch.Tensor, w: torch.Tensor, delta: float):
    """Final (1,3,3) C->1 convs of both prediction heads + tanh offset squashing (module.py:380-398)."""
    _chk(feat, w)
    B, C2, D, H, W = feat.shape
    cost = torch.empty((B, D, H, W), device=feat.device, dtype=torch.float32)
    off = torch.empty_like(cost)
    _lib.call("tstereo_heads", _p(feat), _p(w), _p(cost), _p(off), B, C2 // 2, D, H, W, float(delta), _stream())
    return cost, off


def predict_disp(cost: torch.Tensor, samples: torch.Tensor, off: torch.Tensor, want_top: bool = False):
    """top-2 soft-argmin (coarse.py:69-75): disp [B,1,H,W] (+ top-2 disparities / costs)."""
    _chk(cost, samples, off)
    B, D, H, W = cost.shape
    disp = torch.empty((B, 1, H, W), device=cost.device, dtype=torch.float32)
    td = tc = None
    if want_top:
        td = torch.empty((B, 2, H, W), device=cost.device, dtype=torch.float32)
        tc = torch.empty_like(td)
    _lib.call("tstereo_predict_disp", _p(cost), _p(samples), _p(off), _p(disp), _p(td), _p(tc), B, D, H, W, _stream())
    return disp, td, tc


def range_samples(disp: torch.Tensor, radius: float, samples: torch.Tensor, c_off: int = 0):
    """low/high = disp -/+ radius and the 5 range candidates written at channel c_off of `samples`
    (aggregation/TemporalStereo/TemporalStereo.py:103-110, fine.py:78-86)."""
    _chk(disp, samples)
    B, one, H, W = disp.shape
    if one != 1 or samples.shape[0] != B or tuple(samples.shape[-2:]) != (H, W) or c_off + 5 > samples.shape[1]:
        raise ValueError(f"range_samples: disp {tuple(disp.shape)} does not match samples {tuple(samples.shape)} at channel {c_off}")
    low = torch.empty_like(disp)
    high = torch.empty_like(disp)
    _lib.call("tstereo_range_samples", _p(disp), float(radius), _p(low), _p(high), _p(samples), samples.shape[1],
              c_off, B, H, W, _stream())
    return low, high


def convex_upsample(m: torch.Tensor, w: torch.Tensor, b: torch.Tensor, disp: torch.Tensor) -> torch.Tensor:
    """ConvexUpsample tail (module.py:318-353) given the 64-channel mask features."""
    _chk(m, w, b, disp)
    B, _, H, W = disp.shape
    assert m.shape == (B, 64, H, W)
    out = torch.empty((B, 1, 2 * H, 2 * W), device=disp.device, dtype=torch.float32)
    _lib.call("tstereo_convex_upsample", _p(m), _p(w), _p(b), _p(disp), _p(out), B, H, W, _stream())
    return out


def unet_upsample(logits: torch.Tensor, disp: torch.Tensor) -> torch.Tensor:
    """UNet.upsample (module.py:468-483)."""
    _chk(logits, disp)
    B, nine, H, W = logits.shape
    assert nine == 9
    h, w = disp.shape[-2:]
    out = torch.empty((B, 1, H, W), device=disp.device, dtype=torch.float32)
    _lib.call("tstereo_unet_upsample", _p(logits), _p(disp), _p(out), B, H, W, h, w, _stream())
    return out


def bilinear_resize(x: torch.Tensor, size, mul: float = 1.0, div: float = 1.0, out: Optional[torch.Tensor] = None,
                    c_off: int = 0) -> torch.Tensor:
    """F.interpolate(x * mul / div, size, 'bilinear', align_corners=True), optionally into a channel slice."""
    _chk(x, out)
    B, Cc, Hi, Wi = x.shape
    Ho, Wo = size
    if out is None:
        out = torch.empty((B, Cc, Ho, Wo), device=x.device, dtype=torch.float32)
    elif out.shape[0] != B or tuple(out.shape[-2:]) != (Ho, Wo) or c_off + Cc > out.shape[1]:
        raise ValueError(f"bilinear_resize: out {tuple(out.shape)} cannot hold {Cc} channels of {(Ho, Wo)} at channel {c_off}")
    _lib.call("tstereo_bilinear_resize", _p(x), _p(out), float(mul), float(div), B, Cc, Hi, Wi, Ho, Wo,
              out.shape[1], c_off, _stream())
    return out


# --------------------------------------------------------------------------- temporal warp
def pose_prep(K: torch.Tensor, T_now: torch.Tensor, inv_T_prev: torch.Tensor, baseline: torch.Tensor,
              factor: float) -> torch.Tensor:
    _chk(K, T_now, inv_T_prev, baseline)
    B = K.shape[0]
    assert K.shape == (B, 4, 4) and T_now.shape == (B, 4, 4) and inv_T_prev.shape == (B, 4, 4)
    assert baseline.numel() == B
    params = torch.empty((B, 24), device=K.device, dtype=torch.float32)
    _lib.call("tstereo_pose_prep", _p(K), _p(T_now), _p(inv_T_prev), _p(baseline), float(factor), _p(params), B,
              _stream())
    return params


def reproject_disp(disp: torch.Tensor, params: torch.Tensor, want_flow: bool = True, want_disp: bool = True,
                   out: Optional[torch.Tensor] = None, c_off: int = 0):
    _chk(disp, params, out)
    B, Cc, h, w = disp.shape
    flow = torch.empty((B, 2, h, w), device=disp.device, dtype=torch.float32) if want_flow else None
    if want_disp and out is None:
        out = torch.empty((B, Cc, h, w), device=disp.device, dtype=torch.float32)
    ct = out.shape[1] if out is not None else Cc
    if out is not None and (out.shape[0] != B or tuple(out.shape[-2:]) != (h, w) or c_off + Cc > ct):
        raise ValueError(f"reproject_disp: out {tuple(out.shape)} cannot hold {Cc} channels of {(h, w)} at channel {c_off}")
    _lib.call("tstereo_reproject_disp", _p(disp), _p(params), _p(flow), _p(out if want_disp else None), ct, c_off,
              B, Cc, h, w, _stream())
    return flow, out


def project_depth(depth: torch.Tensor, params: torch.Tensor):
    """Kernel behind the `project_to_3d` drop-in: (optical_flow [B,2C,h,w], triangular_depth [B,C,h,w])."""
    _chk(depth, params)
    B, Cc, h, w = depth.shape
    flow = torch.empty((B, 2 * Cc, h, w), device=depth.device, dtype=torch.float32)
    tri = torch.empty_like(depth)
    _lib.call("tstereo_project_to_3d", _p(depth), _p(params), _p(flow), _p(tri), B, Cc, h, w, _stream())
    return flow, tri


def splat_metric(pd: torch.Tensor) -> torch.Tensor:
    _chk(pd)
    B, Cc, h, w = pd.shape
    metric = torch.empty((B, 1, h, w), device=pd.device, dtype=torch.float32)
    scratch = torch.empty((1024,), device=pd.device, dtype=torch.float32)
    _lib.call("tstereo_splat_metric", _p(pd), _p(metric), _p(scratch), B, Cc, h, w, _stream())
    return metric


def update_map_fused(prev_disp: torch.Tensor, K: torch.Tensor, T_now: torch.Tensor, inv_T_prev: torch.Tensor,
                     baseline: torch.Tensor, mem_sample: Optional[torch.Tensor], mem_cost: Optional[torch.Tensor],
                     local_map: Optional[torch.Tensor], n_lm_out: int, hw: Tuple[int, int]):
    """The whole temporal warp in three launches (TemporalStereo.py:326-461): returns (warped samples, warped costs,
    warped local map), each None when its input group is absent.  `hw` = the 1/8-scale size of the state."""
    _chk(prev_disp, K, T_now, inv_T_prev, baseline, mem_sample, mem_cost, local_map)
    B, _, Hf, Wf = prev_disp.shape
    h, w = hw
    M = mem_sample.shape[1] if mem_sample is not None else 0
    n_in = local_map.shape[1] if local_map is not None else 0
    if mem_sample is not None:
        assert mem_sample.shape == (B, M, h, w) and mem_cost.shape == (B, M, h, w)
    if local_map is not None:
        assert local_map.shape == (B, n_in, h, w)
    dev = prev_disp.device
    out_s = torch.empty((B, M, h, w), device=dev, dtype=torch.float32) if M else None
    out_c = torch.empty((B, M, h, w), device=dev, dtype=torch.float32) if M else None
    out_l = torch.empty((B, n_lm_out, h, w), device=dev, dtype=torch.float32) if n_lm_out else None
    n = _lib.load().tstereo_update_map_scratch_floats(B, h, w, M, n_lm_out)
    scratch = torch.empty((int(n),), device=dev, dtype=torch.float32)
    _lib.call("tstereo_update_map", _p(prev_disp), Hf, Wf, _p(K), _p(T_now), _p(inv_T_prev), _p(baseline), _p(mem_sample),
              _p(mem_cost), M, _p(local_map), n_in, n_lm_out, _p(out_s), _p(out_c), _p(out_l), _p(scratch), B, h, w, _stream())
    return out_s, out_c, out_l


def softsplat(x: torch.Tensor, flow: torch.Tensor, metric: torch.Tensor) -> torch.Tensor:
    _chk(x, flow, metric)
    B, Cc, h, w = x.shape
    assert flow.shape == (B, 2, h, w) and metric.shape == (B, 1, h, w)
    acc = torch.empty((B, Cc + 1, h, w), device=x.device, dtype=torch.float32)
    out = torch.empty_like(x)
    _lib.call("tstereo_softsplat", _p(x), _p(flow), _p(metric), _p(acc), _p(out), B, Cc, h, w, _stream())
    return out


# --------------------------------------------------------------------------- formats either side of the path
IMAGENET_MEAN, IMAGENET_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)      # data/datasets/base.py:60-61


def normalize_u8(img: torch.Tensor, mean=IMAGENET_MEAN, std=IMAGENET_STD, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """uint8 [B,H,W,3] (the decoded image as it leaves PIL) -> ImageNet-normalised fp32 [B,3,H,W]: ToTensor + Normalize of
    the reference's data pipeline (data/datasets/base.py:120-127) on the device, so images cross PCIe as bytes."""
    import ctypes
    if not img.is_cuda or img.dtype != torch.uint8 or img.dim() != 4 or img.shape[-1] != 3 or not img.is_contiguous():
        raise TypeError("normalize_u8 needs a contiguous CUDA uint8 tensor [B,H,W,3]")
    _same_device(img)
    B, H, W, _ = img.shape
    out = _out(out, (B, 3, H, W), img)
    osB, osC, _ = _view5(out)
    m, s = (ctypes.c_float * 3)(*mean), (ctypes.c_float * 3)(*std)
    _lib.call("tstereo_normalize_u8", img.data_ptr(), _p(out), osB, osC, B, H, W, ctypes.addressof(m), ctypes.addressof(s), _stream())
    return out


def disp_error(est: torch.Tensor, gt: torch.Tensor, lb: Optional[float] = None, ub: Optional[float] = None) -> torch.Tensor:
    """calc_error (data/evaluation/pixel_error.py:6-71) on the device: returns a float64 tensor
    [sum |gt - est|, count, #>1px, #>2px, #>3px, #>5px] over lb < gt < ub; `error_dict` turns it into the reference's dict."""
    _chk(est, gt)
    if est.shape != gt.shape:
        raise ValueError("disp_error: shapes differ")
    acc = torch.empty((6,), device=est.device, dtype=torch.float64)
    _lib.call("tstereo_disp_error", _p(est), _p(gt), float(lb or 0.0), float(ub or 0.0), int(lb is not None), int(ub is not None),
              est.numel(), acc.data_ptr(), _stream())
    return acc


def loss_smooth_l1(est: torch.Tensor, gt: torch.Tensor, max_disp: float = 192, start_disp: float = 0, sparse: bool = False) -> torch.Tensor:
    """`DispSmoothL1Loss.loss_per_level` (losses/smooth_l1_loss.py:49-74) on the device, forward only: a 0-dim float32 tensor."""
    _chk(est, gt)
    B, _, H, W = est.shape
    Hg, Wg = gt.shape[-2:]
    acc = torch.empty((2,), device=est.device, dtype=torch.float64)
    _lib.call("tstereo_loss_smooth_l1", _p(est), _p(gt), B, H, W, Hg, Wg, float(max_disp), float(start_disp), int(bool(sparse)),
              acc.data_ptr(), _stream())
    return (acc[0] / acc[1].clamp_min(1.0)).float()


def loss_wasserstein(cost: torch.Tensor, off: torch.Tensor, samples: torch.Tensor, gt: torch.Tensor, max_disp: float = 192,
                     start_disp: float = 0, sparse: bool = False) -> torch.Tensor:
    """`WarssersteinDistanceLoss.loss_per_level` (losses/warsserstein_distance_loss.py:53-81) on the device, forward only."""
    _chk(cost, off, samples, gt)
    B, D, H, W = cost.shape
    if off.shape != cost.shape or samples.shape != cost.shape:
        raise ValueError(f"cost {tuple(cost.shape)}, offsets {tuple(off.shape)} and samples {tuple(samples.shape)} must agree")
    Hg, Wg = gt.shape[-2:]
    acc = torch.empty((2,), device=cost.device, dtype=torch.float64)
    _lib.call("tstereo_loss_wasserstein", _p(cost), _p(off), _p(samples), _p(gt), B, D, H, W, Hg, Wg, float(max_disp),
              float(start_disp), int(bool(sparse)), acc.data_ptr(), _stream())
    return (acc[0] / float(B * H * W)).float()


def error_dict(acc: torch.Tensor) -> dict:
    """The reference's result dict (percentages and EPE) from `disp_error`'s accumulator; one 48-byte read-back."""
    s, n, c1, c2, c3, c5 = [float(v) for v in acc.cpu()]
    if n < 1.0:
        return {"1px": 0.0, "2px": 0.0, "3px": 0.0, "5px": 0.0, "epe": 0.0}
    return {"1px": 100.0 * c1 / n, "2px": 100.0 * c2 / n, "3px": 100.0 * c3 / n, "5px": 100.0 * c5 / n, "epe": s / n}
