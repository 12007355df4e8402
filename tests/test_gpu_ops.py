"""Per-operator parity of the CUDA kernels (through the C ABI) against the CPU oracle.

Every test feeds the same seeded inputs to oracle/oracle.py (or the torch fp32 CPU op the oracle
itself calls) and to libtstereo.so, at sizes the oracle finishes in well under a second, including
odd / ragged sizes (floor pooling, tile tails, W not a multiple of 32).  Index work (sort order,
top-2 selection) must be exact; floating point is compared with the tolerance written per test.
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import oracle as O
from temporalstereo_b200 import synth

pytestmark = pytest.mark.gpu

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


@pytest.fixture(scope="module")
def ops():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from temporalstereo_b200 import ops as _ops
    return _ops


def rnd(*shape, seed=0, scale=1.0):
    rng = np.random.RandomState(seed)
    return torch.from_numpy((scale * rng.standard_normal(shape)).astype(np.float32))


def pack(w):
    """[Cout, Cin, taps...] -> [Cin][taps][CoutP] (the layout include/tstereo.h documents)."""
    cout, cin = w.shape[:2]
    w = w.reshape(cout, cin, -1)
    coutp = (cout + 3) // 4 * 4
    p = torch.zeros(cin, w.shape[2], coutp)
    p[:, :, :cout] = w.permute(1, 2, 0)
    return p.contiguous()


def close(got, want, atol, rtol=0.0, what=""):
    got = got.detach().cpu()
    assert got.shape == want.shape, (what, got.shape, want.shape)
    assert torch.isfinite(got).all(), f"{what}: non-finite output"
    torch.testing.assert_close(got, want, atol=atol, rtol=rtol, msg=lambda m: f"{what}: {m}")


# --------------------------------------------------------------------------- cost volume (a1-a3)
def _load(golden_dir, name):
    return {k: torch.from_numpy(v) for k, v in np.load(os.path.join(golden_dir, name)).items()}


def _op_inputs(seed, B, C, H, W, S):
    rng = np.random.RandomState(seed)
    L = torch.from_numpy(rng.standard_normal((B, C, H, W)).astype(np.float32))
    R = torch.from_numpy(rng.standard_normal((B, C, H, W)).astype(np.float32))
    smp = torch.from_numpy(rng.uniform(-3.0, W / 2.0, (B, S, H, W)).astype(np.float32))
    return L, R, smp


@pytest.mark.parametrize("tag", ["a", "b"])
def test_block_cost_golden(ops, golden_dir, tag):
    """Against outputs of the real reference block_cost (tests/golden, odd sizes)."""
    g = _load(golden_dir, f"block_cost_warp_{tag}.npz")
    B, C, H, W, S = [int(v) for v in g["shape"]]
    L, R, smp = _op_inputs(10, B, C, H, W, S)
    out = ops.block_cost(L.cuda(), R.cuda(), smp.cuda())
    # the [L, warp(R)] half is a copy + one bilinear blend; group terms are sums of 8 squares
    close(out[:, :C], g["out"][:, :C], 0.0, what="L broadcast (bit-exact)")
    close(out, g["out"], 2e-5, what="block_cost warp")
    g = _load(golden_dir, f"block_cost_shift_{tag}.npz")
    out = ops.block_cost(L.cuda(), R.cuda(), S)
    close(out[:, :C], g["out"][:, :C], 0.0, what="-(L-R_d)^2 (bit-exact)")
    close(out, g["out"], 2e-5, what="block_cost shift")


@pytest.mark.parametrize("B,C,H,W,S", [(1, 32, 34, 60, 5), (2, 16, 17, 45, 7), (1, 128, 24, 40, 5), (1, 8, 4, 4, 2),
                                        (1, 16, 33, 70, 3)])
def test_block_cost_warp_vs_oracle(ops, B, C, H, W, S):
    L, R, smp = _op_inputs(3, B, C, H, W, S)
    smp[:, 0] = torch.round(smp[:, 0])          # integer candidates: exact taps
    smp[:, -1] = smp[:, -1] + W                 # fully out-of-range candidates -> zeros
    want = O.block_cost(L, R, smp, 3)
    out = ops.block_cost(L.cuda(), R.cuda(), smp.cuda())
    close(out[:, :C], want[:, :C], 0.0, what="L half")
    close(out[:, C:2 * C], want[:, C:2 * C], 1e-5, what="warped R half")
    close(out[:, 2 * C:], want[:, 2 * C:], 5e-5, rtol=1e-5, what="group terms")


@pytest.mark.parametrize("B,C,H,W,D", [(1, 256, 20, 36, 12), (2, 16, 9, 13, 20), (1, 8, 6, 5, 8), (1, 64, 34, 60, 16)])
def test_block_cost_shift_vs_oracle(ops, B, C, H, W, D):
    L, R, _ = _op_inputs(4, B, C, H, W, 1)
    want = O.block_cost(L, R, D, 3)
    out = ops.block_cost(L.cuda(), R.cuda(), D)
    close(out[:, :C], want[:, :C], 0.0, what="difference half")
    close(out[:, C:], want[:, C:], 5e-5, rtol=1e-5, what="group terms")


@pytest.mark.parametrize("B,C,H,W,S", [(1, 32, 34, 60, 5), (2, 16, 17, 45, 7), (1, 8, 4, 4, 2)])
def test_group_cost_equals_the_volume_tail(ops, B, C, H, W, S):
    """The group-only kernel is the materialising kernel minus its stores: bit-equal to the last 3C/8 channels."""
    L, R, smp = _op_inputs(5, B, C, H, W, S)
    full = ops.block_cost(L.cuda(), R.cuda(), smp.cuda())
    g = ops.group_cost(L.cuda(), R.cuda(), smp.cuda())
    assert torch.equal(g, full[:, 2 * C:])
    full = ops.block_cost(L.cuda(), R.cuda(), S)
    g = ops.group_cost(L.cuda(), R.cuda(), S)
    assert torch.equal(g, full[:, C:])


# fused cost volume -> first (1,3,3) conv: against the fp64 conv of the ORACLE's volume; tolerance = the fp32 conv's
COST_CONV_CASES = [
    # B, C, Cout, S, H, W
    (1, 128, 8, 5, 24, 40),      # precise-level channel counts
    (2, 128, 16, 8, 17, 45),     # fine level with local-map candidates, ragged tiles
    (1, 16, 8, 3, 33, 70),       # tile tails in x (W > 2 tiles) and y
    (1, 64, 32, 2, 9, 13),       # Cout = 32 (CP = 32 instance)
    (1, 8, 20, 4, 12, 31),       # padded output channels, group terms = 3 channels (one partial chunk)
]


@pytest.mark.parametrize("half", [False, True])
@pytest.mark.parametrize("B,C,Cout,S,H,W", COST_CONV_CASES)
def test_cost_conv_warp(ops, B, C, Cout, S, H, W, half):
    L, R, smp = _op_inputs(6, B, C, H, W, S)
    smp[:, 0] = torch.round(smp[:, 0])
    smp[:, -1] = smp[:, -1] + W                 # a fully out-of-range candidate
    planes = 2 * C + 3 * (C // 8)
    w = rnd(Cout, planes, 1, 3, 3, seed=61, scale=(2.0 / (9 * planes)) ** 0.5)
    b = rnd(Cout, seed=62, scale=0.1)
    vol = O.block_cost(L, R, smp, 3)
    want = O._act(F.conv3d(vol.double(), w.double(), b.double(), 1, (0, 1, 1)), "SiLU").float()
    w9 = w.reshape(Cout, planes, 9)
    g = ops.group_cost(L.cuda(), R.cuda(), smp.cuda())
    addl = ops.conv_hw3_tc2(L.cuda(), ops.pack_conv_hw3_tc2(w9[:, :C].contiguous(), half).cuda(), None, Cout, 1, None, half=half)
    got = ops.cost_conv_warp(R.cuda(), smp.cuda(), g, addl, ops.pack_conv_hw3_tc2(w9[:, C:].contiguous(), half).cuda(), b.cuda(),
                             Cout, "SiLU", half=half)
    close(got, want, 2e-5, rtol=1e-5, what="cost_conv_warp")
    # and against the two-step path it replaces (materialised volume -> tensor-core conv): same math, other summation order
    two = ops.conv_hw3_tc2(ops.block_cost(L.cuda(), R.cuda(), smp.cuda()), ops.pack_conv_hw3_tc2(w9, half).cuda(), b.cuda(), Cout, 1,
                           "SiLU", half=half)
    close(got, two.cpu(), 2e-5, rtol=1e-5, what="fused vs materialised")


@pytest.mark.parametrize("B,C,Cout,S,H,W", COST_CONV_CASES[:4] + [(1, 128, 8, 5, 136, 240)])
def test_cost_taps(ops, B, C, Cout, S, H, W):
    """Tap-projection form of the warp levels' first conv (1x1 projection of R once per frame + per-candidate gather / lerp
    + 3x3 conv over the group-wise channels) against the fp64 conv of the ORACLE's volume, and against the producer form."""
    L, R, smp = _op_inputs(6, B, C, H, W, S)
    smp[:, 0] = torch.round(smp[:, 0])
    smp[:, -1] = smp[:, -1] + W                 # a fully out-of-range candidate
    planes = 2 * C + 3 * (C // 8)
    w = rnd(Cout, planes, 1, 3, 3, seed=61, scale=(2.0 / (9 * planes)) ** 0.5)
    b = rnd(Cout, seed=62, scale=0.1)
    vol = O.block_cost(L, R, smp, 3)
    want = O._act(F.conv3d(vol.double(), w.double(), b.double(), 1, (0, 1, 1)), "SiLU").float()
    w9 = w.reshape(Cout, planes, 9)
    Lc, Rc, sc = L.cuda(), R.cuda(), smp.cuda()
    g = ops.group_cost(Lc, Rc, sc)
    addl = ops.conv_hw3_tc2(Lc, ops.pack_conv_hw3_tc2(w9[:, :C].contiguous(), True).cuda(), None, Cout, 1, None, half=True)
    wt, osc_t = ops.fp16_prescale(ops.tap_projection_weights(w9[:, C:2 * C].contiguous()))
    T = ops.conv_d_tc2(Rc.unsqueeze(2), ops.pack_conv_d_tc2(wt, True).cuda(), None, 9 * Cout, 1, 1, 1, False, None, half=True,
                       oscale=osc_t.cuda()).view(B, 9 * Cout, H, W)
    gc = ops.conv_hw3_tc2(g, ops.pack_conv_hw3_tc2(w9[:, 2 * C:].contiguous(), True).cuda(), None, Cout, 1, None, half=True)
    so = ops.Split(B, Cout, S, H, W, 2, device="cuda")
    got, _ = ops.cost_taps(T, sc, gc, addl, b.cuda(), Cout, "SiLU", sout=so, want_f32=True)
    # the 136x240 case: with these white-noise features the group-wise channels reach |g| ~ 30 and the materialising
    # kernel's fp32 evaluation of them differs from the oracle's by up to 8e-4 (3e-5 relative; every GPU form — taps,
    # producer, materialised — then sits 1.2e-4 from the oracle and within 1e-5 of each other, scripts/diag_taps.py)
    close(got, want, 2e-5 if H * W < 10000 else 2e-4, rtol=1e-5, what="cost_taps")
    assert torch.equal(so.t.view(torch.int16), ops.split_pack(got).t.view(torch.int16))
    prod = ops.cost_conv_warp(Rc, sc, g, addl, ops.pack_conv_hw3_tc2(w9[:, C:].contiguous(), True).cuda(), b.cuda(), Cout, "SiLU", half=True)
    close(got, prod.cpu(), 2e-5, rtol=1e-5, what="taps vs producer form")
    # into a strided view; without the optional addends
    buf = torch.full((B, Cout + 8, S, H, W), 7.0, device="cuda")
    ops.cost_taps(T, sc, None, None, b.cuda(), Cout, None, out=buf[:, 8:])
    ref, _ = ops.cost_taps(T, sc, None, None, b.cuda(), Cout, None)
    assert torch.equal(buf[:, 8:], ref) and (buf[:, :8] == 7).all()


@pytest.mark.parametrize("half", [False, True])
@pytest.mark.parametrize("B,C,Cout,D,H,W", [(1, 256, 32, 12, 20, 36), (2, 16, 8, 20, 9, 45), (1, 64, 16, 16, 34, 60)])
def test_cost_conv_shift(ops, B, C, Cout, D, H, W, half):
    L, R, _ = _op_inputs(7, B, C, H, W, 1)
    planes = C + 3 * (C // 8)
    w = rnd(Cout, planes, 1, 3, 3, seed=63, scale=(2.0 / (9 * planes)) ** 0.5)
    b = rnd(Cout, seed=64, scale=0.1)
    vol = O.block_cost(L, R, D, 3)
    want = O._act(F.conv3d(vol.double(), w.double(), b.double(), 1, (0, 1, 1)), "SiLU").float()
    g = ops.group_cost(L.cuda(), R.cuda(), D)
    got = ops.cost_conv_shift(L.cuda(), R.cuda(), g, ops.pack_conv_hw3_tc2(w.reshape(Cout, planes, 9), half).cuda(), b.cuda(), Cout,
                              "SiLU", half=half)
    close(got, want, 2e-5, rtol=1e-5, what="cost_conv_shift")


def test_cost_conv_into_a_strided_view_and_bad_arguments(ops):
    from temporalstereo_b200._lib import TStereoError
    B, C, Cout, S, H, W = 1, 16, 8, 3, 12, 20
    L, R, smp = _op_inputs(8, B, C, H, W, S)
    planes = 2 * C + 3 * (C // 8)
    w = rnd(Cout, planes, 9, seed=65, scale=0.05)
    g = ops.group_cost(L.cuda(), R.cuda(), smp.cuda())
    wp = ops.pack_conv_hw3_tc2(w[:, C:].contiguous(), True).cuda()
    buf = torch.full((B, 24, S, H, W), 7.0, device="cuda")
    ops.cost_conv_warp(R.cuda(), smp.cuda(), g, None, wp, None, Cout, None, out=buf[:, 8:16], half=True)
    ref = ops.cost_conv_warp(R.cuda(), smp.cuda(), g, None, wp, None, Cout, None, half=True)
    assert torch.equal(buf[:, 8:16], ref) and (buf[:, :8] == 7).all() and (buf[:, 16:] == 7).all()
    with pytest.raises(ValueError):
        ops.cost_conv_warp(R.cuda(), smp.cuda(), g, None, wp, None, Cout, None, out=buf[:, 8:17], half=True)
    with pytest.raises((TStereoError, AssertionError)):
        ops.cost_conv_warp(R.cuda()[:, :12].contiguous(), smp.cuda(), g, None, wp, None, Cout, None, half=True)


def test_block_cost_rejects_bad_arguments(ops):
    from temporalstereo_b200._lib import TStereoError
    L = torch.zeros(1, 12, 8, 8, device="cuda")
    with pytest.raises(TStereoError):
        ops.block_cost(L, L, 4)                  # C not a multiple of 8 (reference asserts, block_cost.py:9)
    with pytest.raises(TypeError):
        ops.block_cost(L.cpu(), L.cpu(), 4)      # no CPU fallback


# --------------------------------------------------------------------------- convolutions (a4-a7)
HW3_CASES = [
    # B, Cin, Cout, D, H, W, stride, dil, act, bias
    (1, 44, 32, 3, 17, 30, 1, 1, "SiLU", True),
    (1, 304, 8, 2, 20, 37, 1, 1, "SiLU", True),
    (2, 16, 16, 2, 9, 33, 1, 1, None, False),
    (1, 32, 64, 3, 17, 31, 2, 1, "SiLU", True),
    (1, 8, 16, 5, 34, 60, 2, 1, "SiLU", True),
    (1, 32, 32, 2, 17, 30, 1, 2, "SiLU", True),
    (1, 8, 8, 3, 12, 70, 1, 2, "SiLU", True),
    (1, 16, 16, 2, 11, 19, 1, 2, None, True),
    (1, 3, 32, 1, 32, 48, 2, 1, "ReLU", True),
    (1, 64, 64, 1, 9, 15, 1, 1, "ReLU", True),
    (1, 128, 36, 1, 10, 12, 1, 1, None, True),
    (1, 64, 9, 1, 10, 40, 1, 1, None, True),
    (1, 8, 8, 2, 70, 40, 1, 1, "SiLU", True),
    (1, 16, 32, 2, 8, 12, 2, 1, None, True),
]


@pytest.mark.parametrize("B,Cin,Cout,D,H,W,stride,dil,act,bias", HW3_CASES)
def test_conv_hw3(ops, B, Cin, Cout, D, H, W, stride, dil, act, bias):
    x = rnd(B, Cin, D, H, W, seed=1)
    w = rnd(Cout, Cin, 1, 3, 3, seed=2, scale=(2.0 / (9 * Cin)) ** 0.5)
    b = rnd(Cout, seed=3, scale=0.1) if bias else None
    want = O._act(F.conv3d(x, w, b, (1, stride, stride), (0, dil, dil), (1, dil, dil)), act)
    got = ops.conv_hw3(x.cuda(), pack(w).cuda(), None if b is None else b.cuda(), Cout, stride, dil, act)
    close(got, want, 2e-5, rtol=1e-5, what="conv_hw3")


def test_conv_hw3_2d_and_strided_views(ops):
    """4-D input (Conv2d) written into a channel slice of a wider tensor (the concat targets)."""
    x = rnd(2, 24, 13, 21, seed=5)
    w = rnd(16, 24, 3, 3, seed=6, scale=0.1)
    want = F.relu(F.conv2d(x, w, None, 1, 1))
    buf = torch.full((2, 40, 13, 21), 7.0, device="cuda")
    ops.conv_hw3(x.cuda(), pack(w).cuda(), None, 16, 1, 1, "ReLU", out=buf[:, 8:24])
    close(buf[:, 8:24], want, 2e-5, what="conv2d into slice")
    assert (buf[:, :8] == 7).all() and (buf[:, 24:] == 7).all(), "wrote outside the channel slice"
    # channel-sliced *input* view
    xin = torch.zeros(2, 40, 13, 21)
    xin[:, 8:32] = x
    got = ops.conv_hw3(xin.cuda()[:, 8:32], pack(w).cuda(), None, 16, 1, 1, "ReLU")
    close(got, want, 2e-5, what="conv2d from slice")


D_CASES = [
    # B, Cin, Cout, Din, HW(h,w), k, stride, dil, transposed, act
    (1, 32, 32, 12, (9, 14), 3, 1, 1, False, "SiLU"),
    (1, 64, 64, 12, (7, 9), 3, 2, 1, False, "SiLU"),
    (1, 16, 16, 5, (11, 13), 3, 2, 1, False, "SiLU"),
    (1, 16, 16, 3, (11, 13), 3, 2, 1, False, None),
    (2, 8, 8, 5, (20, 33), 3, 1, 2, False, "SiLU"),
    (1, 32, 32, 14, (9, 14), 5, 1, 1, False, "SiLU"),
    (1, 16, 32, 7, (9, 14), 3, 1, 1, False, "SiLU"),
    (1, 64, 64, 3, (5, 8), 3, 1, 1, True, None),
    (1, 16, 8, 2, (12, 19), 3, 1, 1, True, None),
    (1, 32, 32, 6, (600, 1), 3, 1, 1, True, None),
]


@pytest.mark.parametrize("B,Cin,Cout,Din,hw,k,stride,dil,transposed,act", D_CASES)
def test_conv_d(ops, B, Cin, Cout, Din, hw, k, stride, dil, transposed, act):
    H, W = hw
    x = rnd(B, Cin, Din, H, W, seed=7)
    b = rnd(Cout, seed=9, scale=0.1)
    if transposed:
        w = rnd(Cin, Cout, 3, 1, 1, seed=8, scale=0.1)
        want = O._act(F.conv_transpose3d(x, w, b, (2, 1, 1), (1, 0, 0), (1, 0, 0)), act)
        pw = pack(w.transpose(0, 1).contiguous())
    else:
        w = rnd(Cout, Cin, k, 1, 1, seed=8, scale=0.1)
        want = O._act(F.conv3d(x, w, b, (stride, 1, 1), (dil * (k // 2), 0, 0), (dil, 1, 1)), act)
        pw = pack(w)
    got = ops.conv_d(x.cuda(), pw.cuda(), b.cuda(), Cout, k, stride, dil, transposed, act)
    close(got, want, 2e-5, rtol=1e-5, what="conv_d")


@pytest.mark.parametrize("B,Cin,Cout,D,H,W,k,act", [
    (1, 64, 64, 3, 9, 15, 3, None), (1, 64, 32, 2, 17, 30, 3, None), (2, 16, 8, 2, 5, 7, 3, None),
    (1, 32, 32, 1, 17, 30, 4, "ReLU"), (1, 32, 9, 1, 20, 70, 4, None), (1, 16, 16, 2, 6, 65, 3, None)])
def test_deconv_hw(ops, B, Cin, Cout, D, H, W, k, act):
    x = rnd(B, Cin, D, H, W, seed=11)
    w = rnd(Cin, Cout, 1, k, k, seed=12, scale=0.1)
    b = rnd(Cout, seed=13, scale=0.1)
    op = (0, 1, 1) if k == 3 else (0, 0, 0)
    want = O._act(F.conv_transpose3d(x, w, b, (1, 2, 2), (0, 1, 1), op), act)
    got = ops.deconv_hw(x.cuda(), pack(w.transpose(0, 1).contiguous()).cuda(), b.cuda(), Cout, k, act)
    close(got, want, 2e-5, rtol=1e-5, what="deconv_hw")


@pytest.mark.parametrize("src,dst", [((6, 18, 30), (6, 17, 30)), ((4, 10, 16), (3, 9, 15)), ((6, 18, 30), (5, 17, 30)),
                                     ((12, 34, 60), (12, 34, 60)), ((2, 4, 6), (5, 9, 13))])
def test_resize_add_act(ops, src, dst):
    a = rnd(2, 5, *src, seed=14)
    skip = rnd(2, 5, *dst, seed=15)
    want = F.silu(F.interpolate(a, size=dst, mode="trilinear", align_corners=True) + skip)
    got = ops.resize_add_act(a.cuda(), dst, skip.cuda(), "SiLU")
    close(got, want, 2e-6, what="resize_add_act")
    got = ops.resize_add_act(a.cuda(), dst, None, None)
    close(got, F.interpolate(a, size=dst, mode="trilinear", align_corners=True), 2e-6, what="resize only")


@pytest.mark.parametrize("B,C,D,H,W", [(1, 4, 14, 34, 60), (2, 3, 7, 9, 33), (1, 2, 5, 5, 5), (1, 1, 6, 40, 5)])
def test_pool5(ops, B, C, D, H, W):
    x = rnd(B, C, D, H, W, seed=16)
    cat = torch.zeros(B, 3 * C, D, H, W, device="cuda")
    cat[:, :C] = x.cuda()
    ops.pool5(cat[:, :C], cat[:, C:2 * C], cat[:, 2 * C:])
    close(cat[:, C:2 * C], F.avg_pool3d(x, 5, 1, 2), 1e-6, what="avg_pool3d 5^3")
    close(cat[:, 2 * C:], F.max_pool3d(x, 5, 1, 2), 0.0, what="max_pool3d 5^3 (exact)")


# --------------------------------------------------------------------------- memory merge (a8)
def _past_conv_sd(C, seed):
    rng = np.random.RandomState(seed)
    f = lambda *s: torch.from_numpy(rng.standard_normal(s).astype(np.float32))
    return {"m.past_conv.weight": f(C, 1, 1, 1, 1), "m.past_conv.norm.weight": 1 + 0.1 * f(C),
            "m.past_conv.norm.bias": 0.1 * f(C), "m.past_conv.norm.running_mean": 0.1 * f(C),
            "m.past_conv.norm.running_var": 1 + 0.1 * f(C).abs()}


def _fold_past(sd):
    s = sd["m.past_conv.norm.weight"] / torch.sqrt(sd["m.past_conv.norm.running_var"] + 1e-5)
    w = sd["m.past_conv.weight"].reshape(-1) * s
    b = (0 - sd["m.past_conv.norm.running_mean"]) * s + sd["m.past_conv.norm.bias"]
    return w.contiguous(), b.contiguous()


@pytest.mark.parametrize("with_memory", [False, True])
@pytest.mark.parametrize("B,C,D,H,W", [(1, 32, 12, 9, 14), (2, 16, 8, 7, 33)])
def test_merge_memory(ops, with_memory, B, C, D, H, W):
    vol = rnd(B, C, D, H, W, seed=17)
    if D == 12:   # the coarse level's constant integer candidates: ties with the zero memory samples
        samples = torch.linspace(0, D - 1, D).view(1, D, 1, 1).expand(B, D, H, W).contiguous()
    else:
        samples = rnd(B, D, H, W, seed=18, scale=5.0)
        samples[:, 3] = samples[:, 1]            # exact ties inside the candidate list
    sd = _past_conv_sd(C, 19)
    prev = {}
    ms = mc = None
    if with_memory:
        ms = rnd(B, 2, H, W, seed=20, scale=4.0)
        ms[:, 1] = samples[:, 2]                 # tie between a memory sample and a candidate
        mc = rnd(B, 2, H, W, seed=21)
        prev = {"cost_memory": {"disp_sample": ms, "cost_volume": mc}, "use_past_cost": True}
    want_vol, want_s = O.merge_memory(vol, samples, sd, "m", prev, 2, coarse=False)
    pw, pb = _fold_past(sd)
    got_vol, got_s = ops.merge_memory(vol.cuda(), samples.cuda(), None if ms is None else ms.cuda(),
                                      None if mc is None else mc.cuda(), pw.cuda(), pb.cuda(), 2)
    close(got_s, want_s, 0.0, what="sorted samples (exact)")
    # planes coming from the volume are pure gathers -> exact; memory planes go through SiLU
    close(got_vol, want_vol, 2e-6, what="permuted volume")
    # the permutation itself: every plane taken from the input volume must match bit-for-bit
    src = torch.cat([samples, ms if ms is not None else torch.zeros(B, 2, H, W)], 1)
    order = torch.sort(src, dim=1, stable=True)[1]
    from_vol = (order < D).unsqueeze(1).expand_as(want_vol)
    assert torch.equal(got_vol.cpu()[from_vol], want_vol[from_vol]), "gathered planes differ: sort order mismatch"


# --------------------------------------------------------------------------- heads + regression (a10, a11, a14)
def test_heads(ops):
    B, C, D, H, W = 2, 16, 7, 9, 33
    feat = rnd(B, 2 * C, D, H, W, seed=22)
    wc = rnd(1, C, 1, 3, 3, seed=23, scale=3.0)
    wo = rnd(1, C, 1, 3, 3, seed=24, scale=30.0)     # large: exercises the tanh squashing
    want_cost = F.conv3d(feat[:, :C], wc, None, 1, (0, 1, 1)).squeeze(1)
    want_off = torch.tanh(F.conv3d(feat[:, C:], wo, None, 1, (0, 1, 1)).squeeze(1) / 100).clamp(-1, 1) * 1.0
    w = torch.stack([wc.reshape(C, 9), wo.reshape(C, 9)]).contiguous()
    cost, off = ops.heads(feat.cuda(), w.cuda(), 1.0)
    close(cost, want_cost, 5e-5, rtol=1e-5, what="cost head")
    close(off, want_off, 5e-6, what="offset head")


@pytest.mark.parametrize("B,D,H,W", [(1, 14, 9, 14), (2, 5, 7, 33), (1, 2, 3, 5), (1, 22, 4, 40)])
def test_predict_disp(ops, B, D, H, W):
    cost = rnd(B, D, H, W, seed=25, scale=2.0)
    samples = rnd(B, D, H, W, seed=26, scale=10.0)
    off = rnd(B, D, H, W, seed=27, scale=0.3)
    want, want_td, want_tc = O.predict_disp(cost, samples, off, 2)
    disp, td, tc = ops.predict_disp(cost.cuda(), samples.cuda(), off.cuda(), True)
    close(tc, want_tc, 0.0, what="top-2 costs (exact)")
    close(td, want_td, 0.0, what="top-2 disparities (exact: index work)")
    close(disp, want, 2e-6, what="regressed disparity")


def test_predict_disp_ties(ops):
    """Equal costs: ties resolve to the lowest index first, and the selected *values* equal torch.topk's."""
    cost = torch.tensor([1.0, 3.0, 3.0, 2.0, 3.0]).view(1, 5, 1, 1).repeat(1, 1, 2, 3).contiguous()
    samples = torch.arange(5.0).view(1, 5, 1, 1).repeat(1, 1, 2, 3).contiguous()
    off = torch.zeros_like(cost)
    disp, td, tc = ops.predict_disp(cost.cuda(), samples.cuda(), off.cuda(), True)
    assert torch.equal(tc.cpu(), torch.full((1, 2, 2, 3), 3.0))
    assert torch.equal(td.cpu()[:, 0], torch.full((1, 2, 3), 1.0)) and torch.equal(td.cpu()[:, 1], torch.full((1, 2, 3), 2.0))
    # all-equal costs (e.g. a constant volume): indices 0 and 1
    cost = torch.zeros(1, 4, 2, 2)
    disp, td, tc = ops.predict_disp(cost.cuda(), torch.arange(4.0).view(1, 4, 1, 1).expand(1, 4, 2, 2).contiguous().cuda(),
                                    cost.cuda(), True)
    assert torch.equal(td.cpu()[:, 0], torch.zeros(1, 2, 2)) and torch.equal(td.cpu()[:, 1], torch.ones(1, 2, 2))
    close(disp, torch.full((1, 1, 2, 2), 0.5), 0.0, what="tie regression")


def test_range_samples(ops):
    disp = rnd(2, 1, 9, 14, seed=28, scale=20.0)
    want = O.range_samples(disp - 4.0, disp + 4.0)
    buf = torch.full((2, 8, 9, 14), -1.0, device="cuda")
    low, high = ops.range_samples(disp.cuda(), 4.0, buf, 3)
    close(buf[:, 3:], want, 0.0, what="range candidates (exact)")
    close(low, disp - 4.0, 0.0, what="low")
    close(high, disp + 4.0, 0.0, what="high")
    assert (buf[:, :3] == -1).all()


# --------------------------------------------------------------------------- up-sampling (a12, a13, a16, a20)
def test_convex_upsample(ops):
    B, H, W = 2, 9, 33
    m = rnd(B, 64, H, W, seed=29)
    w = rnd(36, 64, 1, 1, seed=30, scale=0.3)
    b = rnd(36, seed=31, scale=0.1)
    disp = rnd(B, 1, H, W, seed=32, scale=10.0)
    logits = F.conv2d(m, w, b)
    mm = torch.softmax(logits.view(B, 1, 9, 2, 2, H, W), 2)
    u = F.unfold(disp * 2, (3, 3), padding=1).view(B, 1, 9, 1, 1, H, W)
    want = (mm * u).sum(2).permute(0, 1, 4, 2, 5, 3).reshape(B, 1, 2 * H, 2 * W)
    got = ops.convex_upsample(m.cuda(), w.reshape(36, 64).contiguous().cuda(), b.cuda(), disp.cuda())
    close(got, want, 2e-5, what="convex_upsample")


@pytest.mark.parametrize("h,w", [(9, 14), (24, 40), (5, 33)])
def test_unet_upsample(ops, h, w):
    B, H, W = 2, 4 * h, 4 * w
    logits = rnd(B, 9, H, W, seed=33, scale=2.0)
    disp = rnd(B, 1, h, w, seed=34, scale=10.0)
    want = O.unet_upsample(logits, disp)
    got = ops.unet_upsample(logits.cuda(), disp.cuda())
    close(got, want, 2e-5, what="unet_upsample")


@pytest.mark.parametrize("src,dst", [((24, 40), (12, 20)), ((12, 20), (6, 10)), ((6, 10), (96, 160)), ((7, 9), (7, 9))])
def test_bilinear_resize(ops, src, dst):
    x = rnd(2, 3, *src, seed=35, scale=5.0)
    want = F.interpolate(x * dst[1] / src[1], size=dst, mode="bilinear", align_corners=True)
    got = ops.bilinear_resize(x.cuda(), dst, mul=dst[1], div=src[1])
    close(got, want, 2e-6 * max(1.0, dst[1] / src[1]), rtol=1e-6, what="bilinear_resize")
    buf = torch.zeros(2, 5, *dst, device="cuda")
    ops.bilinear_resize(x.cuda(), dst, out=buf, c_off=2)
    close(buf[:, 2:], F.interpolate(x, size=dst, mode="bilinear", align_corners=True), 2e-6, what="into slice")
    assert (buf[:, :2] == 0).all()


# --------------------------------------------------------------------------- temporal warp (a17-a19)
def test_project_to_3d_golden(golden_dir):
    from temporalstereo_b200 import temporal
    g = _load(golden_dir, "project_to_3d.npz")
    st = synth.synthetic_temporal_state(64, 96, B=2)
    K8 = st["K"].clone(); K8[:, :2] /= 8.0
    T = torch.bmm(st["T_now"], st["inv_T_prev"])
    depth = 0.54 * K8[:, 0, 0].view(-1, 1, 1, 1) / (st["cost_memory"]["disp_sample"] + 1e-5)
    out = temporal.project_to_3d(depth.cuda(), K8.cuda(), None, T.cuda())
    close(out["optical_flow"], g["flow"], 2e-4, rtol=1e-5, what="optical_flow vs reference")
    close(out["triangular_depth"], g["tri"], 1e-5, rtol=1e-6, what="triangular_depth vs reference")


def test_softsplat_vs_oracle():
    from temporalstereo_b200 import temporal
    B, C, h, w = 2, 4, 12, 20
    x = rnd(B, C, h, w, seed=36)
    flow = rnd(B, 2, h, w, seed=37, scale=2.5)
    flow[0, :, 0, 0] = torch.tensor([-30.0, 4.0])      # lands outside: contributes nowhere
    flow[0, :, 1, 1] = torch.tensor([1.0, -1.0])       # exactly on a pixel centre
    metric = rnd(B, 1, h, w, seed=38, scale=3.0)
    want = O.softsplat_softmax(x, flow, metric)
    got = temporal.FunctionSoftsplat(x.cuda(), flow.cuda(), metric.cuda(), "softmax")
    # float atomics: summation order differs from the oracle's index_add_
    close(got, want, 2e-5, rtol=1e-5, what="softmax splat")
    holes = (want == 0).all(1)
    assert torch.equal((got.cpu() == 0).all(1), holes), "hole pattern (pixels nobody splats to) differs"


def _ref_splat_lib():
    import ctypes
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = os.path.join(root, "oracle", "_ref", "libsoftsplat_ref.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libsoftsplat_ref.so not built (python oracle/build_softsplat_ref.py, needs /root/reference)")
    lib = ctypes.CDLL(path)
    lib.softsplat_ref_launch.restype = ctypes.c_int
    lib.softsplat_ref_launch.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 4 + [ctypes.c_void_p]
    return lib


def _splat_case(i, B, C, h, w):
    x = rnd(B, C, h, w, seed=70 + i)
    flow = rnd(B, 2, h, w, seed=80 + i, scale=2.5)
    flow[0, :, 0, 0] = torch.tensor([-30.0, 4.0])      # lands outside: contributes nowhere
    flow[0, :, 1, 1] = torch.tensor([1.0, -1.0])       # exactly on a pixel centre
    flow[0, :, 2, 2] = torch.tensor([w - 3.5, 0.25])   # straddles the right border: two corners dropped
    metric = rnd(B, 1, h, w, seed=90 + i, scale=3.0)
    return x, flow, metric


def test_softsplat_vs_reference_kernel():
    """Row a19 pinned to the reference's OWN kernel: oracle/_ref/libsoftsplat_ref.so is kernel_Softsplat_updateOutput
    (softsplat.py:8-53) expanded by the reference's cupy_kernel (softsplat.py:179-232) and compiled with nvcc.  The packing
    and normalisation around it follow FunctionSoftsplat 'softmax' (softsplat.py:344-356).  Both sides add with float
    atomics, so the comparison carries a summation-order tolerance; the hole pattern must be identical."""
    from oracle.softsplat_shapes import SHAPES
    from temporalstereo_b200 import temporal
    lib = _ref_splat_lib()
    for i, (B, C, h, w) in enumerate(SHAPES):
        x, flow, metric = _splat_case(i, B, C, h, w)
        xg, fg, mg = x.cuda(), flow.cuda(), metric.cuda()
        packed = torch.cat([xg * mg.exp(), mg.exp()], 1).contiguous()
        acc = torch.zeros_like(packed)
        rc = lib.softsplat_ref_launch(packed.data_ptr(), fg.data_ptr(), acc.data_ptr(), B, C + 1, h, w,
                                      torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert rc == 0, f"reference kernel launch failed for shape {(B, C, h, w)}: {rc}"
        want = (acc[:, :-1] / (acc[:, -1:] + 1e-22)).cpu()
        got = temporal.FunctionSoftsplat(xg, fg, mg, "softmax")
        close(got, want, 2e-5, rtol=1e-5, what=f"splat vs reference kernel {(B, C, h, w)}")
        assert torch.equal((got.cpu() == 0).all(1), (want == 0).all(1)), "hole pattern differs from the reference kernel"
        # and the CPU restatement the rest of the temporal tests lean on is pinned by the same kernel
        close(O.softsplat_softmax(x, flow, metric), want, 2e-5, rtol=1e-5, what="oracle splat vs reference kernel")


@pytest.mark.parametrize("first_frame", [False, True])
def test_update_map_vs_oracle(golden_dir, first_frame):
    from temporalstereo_b200 import temporal
    H, W = 96, 160
    st = synth.synthetic_temporal_state(H, W, B=2)
    prev = dict(prev_disp=st["prev_disp"], cost_memory=st["cost_memory"], local_map=st["local_map"])
    if first_frame:
        prev.pop("local_map")
    want = O.update_map(dict(prev), st["K"], st["T_now"], st["inv_T_prev"], st["baseline"], H, W, True, 3)
    dev = {k: (v.cuda() if torch.is_tensor(v) else {a: b.cuda() for a, b in v.items()}) for k, v in prev.items()}
    got = temporal.update_map(dev, st["K"].cuda(), st["T_now"].cuda(), st["inv_T_prev"].cuda(), st["baseline"].cuda(),
                              H, W, True, 3)
    close(got["cost_memory"]["disp_sample"], want["cost_memory"]["disp_sample"], 2e-4, rtol=1e-5, what="warped samples")
    close(got["cost_memory"]["cost_volume"], want["cost_memory"]["cost_volume"], 2e-4, rtol=1e-5, what="warped costs")
    close(got["local_map"], want["local_map"], 2e-4, rtol=1e-5, what="local map")
    assert got["local_map_size"] == 3 and got["use_past_cost"] is True


def test_update_map_golden(golden_dir):
    """Against the real reference's update_map (with the CPU splat restatement), B=1."""
    from temporalstereo_b200 import temporal
    gm = _load(golden_dir, "update_map_96x160.npz")
    H, W = 96, 160
    st = synth.synthetic_temporal_state(H, W, B=1)
    dev = dict(prev_disp=st["prev_disp"].cuda(), local_map=st["local_map"].cuda(),
               cost_memory={k: v.cuda() for k, v in st["cost_memory"].items()})
    got = temporal.update_map(dev, st["K"].cuda(), st["T_now"].cuda(), st["inv_T_prev"].cuda(), st["baseline"].cuda(),
                              H, W, True, 3)
    close(got["cost_memory"]["disp_sample"], gm["mem_sample"], 2e-4, rtol=1e-5, what="warped samples")
    close(got["cost_memory"]["cost_volume"], gm["mem_cost"], 2e-4, rtol=1e-5, what="warped costs")
    close(got["local_map"], gm["local_map"], 2e-4, rtol=1e-5, what="local map")


# --------------------------------------------------------------------------- tensor-core convs (tcgen05, hi+lo split operands)
TC2_CASES = [
    # B, Cin, Cout, D, H, W, dil, act, bias
    (1, 8, 16, 1, 8, 16, 1, None, False),
    (1, 16, 8, 2, 9, 20, 1, "SiLU", True),
    (1, 44, 32, 3, 17, 30, 1, "SiLU", True),
    (2, 304, 8, 2, 20, 37, 1, "SiLU", True),
    (1, 352, 32, 2, 34, 60, 1, "SiLU", True),
    (1, 304, 16, 2, 19, 61, 1, "SiLU", True),
    (1, 128, 32, 2, 34, 60, 1, None, True),
    (1, 32, 32, 2, 17, 30, 2, "SiLU", True),
    (1, 13, 20, 1, 10, 12, 1, None, True),
    (1, 64, 32, 1, 40, 240, 1, "ReLU", True),
    (1, 8, 8, 5, 136, 240, 2, "SiLU", True),
    (1, 16, 16, 3, 5, 3, 2, "SiLU", True),
    (2, 32, 3, 1, 1, 1, 1, None, True),
]


@pytest.mark.parametrize("half", [False, True])
@pytest.mark.parametrize("B,Cin,Cout,D,H,W,dil,act,bias", TC2_CASES)
def test_conv_hw3_tc2(ops, B, Cin, Cout, D, H, W, dil, act, bias, half):
    """kx-folded tcgen05 conv (2-D tiles; tf32 hi+lo or fp16 hi+lo operands) == fp64 conv to fp32 rounding."""
    x = rnd(B, Cin, D, H, W, seed=41)
    w = rnd(Cout, Cin, 1, 3, 3, seed=42, scale=(2.0 / (9 * Cin)) ** 0.5)
    b = rnd(Cout, seed=43, scale=0.1) if bias else None
    want = O._act(F.conv3d(x.double(), w.double(), None if b is None else b.double(), 1, (0, dil, dil), (1, dil, dil)), act).float()
    wp = ops.pack_conv_hw3_tc2(w.reshape(Cout, Cin, 9).cuda(), half)
    got = ops.conv_hw3_tc2(x.cuda(), wp, None if b is None else b.cuda(), Cout, dil, act, half=half)
    close(got, want, 1e-5, rtol=1e-5, what="conv_hw3_tc2")


@pytest.mark.parametrize("B,Cin,Cout,D,H,W,dil,act", [
    (1, 16, 8, 2, 9, 20, 1, "SiLU"), (2, 304, 8, 2, 20, 36, 1, "SiLU"), (1, 44, 32, 3, 17, 32, 1, "SiLU"),
    (1, 32, 32, 2, 17, 28, 2, "SiLU"), (1, 64, 32, 1, 40, 240, 1, "ReLU"), (1, 8, 8, 5, 136, 240, 2, "SiLU")])
def test_conv_hw3_tc2_tma(ops, monkeypatch, B, Cin, Cout, D, H, W, dil, act):
    """The TMA producer variant (cp.async.bulk.tensor.5d boxes, zero-filled halos) gives the same results."""
    monkeypatch.setenv("TSTEREO_TC2_TMA", "1")
    x = rnd(B, Cin, D, H, W, seed=81)
    w = rnd(Cout, Cin, 1, 3, 3, seed=82, scale=(2.0 / (9 * Cin)) ** 0.5)
    b = rnd(Cout, seed=83, scale=0.1)
    want = O._act(F.conv3d(x.double(), w.double(), b.double(), 1, (0, dil, dil), (1, dil, dil)), act).float()
    got = ops.conv_hw3_tc2(x.cuda(), ops.pack_conv_hw3_tc2(w.reshape(Cout, Cin, 9).cuda()), b.cuda(), Cout, dil, act)
    close(got, want, 1e-5, rtol=1e-5, what="conv_hw3_tc2 (TMA)")
    # (k,1,1) conv along D: planes outside [0, Din) are zero-filled by the TMA unit
    xd = rnd(1, 16, 7, 9, 16, seed=84)
    wd = rnd(16, 16, 5, 1, 1, seed=85, scale=0.1)
    wantd = F.conv3d(xd, wd, None, 1, (2, 0, 0))
    gotd = ops.conv_d_tc2(xd.cuda(), ops.pack_conv_d_tc2(wd.reshape(16, 16, 5).cuda()), None, 16, 5, 1, 1, False, None)
    close(gotd, wantd, 1e-5, rtol=1e-5, what="conv_d_tc2 (TMA)")


def test_fp16_split_small_weights_need_the_prescale(ops):
    """BN-folded weights of magnitude 1e-3: the lo half of the fp16 split is subnormal (14 of 22 bits survive) unless the
    weights are pre-scaled per output channel by a power of two and the accumulator is scaled back (oscale)."""
    B, Cin, Cout, D, H, W = 1, 64, 16, 2, 20, 37
    x = rnd(B, Cin, D, H, W, seed=71)
    w = rnd(Cout, Cin, 1, 3, 3, seed=72, scale=1e-3)
    w[3] *= 50.0                                     # channels of very different magnitude get their own scale
    b = rnd(Cout, seed=73, scale=1e-3)
    want = F.conv3d(x.double(), w.double(), b.double(), 1, (0, 1, 1)).float()
    w9 = w.reshape(Cout, Cin, 9)
    plain = ops.conv_hw3_tc2(x.cuda(), ops.pack_conv_hw3_tc2(w9, True).cuda(), b.cuda(), Cout, 1, None, half=True).cpu()
    ws, inv = ops.fp16_prescale(w9)
    assert torch.equal(torch.log2(inv), torch.log2(inv).round()), "the scale must be a power of two (exact to undo)"
    m = ws.abs().flatten(1).amax(1)
    assert (m > 511).all() and (m <= 1023).all()
    scaled = ops.conv_hw3_tc2(x.cuda(), ops.pack_conv_hw3_tc2(ws, True).cuda(), b.cuda(), Cout, 1, None, half=True, oscale=inv.cuda()).cpu()
    ref = want.abs().flatten(2).amax(2).view(1, Cout, 1, 1, 1)     # per-channel output magnitude
    e_plain, e_scaled = ((plain - want).abs() / ref).max().item(), ((scaled - want).abs() / ref).max().item()
    print(f"fp16 split, |w| ~ 1e-3: relative error plain {e_plain:.2e}, pre-scaled {e_scaled:.2e}")
    assert e_scaled < 3e-6, e_scaled                 # the fp32 conv's own rounding level
    assert e_plain > 3 * e_scaled, (e_plain, e_scaled)


def test_conv_hw3_tc2_fp16_saturates_instead_of_nan(ops):
    """fp16 split: an activation beyond 65504 clamps (cvt.satfinite) — finite output; the tf32 split has no limit."""
    x = rnd(1, 8, 1, 8, 32, seed=96)
    x[0, 3, 0, 4, 7] = 3.0e5
    w = rnd(8, 8, 1, 3, 3, seed=97, scale=0.1)
    want = F.conv3d(x.double(), w.double(), None, 1, (0, 1, 1)).float()
    got_h = ops.conv_hw3_tc2(x.cuda(), ops.pack_conv_hw3_tc2(w.reshape(8, 8, 9).cuda(), True), None, 8, 1, None, half=True)
    assert torch.isfinite(got_h).all()
    got_t = ops.conv_hw3_tc2(x.cuda(), ops.pack_conv_hw3_tc2(w.reshape(8, 8, 9).cuda(), False), None, 8, 1, None, half=False)
    close(got_t, want, 1e-5, rtol=1e-5, what="tf32 split with a 3e5 activation")


def test_conv_tc2_cp_async_ring(ops, monkeypatch):
    """The cp.async raw-ring producer (opt-in) gives the same results: 3x3 with halo, odd width, (k,1,1) along D."""
    monkeypatch.setenv("TSTEREO_TC2_CPA", "1")
    x = rnd(2, 40, 3, 19, 45, seed=91)
    w = rnd(16, 40, 1, 3, 3, seed=92, scale=0.05)
    b = rnd(16, seed=93, scale=0.1)
    want = O._act(F.conv3d(x.double(), w.double(), b.double(), 1, (0, 1, 1)), "SiLU").float()
    got = ops.conv_hw3_tc2(x.cuda(), ops.pack_conv_hw3_tc2(w.reshape(16, 40, 9).cuda(), True), b.cuda(), 16, 1, "SiLU", half=True)
    close(got, want, 1e-5, rtol=1e-5, what="conv_hw3_tc2 (cp.async ring)")
    xd = rnd(1, 64, 6, 17, 30, seed=94)
    wd = rnd(64, 64, 3, 1, 1, seed=95, scale=0.05)
    wantd = F.conv3d(xd.double(), wd.double(), None, (2, 1, 1), (1, 0, 0)).float()
    gotd = ops.conv_d_tc2(xd.cuda(), ops.pack_conv_d_tc2(wd.reshape(64, 64, 3).cuda(), True), None, 64, 3, 2, 1, False, None, half=True)
    close(gotd, wantd, 1e-5, rtol=1e-5, what="conv_d_tc2 (cp.async ring)")


@pytest.mark.parametrize("mt", [2, 4])
def test_conv_hw3_tc2_tilings(ops, mt, monkeypatch):
    """Both M-tile counts (8- and 16-row tiles) on the same input; ragged right / bottom edges."""
    monkeypatch.setenv("TSTEREO_TC2_MT", str(mt))
    x = rnd(1, 24, 2, 37, 71, seed=61)
    w = rnd(8, 24, 1, 3, 3, seed=62, scale=0.1)
    want = F.conv3d(x.double(), w.double(), None, 1, (0, 1, 1)).float()
    got = ops.conv_hw3_tc2(x.cuda(), ops.pack_conv_hw3_tc2(w.reshape(8, 24, 9).cuda()), None, 8, 1, None)
    close(got, want, 1e-5, rtol=1e-5, what=f"conv_hw3_tc2 MT={mt}")


def test_conv_hw3_tc2_views(ops):
    """2-D input, channel-sliced input and output views (the concat buffers)."""
    x = rnd(2, 24, 13, 21, seed=45)
    w = rnd(16, 24, 3, 3, seed=46, scale=0.1)
    want = F.relu(F.conv2d(x, w, None, 1, 1))
    wp = ops.pack_conv_hw3_tc2(w.reshape(16, 24, 9).cuda())
    buf = torch.full((2, 40, 13, 21), 7.0, device="cuda")
    xin = torch.zeros(2, 40, 13, 21)
    xin[:, 8:32] = x
    ops.conv_hw3_tc2(xin.cuda()[:, 8:32], wp, None, 16, 1, "ReLU", out=buf[:, 8:24])
    close(buf[:, 8:24], want, 1e-5, rtol=1e-5, what="tc2 conv2d slices")
    assert (buf[:, :8] == 7).all() and (buf[:, 24:] == 7).all()


def test_conv_hw3_tc2_cout64(ops):
    """Cout > 32 runs as groups of 32 output channels."""
    x = rnd(1, 64, 1, 19, 45, seed=63)
    w = rnd(64, 64, 1, 3, 3, seed=64, scale=0.05)
    b = rnd(64, seed=65, scale=0.1)
    want = F.relu(F.conv3d(x.double(), w.double(), b.double(), 1, (0, 1, 1))).float()
    got = ops.conv_hw3_tc2(x.cuda(), ops.pack_conv_hw3_tc2(w.reshape(64, 64, 9).cuda()), b.cuda(), 64, 1, "ReLU")
    close(got, want, 1e-5, rtol=1e-5, what="conv_hw3_tc2 Cout=64")
    x = rnd(1, 256, 1, 9, 15, seed=66)
    w = rnd(64, 256, 1, 3, 3, seed=67, scale=0.03)
    want = O._act(F.conv3d(x.double(), w.double(), b.double(), 1, (0, 1, 1)), "SiLU").float()
    got = ops.conv_hw3_tc2(x.cuda(), ops.pack_conv_hw3_tc2(w.reshape(64, 256, 9).cuda()), b.cuda(), 64, 1, "SiLU")
    close(got, want, 1e-5, rtol=1e-5, what="conv_hw3_tc2 256->64")


S2_CASES = [
    # B, Cin, Cout, D, Hin, Win, act
    (1, 3, 32, 1, 32, 48, "ReLU"),
    (2, 8, 16, 3, 17, 30, "SiLU"),
    (1, 16, 16, 2, 9, 15, "SiLU"),
    (1, 32, 64, 1, 34, 61, "ReLU"),
    (1, 64, 64, 4, 17, 30, "SiLU"),
    (1, 32, 32, 5, 68, 120, None),
    (1, 12, 20, 1, 5, 7, None),
    (1, 8, 8, 1, 1, 1, None),
]


@pytest.mark.parametrize("half", [False, True])
@pytest.mark.parametrize("B,Cin,Cout,D,Hin,Win,act", S2_CASES)
def test_conv_hw3s2_tc2(ops, B, Cin, Cout, D, Hin, Win, act, half):
    """Stride-2 3x3 conv through the phase-decomposed tensor-core kernel vs fp64 (odd sizes: the odd-parity
    phases end one row / column early)."""
    x = rnd(B, Cin, D, Hin, Win, seed=71)
    w = rnd(Cout, Cin, 1, 3, 3, seed=72, scale=(2.0 / (9 * Cin)) ** 0.5)
    b = rnd(Cout, seed=73, scale=0.1)
    want = O._act(F.conv3d(x.double(), w.double(), b.double(), (1, 2, 2), (0, 1, 1)), act).float()
    wp = ops.pack_conv_hw3s2_tc2(w.reshape(Cout, Cin, 9).cuda(), half)
    got = ops.conv_hw3s2_tc2(x.cuda(), wp, b.cuda(), Cout, act, half=half)
    assert got.shape == want.shape
    close(got, want, 1e-5, rtol=1e-5, what="conv_hw3s2_tc2")


@pytest.mark.parametrize("B,Cin,Cout,D,Hin,Win,k,act", [
    (1, 32, 32, 1, 17, 30, 4, "ReLU"), (2, 32, 9, 1, 34, 60, 4, None), (1, 64, 32, 3, 9, 15, 3, None),
    (1, 16, 8, 5, 34, 60, 3, None), (1, 8, 8, 1, 1, 1, 3, None), (1, 40, 40, 2, 5, 33, 4, "SiLU")])
@pytest.mark.parametrize("half", [False, True])
def test_deconv_hw_tc2(ops, B, Cin, Cout, D, Hin, Win, k, act, half):
    """Stride-2 transposed convs (k3 p1 op1 and k4 p1) as four output-phase launches of the tensor-core kernel."""
    x = rnd(B, Cin, D, Hin, Win, seed=74)
    w = rnd(Cin, Cout, 1, k, k, seed=75, scale=(2.0 / (k * k * Cin)) ** 0.5)
    b = rnd(Cout, seed=76, scale=0.1)
    op = 1 if k == 3 else 0
    want = O._act(F.conv_transpose3d(x.double(), w.double(), b.double(), (1, 2, 2), (0, 1, 1), (0, op, op)), act).float()
    wp = ops.pack_deconv_hw_tc2(w.reshape(Cin, Cout, k * k).transpose(0, 1).contiguous().cuda(), k, half)
    got = ops.deconv_hw_tc2(x.cuda(), wp, b.cuda(), Cout, act, half=half)
    assert got.shape == want.shape
    close(got, want, 1e-5, rtol=1e-5, what="deconv_hw_tc2")


def test_conv_s2_deconv_views(ops):
    """Channel-sliced output (the UNet concat buffers) for the stride-2 / transposed forms."""
    x = rnd(1, 16, 12, 20, seed=77)
    w = rnd(8, 16, 3, 3, seed=78, scale=0.1)
    buf = torch.full((1, 24, 6, 10), 7.0, device="cuda")
    ops.conv_hw3s2_tc2(x.cuda(), ops.pack_conv_hw3s2_tc2(w.reshape(8, 16, 9).cuda()), None, 8, "ReLU", out=buf[:, 8:16])
    close(buf[:, 8:16], F.relu(F.conv2d(x, w, None, 2, 1)), 1e-5, rtol=1e-5, what="s2 into slice")
    assert (buf[:, :8] == 7).all() and (buf[:, 16:] == 7).all()
    wt = rnd(16, 8, 4, 4, seed=79, scale=0.1)
    buf = torch.full((1, 24, 24, 40), 7.0, device="cuda")
    ops.deconv_hw_tc2(x.cuda(), ops.pack_deconv_hw_tc2(wt.reshape(16, 8, 16).transpose(0, 1).contiguous().cuda(), 4), None, 8, None,
                      out=buf[:, 8:16])
    close(buf[:, 8:16], F.conv_transpose2d(x, wt, None, 2, 1), 1e-5, rtol=1e-5, what="deconv into slice")
    assert (buf[:, :8] == 7).all() and (buf[:, 16:] == 7).all()


@pytest.mark.parametrize("B,Cin,Cout,Din,hw,k,stride,dil,transposed,act", D_CASES + [
    (1, 8, 16, 5, (136, 240), 3, 1, 1, False, "SiLU"), (2, 64, 64, 6, (17, 30), 3, 2, 1, False, "SiLU"),
    (1, 32, 64, 14, (34, 60), 3, 1, 1, False, "SiLU"), (1, 16, 16, 7, (68, 120), 5, 1, 1, False, "SiLU"),
    (1, 64, 32, 3, (9, 15), 3, 1, 1, True, None), (1, 12, 5, 4, (7, 33), 3, 1, 2, False, "SiLU")])
@pytest.mark.parametrize("half", [False, True])
def test_conv_d_tc2(ops, B, Cin, Cout, Din, hw, k, stride, dil, transposed, act, half):
    """(k,1,1) conv along D through the second-generation tensor-core kernel (planes as K-chunks) vs fp64."""
    H, W = hw
    x = rnd(B, Cin, Din, H, W, seed=47)
    b = rnd(Cout, seed=49, scale=0.1)
    if transposed:
        w = rnd(Cin, Cout, 3, 1, 1, seed=48, scale=0.1)
        want = O._act(F.conv_transpose3d(x.double(), w.double(), b.double(), (2, 1, 1), (1, 0, 0), (1, 0, 0)), act).float()
        wk = w.transpose(0, 1).reshape(Cout, Cin, 3)
    else:
        w = rnd(Cout, Cin, k, 1, 1, seed=48, scale=0.1)
        want = O._act(F.conv3d(x.double(), w.double(), b.double(), (stride, 1, 1), (dil * (k // 2), 0, 0), (dil, 1, 1)), act).float()
        wk = w.reshape(Cout, Cin, k)
    got = ops.conv_d_tc2(x.cuda(), ops.pack_conv_d_tc2(wk.contiguous().cuda(), half), b.cuda(), Cout, k, stride, dil, transposed, act,
                         half=half)
    close(got, want, 1e-5, rtol=1e-5, what="conv_d_tc2")


def test_conv_d_tc2_into_slice(ops):
    x = rnd(1, 16, 7, 9, 14, seed=50)
    w = rnd(16, 16, 5, 1, 1, seed=51, scale=0.1)
    want = F.conv3d(x, w, None, 1, (2, 0, 0))
    cat = torch.zeros(1, 64, 7, 9, 14, device="cuda")
    cat[:, :16] = x.cuda()
    ops.conv_d_tc2(cat[:, :16], ops.pack_conv_d_tc2(w.reshape(16, 16, 5).cuda()), None, 16, 5, 1, 1, False, None, out=cat[:, 16:32])
    close(cat[:, 16:32], want, 1e-5, rtol=1e-5, what="conv_d_tc2 into slice")
    assert (cat[:, 32:] == 0).all()


# --------------------------------------------------------------------------- formats either side of the path (f3, f4)
def test_normalize_u8_matches_totensor_normalize(ops):
    """uint8 HWC -> normalised fp32 CHW, bit-identical to ToTensor().div(255) + Normalize.sub_(mean).div_(std)
    (data/datasets/base.py:120-127), also into a slice of a 2B batch buffer."""
    rng = np.random.RandomState(5)
    img = torch.from_numpy(rng.randint(0, 256, (2, 37, 53, 3)).astype(np.uint8))
    mean = torch.tensor(ops.IMAGENET_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(ops.IMAGENET_STD).view(1, 3, 1, 1)
    want = (img.permute(0, 3, 1, 2).float().div(255).sub(mean)).div(std)
    got = ops.normalize_u8(img.cuda())
    close(got, want, 0.0, what="normalize_u8 (bit-exact)")
    buf = torch.full((4, 3, 37, 53), 9.0, device="cuda")
    ops.normalize_u8(img.cuda(), out=buf[2:])
    assert torch.equal(buf[2:].cpu(), want) and (buf[:2] == 9).all()


@pytest.mark.parametrize("lb,ub", [(None, None), (0.0, 192.0), (5.0, None)])
def test_disp_error_matches_calc_error(ops, lb, ub):
    """calc_error (data/evaluation/pixel_error.py:6-71) restated with torch ops vs the device reduction."""
    gt = rnd(2, 1, 61, 97, seed=11, scale=40.0).abs()
    est = gt + rnd(2, 1, 61, 97, seed=12, scale=2.5)
    mask = torch.ones_like(gt, dtype=torch.bool)
    if lb is not None:
        mask &= gt > lb
    if ub is not None:
        mask &= gt < ub
    err = (gt[mask] - est[mask]).abs()
    want = {"epe": err.mean().item(), **{f"{k}px": 100.0 * (err > k).float().mean().item() for k in (1, 2, 3, 5)}}
    got = ops.error_dict(ops.disp_error(est.cuda(), gt.cuda(), lb, ub))
    assert abs(got["epe"] - want["epe"]) < 1e-5 * max(1.0, want["epe"])
    for k in ("1px", "2px", "3px", "5px"):
        assert abs(got[k] - want[k]) < 1e-4, (k, got[k], want[k])
    # empty mask -> zeros, like the reference
    z = ops.error_dict(ops.disp_error(est.cuda(), gt.cuda(), 1e9, None))
    assert z["epe"] == 0.0 and z["1px"] == 0.0


@pytest.mark.parametrize("Cin,Cout,H,W", [(128, 32, 34, 60), (64, 32, 40, 52), (32, 9, 20, 37)])
def test_single_term_fp16_convs(ops, Cin, Cout, H, W):
    """half=2: one MMA term (A_hi * B_hi): fp16-rounded operands, fp32 accumulate — the UNet decoder's precision.  Against
    the fp64 conv of the fp16-ROUNDED operands it is exact to accumulation rounding; against the fp32 conv it carries the
    operand rounding (2^-11 relative per product)."""
    x = rnd(2, Cin, H, W, seed=81)
    w = rnd(Cout, Cin, 3, 3, seed=82, scale=(2.0 / (9 * Cin)) ** 0.5)
    b = rnd(Cout, seed=83, scale=0.1)
    ws, inv = ops.fp16_prescale(w.reshape(Cout, Cin, 9))
    got = ops.conv_hw3_tc2(x.cuda(), ops.pack_conv_hw3_tc2(ws, True).cuda(), b.cuda(), Cout, 1, "ReLU", half=2, oscale=inv.cuda())
    xr = x.half().double()
    wr = (ws.half().double() * inv.double().view(-1, 1, 1)).reshape(Cout, Cin, 3, 3)
    close(got, F.relu(F.conv2d(xr, wr, b.double(), 1, 1)).float(), 2e-5, rtol=1e-5, what="single-term conv vs rounded-operand reference")
    full = F.relu(F.conv2d(x.double(), w.double(), b.double(), 1, 1)).float()
    err = (got.cpu() - full).abs().max().item()
    assert 1e-6 < err < 5e-3, err
    # transposed 4x4 (deconv4 / deconv2 of the decoder)
    wt = rnd(Cin, Cout, 4, 4, seed=84, scale=(2.0 / (16 * Cin)) ** 0.5)
    wk = wt.reshape(Cin, Cout, 16).transpose(0, 1).contiguous()
    wks, inv = ops.fp16_prescale(wk)
    got = ops.deconv_hw_tc2(x.cuda(), ops.pack_deconv_hw_tc2(wks, 4, True).cuda(), b.cuda(), Cout, None, half=2, oscale=inv.cuda())
    wtr = (wks.half().double() * inv.double().view(-1, 1, 1)).transpose(0, 1).reshape(Cin, Cout, 4, 4)
    close(got, F.conv_transpose2d(xr, wtr, b.double(), 2, 1).float(), 2e-5, rtol=1e-5, what="single-term deconv vs rounded-operand reference")


# --------------------------------------------------------------------------- losses, forward (f2)
@pytest.mark.parametrize("sparse", [False, True])
def test_losses_forward(ops, sparse):
    """On-device smooth-L1 and Wasserstein loss terms (one launch per level, gt scaled + pooled in the kernel) vs the oracle,
    and through the reference-shaped classes; fp32 sums in another order: 2e-6 relative."""
    from oracle.make_golden import loss_inputs
    from temporalstereo_b200.losses import DispSmoothL1Loss, WarssersteinDistanceLoss
    est, costs, offs, smps, gt = loss_inputs(sparse)
    for i, e in enumerate(est):
        want = float(O.smooth_l1_loss_level(e, gt, 192, 0, sparse))
        got = float(ops.loss_smooth_l1(e.cuda(), gt.cuda(), 192, 0, sparse))
        assert abs(got - want) <= 2e-6 * max(1.0, abs(want)), (i, got, want)
    for i, (c, o, s) in enumerate(zip(costs, offs, smps)):
        want = float(O.wasserstein_loss_level(c, o, s, gt, 192, 0, sparse))
        got = float(ops.loss_wasserstein(c.cuda(), o.cuda(), s.cuda(), gt.cuda(), 192, 0, sparse))
        assert abs(got - want) <= 2e-6 * max(1.0, abs(want)), (i, got, want)
    assert float(ops.loss_smooth_l1(est[1].cuda(), torch.zeros_like(gt).cuda(), 192, 0, sparse)) == 0.0
    d1 = DispSmoothL1Loss(192, 0, 0.5, [1.0, 0.7, 0.5, 0.3], sparse)([e.cuda() for e in est], gt.cuda())
    assert set(d1) == {f"l1_loss_lvl{i}" for i in range(4)}
    assert abs(float(d1["l1_loss_lvl1"]) - 0.5 * 0.7 * float(O.smooth_l1_loss_level(est[1], gt, 192, 0, sparse))) < 1e-4
    d2 = WarssersteinDistanceLoss(192, 0, 1.0, None, sparse)([c.cuda() for c in costs], [o.cuda() for o in offs], [s.cuda() for s in smps], gt.cuda())
    assert set(d2) == {f"wars_loss_lvl{i}" for i in range(3)}
