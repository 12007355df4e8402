"""S-format (TMA-fed) tensor-core convolutions: parity with the fp32-input forms and with fp64 torch convolutions.

An S-format tensor holds the fp16 hi / lo halves of an fp32 activation (include/tstereo.h `tstereo_split`); the kernels
that read it stage a K-chunk with one TMA box instead of converting fp32 in producer warps.  Operand values and MMA
order are the same as in the fp32-input forms, so the results must be BIT-IDENTICAL to them (asserted with torch.equal),
and within the written tolerance (1e-5 abs + 1e-5 rel) of the fp64 convolution like every other contraction.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import oracle as O

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


def close(got, want, atol, rtol=0.0, what=""):
    got = got.detach().cpu()
    assert got.shape == want.shape, (what, got.shape, want.shape)
    assert torch.isfinite(got).all(), f"{what}: non-finite output"
    torch.testing.assert_close(got, want, atol=atol, rtol=rtol, msg=lambda m: f"{what}: {m}")


def same_split(a, b, what=""):
    """Two S-format tensors hold the same bit patterns."""
    assert a.shape == b.shape, (what, a.shape, b.shape)
    assert torch.equal(a.t.view(torch.int16), b.t.view(torch.int16)), f"{what}: S-format bits differ"


@pytest.mark.parametrize("shape", [(2, 32, 9, 20), (1, 13, 3, 5, 7), (1, 8, 1, 1, 1), (2, 44, 2, 17, 30)])
@pytest.mark.parametrize("parts", [1, 2])
def test_split_pack(ops, shape, parts):
    """hi = fp16(x) exactly (round to nearest even), lo = fp16(x - hi); padding channels are zero."""
    x = rnd(*shape, seed=1, scale=3.0).cuda()
    s = ops.split_pack(x, parts)
    B, C = shape[:2]
    hi = x.half()
    xs = x if x.dim() == 5 else x.unsqueeze(2)
    his = hi if x.dim() == 5 else hi.unsqueeze(2)
    C8 = (C + 7) // 8
    pad = torch.zeros(B, C8 * 8, *xs.shape[2:], device="cuda", dtype=torch.float16)
    pad[:, :C] = his
    want_hi = pad.view(B, C8, 8, *xs.shape[2:]).permute(0, 3, 1, 4, 5, 2)      # [B, D, C8, H, W, 8]
    assert torch.equal(s.t[:, :, 0], want_hi)
    if parts == 2:
        pad32 = torch.zeros(B, C8 * 8, *xs.shape[2:], device="cuda")
        pad32[:, :C] = xs - his.float()
        want_lo = pad32.half().view(B, C8, 8, *xs.shape[2:]).permute(0, 3, 1, 4, 5, 2)
        assert torch.equal(s.t[:, :, 1], want_lo)
        assert (s.float() - x).abs().max() <= x.abs().max() * 2.0 ** -21
    else:
        assert torch.equal(s.float(), hi.float())


S_HW3 = [
    # B, Cin, Cout, D, H, W, dil, act
    (1, 8, 16, 1, 8, 16, 1, None),          # one 8-channel unit: the second K half is a box of zeros
    (1, 16, 8, 2, 9, 20, 1, "SiLU"),
    (1, 44, 32, 3, 17, 30, 1, "SiLU"),      # padded chunk, odd number of units
    (2, 32, 32, 1, 40, 70, 1, "ReLU"),      # the UNet 32 -> 32 form, several tiles per CTA
    (1, 32, 32, 2, 17, 30, 2, "SiLU"),      # dilation 2
    (1, 128, 32, 2, 34, 60, 1, None),       # accumulation groups (ACC mode)
    (1, 304, 8, 2, 20, 37, 1, "SiLU"),
    (1, 64, 64, 1, 33, 50, 1, "ReLU"),      # two output-channel groups
    (2, 64, 9, 1, 12, 33, 1, None),         # Cout not a multiple of 8
    (1, 16, 16, 3, 5, 3, 2, "SiLU"),
    (2, 32, 3, 1, 1, 1, 1, None),
    (1, 8, 8, 5, 136, 240, 2, "SiLU"),
]


@pytest.mark.parametrize("half", [1, 2])
@pytest.mark.parametrize("B,Cin,Cout,D,H,W,dil,act", S_HW3)
def test_conv_hw3_s_input(ops, B, Cin, Cout, D, H, W, dil, act, half):
    """S-format input through TMA == the fp32-input kernel bit for bit; both == fp64 conv to fp32 rounding (3 terms)."""
    x = rnd(B, Cin, D, H, W, seed=141).cuda()
    w = rnd(Cout, Cin, 1, 3, 3, seed=142, scale=(2.0 / (9 * Cin)) ** 0.5)
    b = rnd(Cout, seed=143, scale=0.1).cuda()
    wp = ops.pack_conv_hw3_tc2(w.reshape(Cout, Cin, 9).cuda(), True)
    ref = ops.conv_hw3_tc2(x, wp, b, Cout, dil, act, half=half)
    xs = ops.split_pack(x, 2 if half == 1 else 1)
    got, _ = ops.conv_hw3_s(xs, wp, b, Cout, dil, act, half=half)
    assert torch.equal(got, ref), f"max diff {(got - ref).abs().max().item():.3e}"
    if half == 1:
        want = O._act(F.conv3d(x.cpu().double(), w.double(), b.cpu().double(), 1, (0, dil, dil), (1, dil, dil)), act).float()
        close(got, want, 1e-5, rtol=1e-5, what="conv_hw3_s")


@pytest.mark.parametrize("parts", [1, 2])
@pytest.mark.parametrize("B,Cin,Cout,D,H,W,dil,act", S_HW3[:9])
def test_conv_hw3_s_output(ops, B, Cin, Cout, D, H, W, dil, act, parts):
    """The epilogue's S-format output == split_pack of its fp32 output, from either input format."""
    x = rnd(B, Cin, D, H, W, seed=151).cuda()
    w = rnd(Cout, Cin, 1, 3, 3, seed=152, scale=(2.0 / (9 * Cin)) ** 0.5)
    b = rnd(Cout, seed=153, scale=0.1).cuda()
    wp = ops.pack_conv_hw3_tc2(w.reshape(Cout, Cin, 9).cuda(), True)
    ref = ops.conv_hw3_tc2(x, wp, b, Cout, dil, act, half=True)
    want = ops.split_pack(ref, parts)
    for src in (x, ops.split_pack(x)):
        so = ops.Split(B, Cout, D, H, W, parts, device="cuda")
        so.t.fill_(7.0)
        out, _ = ops.conv_hw3_s(src, wp, b, Cout, dil, act, half=1, sout=so, want_f32=True)
        assert torch.equal(out, ref)
        same_split(so, want, "conv_hw3_s sout")
        so2 = ops.Split(B, Cout, D, H, W, parts, device="cuda")
        out2, _ = ops.conv_hw3_s(src, wp, b, Cout, dil, act, half=1, sout=so2)     # S-format only
        assert out2 is None
        same_split(so2, want, "conv_hw3_s sout only")


def test_conv_hw3_s_concat_and_partial_batch(ops):
    """Two producers write chunk ranges of one S-format tensor (a channel concat); a consumer reads it; `nb` limits the
    S-format output to the leading batches."""
    B, H, W = 3, 20, 45
    xa, xb = rnd(B, 16, H, W, seed=161).cuda(), rnd(B, 24, H, W, seed=162).cuda()
    wa, wb = rnd(32, 16, 3, 3, seed=163, scale=0.1), rnd(32, 24, 3, 3, seed=164, scale=0.1)
    wc = rnd(16, 64, 3, 3, seed=165, scale=0.05)
    pa = ops.pack_conv_hw3_tc2(wa.reshape(32, 16, 9).cuda(), True)
    pb = ops.pack_conv_hw3_tc2(wb.reshape(32, 24, 9).cuda(), True)
    pc = ops.pack_conv_hw3_tc2(wc.reshape(16, 64, 9).cuda(), True)
    cat = ops.Split(B, 64, 1, H, W, 2, device="cuda", five=False)
    cat.t.zero_()
    fa, _ = ops.conv_hw3_s(xa, pa, None, 32, 1, "ReLU", sout=cat.channels(0, 32), want_f32=True)
    fb, _ = ops.conv_hw3_s(xb, pb, None, 32, 1, "ReLU", sout=cat.channels(32, 64), want_f32=True, nb=2)
    fcat = torch.cat([fa, fb], 1)
    fcat[2:, 32:] = 0                                     # batch 2 of the second producer was not written
    same_split(cat, ops.split_pack(fcat), "concat")
    got, _ = ops.conv_hw3_s(cat, pc, None, 16, 1, None)
    ref = ops.conv_hw3_tc2(fcat, pc, None, 16, 1, None, half=True)
    assert torch.equal(got, ref)
    # a consumer of a chunk slice and of a batch slice
    p8 = ops.pack_conv_hw3_tc2(rnd(8, 32, 9, seed=166, scale=0.1).cuda(), True)
    got2, _ = ops.conv_hw3_s(cat.channels(32, 64).batches(1, 3), p8, None, 8, 1, None)
    ref2 = ops.conv_hw3_tc2(fcat[1:3, 32:], p8, None, 8, 1, None, half=True)
    assert torch.equal(got2, ref2)


S_S2 = [
    # B, Cin, Cout, D, Hin, Win, act
    (2, 8, 16, 3, 18, 30, "SiLU"),
    (1, 16, 16, 2, 9, 15, "SiLU"),          # odd sizes: the odd-parity phases end one row / column early
    (1, 32, 64, 1, 34, 61, "ReLU"),
    (1, 64, 64, 4, 17, 30, "SiLU"),
    (2, 32, 32, 1, 68, 120, None),
    (1, 12, 20, 1, 5, 7, None),
    (1, 8, 8, 1, 1, 1, None),
]


@pytest.mark.parametrize("B,Cin,Cout,D,Hin,Win,act", S_S2)
def test_conv_hw3s2_s(ops, B, Cin, Cout, D, Hin, Win, act):
    """Stride-2 conv reading the four parity phases of an S-format input through the TMA map's element strides."""
    x = rnd(B, Cin, D, Hin, Win, seed=171).cuda()
    w = rnd(Cout, Cin, 1, 3, 3, seed=172, scale=(2.0 / (9 * Cin)) ** 0.5)
    b = rnd(Cout, seed=173, scale=0.1).cuda()
    wp = ops.pack_conv_hw3s2_tc2(w.reshape(Cout, Cin, 9).cuda(), True)
    ref = ops.conv_hw3s2_tc2(x, wp, b, Cout, act, half=True)
    H, W = ref.shape[-2:]
    so = ops.Split(B, Cout, D, H, W, 2, device="cuda")
    got, _ = ops.conv_hw3s2_s(ops.split_pack(x), wp, b, Cout, act, sout=so, want_f32=True)
    assert torch.equal(got, ref), f"max diff {(got - ref).abs().max().item():.3e}"
    same_split(so, ops.split_pack(ref), "conv_hw3s2_s sout")
    want = O._act(F.conv3d(x.cpu().double(), w.double(), b.cpu().double(), (1, 2, 2), (0, 1, 1)), act).float()
    close(got, want, 1e-5, rtol=1e-5, what="conv_hw3s2_s")


@pytest.mark.parametrize("half", [1, 2])
@pytest.mark.parametrize("B,Cin,Cout,D,Hin,Win,k,act", [
    (1, 32, 32, 1, 17, 30, 4, "ReLU"), (2, 32, 9, 1, 34, 60, 4, None), (1, 64, 32, 3, 9, 15, 3, None),
    (1, 16, 8, 5, 34, 60, 3, None), (1, 8, 8, 1, 1, 1, 3, None), (1, 40, 40, 2, 5, 33, 4, "SiLU")])
def test_deconv_hw_s(ops, B, Cin, Cout, D, Hin, Win, k, act, half):
    """Stride-2 transposed convs: S-format in, both formats out (one output parity phase per launch)."""
    x = rnd(B, Cin, D, Hin, Win, seed=174).cuda()
    w = rnd(Cin, Cout, 1, k, k, seed=175, scale=(2.0 / (k * k * Cin)) ** 0.5)
    b = rnd(Cout, seed=176, scale=0.1).cuda()
    wp = ops.pack_deconv_hw_tc2(w.reshape(Cin, Cout, k * k).transpose(0, 1).contiguous().cuda(), k, True)
    ref = ops.deconv_hw_tc2(x, wp, b, Cout, act, half=half)
    parts = 2 if half == 1 else 1
    so = ops.Split(B, Cout, D, 2 * Hin, 2 * Win, parts, device="cuda")
    got, _ = ops.deconv_hw_s(ops.split_pack(x, parts), wp, b, Cout, act, half=half, sout=so, want_f32=True)
    assert torch.equal(got, ref), f"max diff {(got - ref).abs().max().item():.3e}"
    same_split(so, ops.split_pack(ref, parts), "deconv_hw_s sout")


@pytest.mark.parametrize("B,Cin,Cout,Din,hw,k,stride,dil,transposed,act", [
    (1, 32, 32, 12, (34, 60), 3, 1, 1, False, "SiLU"), (2, 64, 64, 6, (17, 30), 3, 2, 1, False, "SiLU"),
    (1, 8, 16, 5, (36, 70), 3, 1, 1, False, "SiLU"), (1, 16, 16, 7, (20, 33), 5, 1, 1, False, "SiLU"),
    (1, 64, 32, 3, (9, 15), 3, 1, 1, True, None), (1, 12, 5, 4, (7, 33), 3, 1, 2, False, "SiLU"),
    (1, 32, 32, 14, (9, 14), 3, 1, 2, False, None), (2, 16, 16, 3, (5, 40), 3, 2, 1, False, None)])
def test_conv_d_s(ops, B, Cin, Cout, Din, hw, k, stride, dil, transposed, act):
    """(k,1,1) conv along D from an S-format input: planes outside [0, Din) are boxes of zeros from the TMA unit."""
    H, W = hw
    x = rnd(B, Cin, Din, H, W, seed=147).cuda()
    b = rnd(Cout, seed=149, scale=0.1).cuda()
    if transposed:
        w = rnd(Cin, Cout, 3, seed=148, scale=0.1)
        wk = w.transpose(0, 1).contiguous()
    else:
        wk = rnd(Cout, Cin, k, seed=148, scale=0.1)
    wp = ops.pack_conv_d_tc2(wk.cuda(), True)
    ref = ops.conv_d_tc2(x, wp, b, Cout, k, stride, dil, transposed, act, half=True)
    so = ops.Split(*ref.shape, 2, device="cuda")
    got, _ = ops.conv_d_s(ops.split_pack(x), wp, b, Cout, k, stride, dil, transposed, act, sout=so, want_f32=True)
    assert torch.equal(got, ref), f"max diff {(got - ref).abs().max().item():.3e}"
    same_split(so, ops.split_pack(ref), "conv_d_s sout")


def test_split_rejects_bad_arguments(ops):
    from temporalstereo_b200._lib import TStereoError
    x = rnd(1, 16, 8, 16, seed=1).cuda()
    wp = ops.pack_conv_hw3_tc2(rnd(8, 16, 9, seed=2).cuda(), True)
    with pytest.raises(TStereoError):       # a hi-only input cannot feed the three-term form
        ops.conv_hw3_s(ops.split_pack(x, 1), wp, None, 8, 1, None, half=1)
    with pytest.raises(ValueError):         # wrong-sized S-format output
        ops.conv_hw3_s(x, wp, None, 8, 1, None, sout=ops.Split(1, 8, 1, 8, 15, 2, device="cuda", five=False))
    with pytest.raises(TStereoError):       # S-format input needs the fp16 split
        ops.conv_hw3_s(ops.split_pack(x), wp, None, 8, 1, None, half=0)


@pytest.mark.parametrize("src,dst,C", [((6, 18, 30), (6, 17, 30), 64), ((4, 10, 16), (3, 9, 15), 8), ((6, 18, 30), (5, 17, 30), 12),
                                       ((2, 5, 7), (2, 5, 7), 16)])
@pytest.mark.parametrize("parts", [1, 2])
def test_resize_add_act_s(ops, src, dst, C, parts):
    """resize + add + SiLU writing S-format: the fp32 operator's result (its SiLU is the full-precision form, this kernel's
    the ex2 / rcp approximation of the conv epilogues: 4e-6 abs) in a well-formed split; without an activation bit for bit."""
    a = rnd(2, C, *src, seed=181).cuda()
    skip = rnd(2, C, *dst, seed=182).cuda()
    got = ops.resize_add_act_s(a, dst, skip, "SiLU", parts)
    want = ops.resize_add_act(a, dst, skip, "SiLU")
    tol = 4e-6 if parts == 2 else 4e-3                    # hi half alone carries 11 bits
    assert (got.float() - want).abs().max().item() <= tol * max(1.0, want.abs().max().item())
    same_split(ops.resize_add_act_s(a, dst, None, None, parts), ops.split_pack(ops.resize_add_act(a, dst, None, None), parts), "no skip")


@pytest.mark.parametrize("B,C,H,W,D", [(1, 256, 20, 36, 12), (2, 64, 9, 13, 20), (1, 64, 34, 60, 16)])
def test_block_cost_shift_s(ops, B, C, H, W, D):
    """The coarse shift volume written directly in S-format == split_pack of the materialising operator's fp32 volume."""
    L, R = rnd(B, C, H, W, seed=191).cuda(), rnd(B, C, H, W, seed=192).cuda()
    want = ops.split_pack(ops.block_cost(L, R, D))
    got = ops.block_cost_shift_s(L, R, D)
    same_split(got, want, "block_cost_shift_s")
