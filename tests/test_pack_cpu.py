"""Host-side weight repacking of the tensor-core convolutions (no GPU): the "virtual stride-1 conv" forms that let
one kernel run stride-2 convs and stride-2 transposed convs must be exactly equivalent to the torch operators, and
the packed operand images must have the sizes the C ABI announces and split every weight into hi + lo parts."""
import pytest
import torch
import torch.nn.functional as F

from temporalstereo_b200 import _lib, ops


def _s2d(x, cin8):
    """[B, C, H, W] -> the four parity phases stacked on the channel axis, phase-major, each padded to cin8 channels."""
    B, C, H, W = x.shape
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    out = torch.zeros(B, 4, cin8, Ho, Wo, dtype=x.dtype)
    for pr in (0, 1):
        for pc in (0, 1):
            ph = x[:, :, pr::2, pc::2]
            out[:, pr * 2 + pc, :C, :ph.shape[2], :ph.shape[3]] = ph
    return out.reshape(B, 4 * cin8, Ho, Wo)


@pytest.mark.parametrize("cin,cout,H,W", [(3, 5, 12, 16), (8, 4, 9, 15), (13, 7, 7, 5), (16, 16, 2, 2)])
def test_stride2_conv_as_virtual_stride1(cin, cout, H, W):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, cin, H, W, generator=g, dtype=torch.float64)
    w = torch.randn(cout, cin, 3, 3, generator=g, dtype=torch.float64)
    want = F.conv2d(x, w, None, 2, 1)
    virt = ops.virtual_weights_s2(w.reshape(cout, cin, 9).float()).double()
    cin8 = (cin + 7) // 8 * 8
    got = F.conv2d(_s2d(x, cin8), virt, None, 1, 1)
    assert got.shape == want.shape
    assert torch.allclose(got, want, atol=1e-5, rtol=1e-6)      # weights went through fp32


@pytest.mark.parametrize("k", [3, 4])
@pytest.mark.parametrize("cin,cout,H,W", [(6, 5, 7, 9), (8, 9, 1, 1), (4, 3, 5, 4)])
def test_transposed_conv_as_four_phase_convs(k, cin, cout, H, W):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, cin, H, W, generator=g, dtype=torch.float64)
    wt = torch.randn(cin, cout, k, k, generator=g, dtype=torch.float64)
    want = F.conv_transpose2d(x, wt, None, 2, 1, 1 if k == 3 else 0)
    assert want.shape[-2:] == (2 * H, 2 * W)
    phases = ops.virtual_weights_deconv(wt.transpose(0, 1).reshape(cout, cin, k * k).float(), k).double()
    got = torch.zeros_like(want)
    for py in (0, 1):
        for px in (0, 1):
            got[:, :, py::2, px::2] = F.conv2d(x, phases[py * 2 + px], None, 1, 1)
    assert torch.allclose(got, want, atol=1e-5, rtol=1e-6)


@pytest.mark.parametrize("half", [False, True])
def test_operand_images_match_the_abi_sizes_and_reconstruct_the_weights(half):
    lib = _lib.load()
    g = torch.Generator().manual_seed(2)
    for cin, cout in [(8, 8), (13, 20), (304, 8), (64, 64), (32, 36)]:
        w = torch.randn(cout, cin, 9, generator=g) * 0.1
        p = ops.pack_conv_hw3_tc2(w, half)
        assert p.dtype == torch.float32 and p.numel() == lib.tstereo_conv_hw3_tc2_wpack_floats(cin, cout, int(half))
        assert ops.pack_conv_hw3s2_tc2(w, half).numel() == lib.tstereo_conv_hw3s2_tc2_wpack_floats(cin, cout, int(half))
        assert ops.pack_deconv_hw_tc2(w, 3, half).numel() == lib.tstereo_deconv_hw_tc2_wpack_floats(cin, cout, int(half))
        wd = torch.randn(cout, cin, 5, generator=g) * 0.1
        assert ops.pack_conv_d_tc2(wd, half).numel() == lib.tstereo_conv_d_tc2_wpack_floats(cin, cout, 5, int(half))
    # first output-channel group of a small conv: hi + lo reproduces the weights to 2^-21 (tf32) / 2^-21 (fp16) relative
    cout, cin = 8, 16
    w = torch.randn(cout, cin, 9, generator=g) * 0.1
    p = ops.pack_conv_hw3_tc2(w, half)
    CP, N = 8, 24
    if half:
        img = p.view(torch.float16).view(1, 3, 2, 2 * N, 8).float()          # [chunk, ky, khalf, row, i]
    else:
        img = p.view(2, 3, 2, 2 * N, 4)
    rec = torch.zeros(cout, cin, 3, 3)
    per = 16 if half else 8
    for chunk in range(img.shape[0]):
        for ky in range(3):
            for kh in range(2):
                for kx in range(3):
                    rows_hi = img[chunk, ky, kh, kx * CP:kx * CP + cout]          # [co, i]
                    rows_lo = img[chunk, ky, kh, N + kx * CP:N + kx * CP + cout]
                    c0 = chunk * per + kh * (per // 2)
                    rec[:, c0:c0 + per // 2, ky, kx] = rows_hi + rows_lo
    err = (rec.reshape(cout, cin, 9) - w).abs().max() / w.abs().max()
    assert err < 2e-6, err


def test_tf32_split_is_exact_and_representable():
    g = torch.Generator().manual_seed(3)
    w = torch.randn(1000, generator=g)
    hi, lo = ops.tf32_split(w)
    for t in (hi, lo):
        assert ((t.view(torch.int32) & 0x1FFF) == 0).all(), "parts must be exactly representable in tf32"
    assert ((hi + lo) - w).abs().max() <= w.abs().max() * 2.0 ** -21


def test_fp16_prescale_is_an_exact_power_of_two_per_channel():
    g = torch.Generator().manual_seed(3)
    w = torch.randn(12, 40, 9, generator=g) * torch.logspace(-5, 1, 12).view(-1, 1, 1)
    ws, inv = ops.fp16_prescale(w)
    assert torch.equal(torch.log2(inv), torch.log2(inv).round())
    assert torch.equal(ws * inv.view(-1, 1, 1), w)                   # exact round trip
    m = ws.abs().flatten(1).amax(1)
    assert (m > 511).all() and (m <= 1023).all()
    # hi + lo of the scaled weights reconstructs >= 21 bits of every weight within 2^-10 of its channel's largest
    hi = ws.half()
    lo = (ws - hi.float()).half()
    rec = (hi.float() + lo.float())
    big = ws.abs() > m.view(-1, 1, 1) * 2.0 ** -10
    assert ((rec - ws).abs()[big] <= ws.abs()[big] * 2.0 ** -21).all()
    # unscaled, the small channels lose their lo half to fp16's subnormal range
    hi0 = w.half()
    rec0 = hi0.float() + (w - hi0.float()).half().float()
    assert ((rec0 - w).abs() / w.abs().clamp_min(1e-30))[0].max() > 2.0 ** -15
