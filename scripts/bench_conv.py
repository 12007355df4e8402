"""Times the conv variants (fp32 FMA, tcgen05 kx-folded) on the aggregation's layer shapes and
prints error statistics vs fp64 for the accumulation cadence G (TSTEREO_TC2_G)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from temporalstereo_b200 import ops

B = int(os.environ.get("BATCH", "4"))
SHAPES = [  # name, Cin, Cout, D, H, W, dil
    ("precise.first 304->8", 304, 8, 5, 136, 240, 1),
    ("fine.first 304->16", 304, 16, 5, 68, 120, 1),
    ("coarse.first 352->32", 352, 32, 12, 34, 60, 1),
    ("unet 32->32 @1/2", 32, 32, 1, 272, 480, 1),
    ("unet fuse 128->32 @1/4", 128, 32, 1, 136, 240, 1),
    ("unet concat 64->32 @1/2", 64, 32, 1, 272, 480, 1),
    ("precise 8->8 @1/4", 8, 8, 5, 136, 240, 1),
    ("precise 8->8 dil2", 8, 8, 5, 136, 240, 2),
    ("fine 16->16", 16, 16, 5, 68, 120, 1),
    ("fine fuse 64->16", 64, 16, 7, 68, 120, 1),
    ("coarse 32->32", 32, 32, 12, 34, 60, 1),
    ("coarse fuse 128->32", 128, 32, 14, 34, 60, 1),
    ("hourglass 32->32 @1/32", 32, 32, 6, 17, 30, 1),
    ("hourglass 64->64 @1/32", 64, 64, 6, 17, 30, 1),
    ("hourglass 64->64 @1/64", 64, 64, 3, 9, 15, 1),
    ("fine hg 32->32 @1/16", 32, 32, 3, 34, 60, 1),
]


def pack_simt(w):
    cout, cin = w.shape[:2]
    w = w.reshape(cout, cin, -1)
    p = torch.zeros(cin, w.shape[2], (cout + 3) // 4 * 4, device=w.device)
    p[:, :, :cout] = w.permute(1, 2, 0)
    return p.contiguous()


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for name, Cin, Cout, D, H, W, dil in SHAPES:
    g = torch.Generator(device="cuda").manual_seed(0)
    # rotate over enough inputs to exceed L2
    nbuf = max(2, int(300e6 // (B * Cin * D * H * W * 4)) + 1)
    nbuf = min(nbuf, 16)
    xs = [torch.randn(B, Cin, D, H, W, device="cuda", generator=g) for _ in range(nbuf)]
    w = torch.randn(Cout, Cin, 1, 3, 3, device="cuda", generator=g) * (2.0 / (9 * Cin)) ** 0.5
    bias = torch.randn(Cout, device="cuda", generator=g) * 0.1
    ws, w2 = pack_simt(w), ops.pack_conv_hw3_tc2(w.reshape(Cout, Cin, 9))
    w2h = ops.pack_conv_hw3_tc2(w.reshape(Cout, Cin, 9), True)
    out = torch.empty(B, Cout, D, H, W, device="cuda")
    i = [0]

    def nxt():
        i[0] += 1
        return xs[i[0] % nbuf]
    t_simt = timed(lambda: ops.conv_hw3(nxt(), ws, bias, Cout, 1, dil, "SiLU", out=out))
    res = {}
    for mt in ("2", "4"):
        if Cout > 16 and mt == "4":
            continue
        os.environ["TSTEREO_TC2_MT"] = mt
        res[mt] = timed(lambda: ops.conv_hw3_tc2(nxt(), w2, bias, Cout, dil, "SiLU", out=out))
    os.environ.pop("TSTEREO_TC2_MT")
    t_h = timed(lambda: ops.conv_hw3_tc2(nxt(), w2h, bias, Cout, dil, "SiLU", out=out, half=True))
    gflop = 2.0 * B * Cin * Cout * 9 * D * H * W / 1e9
    mb = 4.0 * B * (Cin + Cout) * D * H * W / 1e6
    best = min(res.values())
    print(f"{name:32s} B={B} {gflop:7.2f} GFLOP {mb:7.1f} MB | fma {t_simt:7.1f} us  tc2 " +
          " ".join(f"MT{k}={v:7.1f}" for k, v in res.items()) +
          f" us | f16 {t_h:7.1f} us {gflop / t_h * 1e-3:6.1f} TFLOP/s {mb / t_h * 1e-3:5.2f} TB/s")

# accumulation cadence: error vs fp64 at the largest K
x = torch.randn(1, 352, 2, 34, 60, device="cuda")
w = torch.randn(32, 352, 1, 3, 3, device="cuda") * (2.0 / (9 * 352)) ** 0.5
want = F.conv3d(x.double().cpu(), w.double().cpu(), None, 1, (0, 1, 1))
fma = ops.conv_hw3(x, pack_simt(w), None, 32, 1, 1, None).double().cpu()
print(f"K=9*352 fp32 FMA      rms err {(fma - want).pow(2).mean().sqrt():.2e} max {(fma - want).abs().max():.2e}")
w2h = ops.pack_conv_hw3_tc2(w.reshape(32, 352, 9), True)
got = ops.conv_hw3_tc2(x, w2h, None, 32, 1, None, half=True).double().cpu()
e = got - want
print(f"K=9*352 tc2 fp16 hi+lo  rms err {e.pow(2).mean().sqrt():.2e} max {e.abs().max():.2e} bias {(e * torch.sign(want)).mean():+.2e}")
w2 = ops.pack_conv_hw3_tc2(w.reshape(32, 352, 9))
for G in (1, 4, 8):
    os.environ["TSTEREO_TC2_G"] = str(G)
    got = ops.conv_hw3_tc2(x, w2, None, 32, 1, None).double().cpu()
    e = got - want
    print(f"K=9*352 tc2 G={G:3d}     rms err {e.pow(2).mean().sqrt():.2e} max {e.abs().max():.2e} bias {(e * torch.sign(want)).mean():+.2e}")
