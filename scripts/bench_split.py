"""Times the S-format (TMA-fed) tensor-core convolutions against the fp32-input forms on the engine's layer shapes
(B = 8 frames of C2, 544x960).  Each timing is a CUDA-graph replay of the launches over inputs rotated beyond L2."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from temporalstereo_b200 import ops

B = int(os.environ.get("BATCH", "8"))
SHAPES = [  # name, kind, Cin, Cout, planes-multiplier(B x), D, H, W, dil, act, half
    ("unet conv2.1 32->32 @1/2 (L+R)", "hw3", 32, 32, 2, 1, 272, 480, 1, "ReLU", 1),
    ("unet conv4.0 32->64 s2 (L+R)", "s2", 32, 64, 2, 1, 272, 480, 1, "ReLU", 1),
    ("unet conv4.1 64->64 @1/4 (L+R)", "hw3", 64, 64, 2, 1, 136, 240, 1, "ReLU", 1),
    ("unet fuse.0 128->32 @1/4 1-term", "hw3", 128, 32, 1, 1, 136, 240, 1, "ReLU", 2),
    ("unet fuse.1 32->32 @1/4 1-term", "hw3", 32, 32, 1, 1, 136, 240, 1, "ReLU", 2),
    ("unet deconv4 32->32 1-term", "dc4", 32, 32, 1, 1, 136, 240, 1, "ReLU", 2),
    ("unet concat 64->32 @1/2 1-term", "hw3", 64, 32, 1, 1, 272, 480, 1, "ReLU", 2),
    ("unet deconv2 32->9 1-term", "dc4", 32, 9, 1, 1, 272, 480, 1, None, 2),
    ("precise 8->16 s2 (hourglass conv1)", "s2", 8, 16, 1, 5, 136, 240, 1, "SiLU", 1),
    ("precise 8->8 shortcut6", "hw3", 8, 8, 1, 5, 136, 240, 1, None, 1),
    ("precise 8->8 dil2", "hw3", 8, 8, 1, 5, 136, 240, 2, "SiLU", 1),
    ("coarse fuse 128->32", "hw3", 128, 32, 1, 14, 34, 60, 1, None, 1),
    ("coarse 32->32", "hw3", 32, 32, 1, 12, 34, 60, 1, "SiLU", 1),
    ("coarse d 32->32 k3", "d", 32, 32, 1, 12, 34, 60, 1, "SiLU", 1),
    ("fine d 16->16 k3", "d", 16, 16, 1, 5, 68, 120, 1, "SiLU", 1),
    ("precise d 8->8 k3", "d", 8, 8, 1, 5, 136, 240, 1, "SiLU", 1),
]


def timed(fns, reps=5):
    """fns: list of thunks (one per rotated input set), replayed as one graph."""
    for f in fns:
        f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for f in fns:
            f()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * len(fns)) * 1e3


print(f"B = {B}; us per launch (graph replay, inputs rotate beyond L2)")
MODES = [{}, {"TSTEREO_TC2_MT": "2"}, {"TSTEREO_TC2_MT": "2", "TSTEREO_TC2_WRES": "0"}, {"TSTEREO_TC2_WRES": "0"}]
print(f"{'layer':40s} {'fp32 in':>9s} {'S->S':>9s} {'S->S mt2':>9s} {'mt2 nowr':>9s} {'nowres':>9s} {'S->S+f32':>9s}  MB(fp32 io)")
for name, kind, Cin, Cout, mul, D, H, W, dil, act, half in SHAPES:
    gen = torch.Generator(device="cuda").manual_seed(0)
    Bn = B * mul
    nbuf = min(6, max(2, int(400e6 // (Bn * Cin * D * H * W * 4)) + 1))
    five = D > 1 or kind == "d"
    shp = (Bn, Cin, D, H, W) if five else (Bn, Cin, H, W)
    xs = [torch.randn(*shp, device="cuda", generator=gen) for _ in range(nbuf)]
    parts = 2 if half == 1 else 1
    ss = [ops.split_pack(x, parts) for x in xs]
    bias = torch.randn(Cout, device="cuda", generator=gen) * 0.1
    if kind == "hw3":
        w = torch.randn(Cout, Cin, 9, device="cuda", generator=gen) * (2.0 / (9 * Cin)) ** 0.5
        wp = ops.pack_conv_hw3_tc2(w, True)
        Ho, Wo = H, W
        f32 = lambda x, out: ops.conv_hw3_tc2(x, wp, bias, Cout, dil, act, out=out, half=half)
        sfn = lambda s, out, so: ops.conv_hw3_s(s, wp, bias, Cout, dil, act, out=out, half=half, sout=so)
    elif kind == "s2":
        w = torch.randn(Cout, Cin, 9, device="cuda", generator=gen) * (2.0 / (9 * Cin)) ** 0.5
        wp = ops.pack_conv_hw3s2_tc2(w, True)
        Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        f32 = lambda x, out: ops.conv_hw3s2_tc2(x, wp, bias, Cout, act, out=out, half=half)
        sfn = lambda s, out, so: ops.conv_hw3s2_s(s, wp, bias, Cout, act, out=out, half=half, sout=so)
    elif kind == "dc4":
        w = torch.randn(Cout, Cin, 16, device="cuda", generator=gen) * (2.0 / (16 * Cin)) ** 0.5
        wp = ops.pack_deconv_hw_tc2(w, 4, True)
        Ho, Wo = 2 * H, 2 * W
        f32 = lambda x, out: ops.deconv_hw_tc2(x, wp, bias, Cout, act, out=out, half=half)
        sfn = lambda s, out, so: ops.deconv_hw_s(s, wp, bias, Cout, act, out=out, half=half, sout=so)
    else:
        w = torch.randn(Cout, Cin, 3, device="cuda", generator=gen) * (2.0 / (3 * Cin)) ** 0.5
        wp = ops.pack_conv_d_tc2(w, True)
        Ho, Wo = H, W
        f32 = lambda x, out: ops.conv_d_tc2(x, wp, bias, Cout, 3, 1, 1, False, act, out=out, half=half)
        sfn = lambda s, out, so: ops.conv_d_s(s, wp, bias, Cout, 3, 1, 1, False, act, out=out, half=half, sout=so)
    oshp = (Bn, Cout, D, Ho, Wo) if five else (Bn, Cout, Ho, Wo)
    out = torch.empty(*oshp, device="cuda")
    so = ops.Split(Bn, Cout, D, Ho, Wo, parts, device="cuda", five=five)
    t0 = timed([(lambda x=x: f32(x, out)) for x in xs])
    cols = []
    for env in MODES:
        for k_ in ("TSTEREO_TC2_MT", "TSTEREO_TC2_WRES"):
            os.environ.pop(k_, None)
        os.environ.update(env)
        cols.append(timed([(lambda s=s: sfn(s, None, so)) for s in ss]))
    for k_ in ("TSTEREO_TC2_MT", "TSTEREO_TC2_WRES"):
        os.environ.pop(k_, None)
    t2 = timed([(lambda s=s: sfn(s, out, so)) for s in ss])
    same = torch.equal(f32(xs[0], torch.empty_like(out)), sfn(ss[0], torch.empty_like(out), None)[0])
    mb = (xs[0].numel() + out.numel()) * 4 / 1e6
    print(f"{name:40s} {t0:9.1f} " + " ".join(f"{c:9.1f}" for c in cols) + f" {t2:9.1f}  {mb:8.1f}  {'bit-identical' if same else 'DIFFERENT'}", flush=True)
