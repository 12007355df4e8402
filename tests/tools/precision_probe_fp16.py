"""CPU probe (test tooling: it patches the oracle): every conv operand rounded to fp16 hi + fp16 lo, products exact."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.nn.functional as F
from oracle import oracle as O
from temporalstereo_b200 import synth
# operand = fp16 hi + fp16 lo (both through real fp16 incl. subnormals / range), products hi*hi + hi*lo + lo*hi exact in fp32
def split16(x, scale=1.0):
    xs = x * scale
    hi = xs.half().float()
    lo = (xs - hi).half().float()
    return hi, lo
orig = dict(conv3d=F.conv3d, conv2d=F.conv2d, ct3=F.conv_transpose3d, ct2=F.conv_transpose2d)
MODE = ["fp32"]
def wrap(fn):
    def f(x, w, bias=None, *a, **k):
        if MODE[0] == "fp32":
            return fn(x, w, bias, *a, **k)
        ws = 2.0 ** MODE[1]
        xh, xl = split16(x)
        wh, wl = split16(w, ws)
        y = fn(xh.double(), wh.double(), None, *a, **k) + fn(xh.double(), wl.double(), None, *a, **k) + fn(xl.double(), wh.double(), None, *a, **k)
        y = (y / ws).float()
        if bias is not None:
            y = y + bias.view(1, -1, *([1] * (y.dim() - 2)))
        return y
    return f
F.conv3d, F.conv2d, F.conv_transpose3d, F.conv_transpose2d = [wrap(orig[k]) for k in ("conv3d", "conv2d", "ct3", "ct2")]
H, W = 320, 576
sd = synth.synthetic_state_dict(seed=0)
lf, rf, li, ri = synth.synthetic_frame(H, W, B=1, seed=3)
with torch.no_grad():
    ref = O.aggregation_forward(sd, lf, rf, li, ri, {})
    for wscale in (0, 6, 10):
        MODE[:] = ["fp16x2", wscale]
        out = O.aggregation_forward(sd, lf, rf, li, ri, {})
        ds = [(a - b).abs() for a, b in zip(out[0], ref[0])]
        print(f"fp16 hi+lo, weight scale 2^{wscale}: " + " ".join(f"disp{i}: mean {d.mean():.2e} max {d.max():.2e}" for i, d in enumerate(ds)), flush=True)
