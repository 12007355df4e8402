"""CPU probe (test tooling: it patches the oracle): ONLY the UNet decoder's convolutions (fuse.0/1, deconv4, concat, deconv2 —
the mask-logit branch after the top-2 selection, module.py:485-492) run with single-term fp16 operands (no lo half);
everything upstream stays fp32.  Measures the full-resolution EPE this costs, to decide whether the decoder may run
1-term MMAs instead of the 3-term hi+lo split."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.nn.functional as F
from oracle import oracle as O
from temporalstereo_b200 import synth

orig_c2d, orig_ct2 = F.conv2d, F.conv_transpose2d
MODE = ["fp32"]


def r(x):
    if MODE[0] == "fp16":
        return x.half().float()
    if MODE[0] == "bf16":
        return x.bfloat16().float()
    if MODE[0] == "fp16x2_act1":        # weights hi+lo (2 terms), activations single fp16
        return x.half().float()
    return x


def wrap(fn):
    def f(x, w, bias=None, *a, **k):
        if MODE[0] == "fp32" or not DEC[0]:
            return fn(x, w, bias, *a, **k)
        wr = w if MODE[0] == "fp16x2_act1" else r(w)
        y = fn(r(x).double(), wr.double(), None, *a, **k).float()
        if bias is not None:
            y = y + bias.view(1, -1, 1, 1)
        return y
    return f


DEC = [False]
F.conv2d, F.conv_transpose2d = wrap(orig_c2d), wrap(orig_ct2)
orig_dec = O.unet_decoder


def dec(*a, **k):
    DEC[0] = True
    try:
        return orig_dec(*a, **k)
    finally:
        DEC[0] = False


O.unet_decoder = dec
sd = synth.synthetic_state_dict(seed=0)
for H, W in ((320, 576), (96, 160)):
    lf, rf, li, ri = synth.synthetic_frame(H, W, B=1, seed=3)
    with torch.no_grad():
        MODE[0] = "fp32"
        ref = O.aggregation_forward(sd, lf, rf, li, ri, {})
        for m in ("fp16", "fp16x2_act1", "bf16"):
            MODE[0] = m
            out = O.aggregation_forward(sd, lf, rf, li, ri, {})
            d = (out[0][0] - ref[0][0]).abs()
            print(f"{H}x{W} decoder operands {m:12s}: full-res EPE {d.mean():.2e} max {d.max():.2e} (lower levels unchanged: {(out[0][1] - ref[0][1]).abs().max():.1e})", flush=True)
