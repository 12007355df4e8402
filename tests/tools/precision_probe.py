"""CPU probe: how many operand mantissa bits do the conv contractions need for EPE < 1e-3 px?
Rounds conv inputs and weights to `bits` explicit mantissa bits (fp32 accumulate) inside the oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.nn.functional as F
from oracle import oracle as O
from temporalstereo_b200 import synth

def rnd(x, bits):
    if bits >= 23: return x
    drop = 23 - bits
    i = x.contiguous().view(torch.int32)
    half = 1 << (drop - 1)
    i = (i + half) & ~((1 << drop) - 1)
    return i.view(torch.float32)

orig = dict(conv3d=F.conv3d, conv2d=F.conv2d, ct3=F.conv_transpose3d, ct2=F.conv_transpose2d)
BITS = [23]
def wrap(fn):
    def f(x, w, *a, **k):
        return fn(rnd(x, BITS[0]), rnd(w, BITS[0]), *a, **k)
    return f
F.conv3d, F.conv2d, F.conv_transpose3d, F.conv_transpose2d = [wrap(orig[k]) for k in ("conv3d", "conv2d", "ct3", "ct2")]

H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (320, 576)
sd = synth.synthetic_state_dict(seed=0)
lf, rf, li, ri = synth.synthetic_frame(H, W, B=1, seed=3)
with torch.no_grad():
    ref = O.aggregation_forward(sd, lf, rf, li, ri, {})
    for bits in (10, 7, 13, 16, 19, 21):
        BITS[0] = bits
        out = O.aggregation_forward(sd, lf, rf, li, ri, {})
        ds = [(a - b).abs() for a, b in zip(out[0], ref[0])]
        print(f"bits={bits:2d} " + " ".join(f"disp{i}: mean {d.mean():.2e} max {d.max():.2e} >0.5px {100*(d>0.5).float().mean():.3f}%" for i, d in enumerate(ds)), flush=True)
