"""Error statistics of the tcgen05 3xTF32 conv against an fp64 reference, as a function of K."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.nn.functional as F
from temporalstereo_b200 import ops
def rnd(*shape, seed=0, scale=1.0):
    return torch.from_numpy((scale * np.random.RandomState(seed).standard_normal(shape)).astype(np.float32))
def pack(w):
    cout, cin = w.shape[:2]; w = w.reshape(cout, cin, -1); p = torch.zeros(cin, w.shape[2], (cout + 3) // 4 * 4); p[:, :, :cout] = w.permute(1, 2, 0); return p.contiguous()
for Cin in (8, 32, 64, 128, 304, 352):
    Cout, D, H, W = 16, 2, 20, 37
    x = rnd(1, Cin, D, H, W, seed=1); w = rnd(Cout, Cin, 1, 3, 3, seed=2, scale=(2.0 / (9 * Cin)) ** 0.5)
    want = F.conv3d(x.double(), w.double(), None, 1, (0, 1, 1))
    got = ops.conv_hw3_tc(x.cuda(), ops.pack_conv_hw3_tc(w.reshape(Cout, Cin, 9).cuda()), None, Cout, 1, None).cpu().double()
    simt = ops.conv_hw3(x.cuda(), pack(w).cuda(), None, Cout, 1, 1, None).cpu().double()
    for name, g in (("tc", got), ("fp32 fma", simt)):
        e = g - want
        bias = (e * torch.sign(want)).mean().item()
        print(f"Cin={Cin:4d} {name:9s} max|e| {e.abs().max():.2e} rms {e.pow(2).mean().sqrt():.2e} signed-bias(towards |x| growth) {bias:+.2e}  rms(want) {want.pow(2).mean().sqrt():.2f}")
