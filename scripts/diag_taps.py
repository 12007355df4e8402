import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch, numpy as np, torch.nn.functional as F
torch.set_num_threads(8)
from oracle import oracle as O
from temporalstereo_b200 import ops
from test_gpu_ops import _op_inputs, rnd
B, C, Cout, S, H, W = 1, 128, 8, 5, 136, 240
L, R, smp = _op_inputs(6, B, C, H, W, S)
smp[:, 0] = torch.round(smp[:, 0]); smp[:, -1] = smp[:, -1] + W
planes = 2 * C + 3 * (C // 8)
w = rnd(Cout, planes, 1, 3, 3, seed=61, scale=(2.0 / (9 * planes)) ** 0.5); b = rnd(Cout, seed=62, scale=0.1)
vol = O.block_cost(L, R, smp, 3)
want = O._act(F.conv3d(vol.double(), w.double(), b.double(), 1, (0, 1, 1)), "SiLU").float()
w9 = w.reshape(Cout, planes, 9)
Lc, Rc, sc = L.cuda(), R.cuda(), smp.cuda()
g = ops.group_cost(Lc, Rc, sc)
addl = ops.conv_hw3_tc2(Lc, ops.pack_conv_hw3_tc2(w9[:, :C].contiguous(), True).cuda(), None, Cout, 1, None, half=True)
wt, osc_t = ops.fp16_prescale(ops.tap_projection_weights(w9[:, C:2 * C].contiguous()))
T = ops.conv_d_tc2(Rc.unsqueeze(2), ops.pack_conv_d_tc2(wt, True).cuda(), None, 9 * Cout, 1, 1, 1, False, None, half=True, oscale=osc_t.cuda()).view(B, 9 * Cout, H, W)
Tref = torch.einsum("ock,chw->kohw", w9[:, C:2*C].double(), R[0].double()).reshape(1, 9*Cout, H, W).float()
print("T err", (T.cpu() - Tref).abs().max().item(), Tref.abs().max().item())
gc = ops.conv_hw3_tc2(g, ops.pack_conv_hw3_tc2(w9[:, 2 * C:].contiguous(), True).cuda(), None, Cout, 1, None, half=True)
got, _ = ops.cost_taps(T, sc, gc, addl, b.cuda(), Cout, "SiLU")
prod = ops.cost_conv_warp(Rc, sc, g, addl, ops.pack_conv_hw3_tc2(w9[:, C:].contiguous(), True).cuda(), b.cuda(), Cout, "SiLU", half=True)
mat = ops.conv_hw3_tc2(ops.block_cost(Lc, Rc, sc), ops.pack_conv_hw3_tc2(w9, True).cuda(), b.cuda(), Cout, 1, "SiLU", half=True)
volg = ops.block_cost(Lc, Rc, sc).cpu()
print("volume gpu vs oracle", (volg - vol).abs().max().item(), "per d:", [(volg[:, :, d] - vol[:, :, d]).abs().max().item() for d in range(S)])
for name, t in (("taps", got), ("producer", prod), ("materialised", mat)):
    d = (t.cpu() - want).abs()
    print(name, "vs oracle max", d.max().item(), "per d:", [d[:, :, k].max().item() for k in range(S)])
print("taps vs producer", (got - prod).abs().max().item())
