"""Launches one conv variant a few times (for ncu): python scripts/one_conv.py Cin Cout D H W dil [B] [kind]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from temporalstereo_b200 import ops
Cin, Cout, D, H, W, dil = map(int, sys.argv[1:7])
B = int(sys.argv[7]) if len(sys.argv) > 7 else 4
kind = sys.argv[8] if len(sys.argv) > 8 else "tc2"
x = torch.randn(B, Cin, D, H, W, device="cuda")
w = torch.randn(Cout, Cin, 9, device="cuda") * 0.05
bias = torch.randn(Cout, device="cuda")
if kind == "tc2":
    wp = ops.pack_conv_hw3_tc2(w)
    f = lambda: ops.conv_hw3_tc2(x, wp, bias, Cout, dil, "SiLU")
elif kind == "f16":
    wp = ops.pack_conv_hw3_tc2(w, True)
    f = lambda: ops.conv_hw3_tc2(x, wp, bias, Cout, dil, "SiLU", half=True)
else:
    wp = ops.pack_conv_hw3_tc(w)
    f = lambda: ops.conv_hw3_tc(x, wp, bias, Cout, dil, "SiLU")
for _ in range(3):
    f()
torch.cuda.synchronize()
