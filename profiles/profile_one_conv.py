"""One large DIRECT-mode convolution (the UNet encoder's 32 -> 32 at half resolution, both images of 8 frames) between
cudaProfilerStart/Stop, for a source-level ncu capture of conv_tc2_kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from temporalstereo_b200 import ops
x = torch.randn(16, 32, 272, 480, device="cuda")
w = torch.randn(32, 32, 9) * 0.06
ws, inv = ops.fp16_prescale(w)
wp, b, inv = ops.pack_conv_hw3_tc2(ws, True).cuda(), torch.randn(32, device="cuda") * 0.1, inv.cuda()
out = torch.empty(16, 32, 272, 480, device="cuda")
for _ in range(3):
    ops.conv_hw3_tc2(x, wp, b, 32, 1, "ReLU", out=out, half=True, oscale=inv)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
ops.conv_hw3_tc2(x, wp, b, 32, 1, "ReLU", out=out, half=True, oscale=inv)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
