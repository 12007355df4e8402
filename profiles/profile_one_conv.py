"""One large DIRECT-mode convolution (the UNet encoder's 32 -> 32 at half resolution, both images of 8 frames) between
cudaProfilerStart/Stop, for a source-level ncu capture of conv_tc2_kernel.  SPLIT=1 (default): the S-format (TMA-fed)
form, S-format in and out; SPLIT=0: fp32 in and out (register producer)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from temporalstereo_b200 import ops
split = os.environ.get("SPLIT", "1") == "1"
x = torch.randn(16, 32, 272, 480, device="cuda")
w = torch.randn(32, 32, 9) * 0.06
ws, inv = ops.fp16_prescale(w)
wp, b, inv = ops.pack_conv_hw3_tc2(ws, True).cuda(), torch.randn(32, device="cuda") * 0.1, inv.cuda()
out = torch.empty(16, 32, 272, 480, device="cuda")
xs = ops.split_pack(x)
so = ops.Split(16, 32, 1, 272, 480, 2, device="cuda", five=False)
def run():
    if split:
        ops.conv_hw3_s(xs, wp, b, 32, 1, "ReLU", half=1, oscale=inv, sout=so)
    else:
        ops.conv_hw3_tc2(x, wp, b, 32, 1, "ReLU", out=out, half=True, oscale=inv)
for _ in range(3):
    run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
