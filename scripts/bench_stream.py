"""Times the streaming (non-convolution) kernels of the frame at the C2 B=8 shapes, each as a CUDA-graph of its launches
(device time, no Python launch overhead), with achieved GB/s on the bytes each must move."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from temporalstereo_b200 import ops
from temporalstereo_b200.graph import CapturedStep

B = int(os.environ.get("BATCH", "8"))
dev = "cuda"


def timed(fn, reps=20):
    step = CapturedStep(lambda: [fn() for _ in range(4)])
    for _ in range(2):
        step.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        step.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps / 4 * 1e3


def report(name, us, nbytes):
    print(f"{name:58s} {us:8.1f} us  {nbytes / 1e6:8.1f} MB  {nbytes / us / 1e3:7.0f} GB/s  ({100 * nbytes / us / 1e3 / 6526.8:4.1f} % of 6526.8)", flush=True)


for lvl, C, D, h, w in (("precise", 8, 5, 136, 240), ("fine", 16, 7, 68, 120), ("coarse", 32, 14, 34, 60)):
    feat = torch.randn(B, 2 * C, D, h, w, device=dev)
    wt = torch.randn(2, C, 9, device=dev) * 0.1
    report(f"heads {lvl} [{2 * C}ch x {D} x {h}x{w}]", timed(lambda: ops.heads(feat, wt, 1.0)), 4 * B * (2 * C + 2) * D * h * w)
    cost, smp, off = (torch.randn(B, D, h, w, device=dev) for _ in range(3))
    report(f"predict_disp {lvl}", timed(lambda: ops.predict_disp(cost, smp, off, True)), 4 * B * (3 * D + 5) * h * w)
for name, C, src, dst in (("resize_add_act precise.conv6", 8, (6, 136, 240), (5, 136, 240)), ("resize_add_act precise.conv5", 16, (4, 68, 120), (3, 68, 120)),
                          ("resize_add_act fine.conv6", 16, (6, 68, 120), (5, 68, 120)), ("resize_add_act coarse.conv6", 32, (12, 34, 60), (12, 34, 60)),
                          ("resize_add_act coarse.conv5", 64, (6, 18, 30), (6, 17, 30))):
    a = torch.randn(B, C, *src, device=dev)
    sk = torch.randn(B, C, *dst, device=dev)
    n = B * C * (src[0] * src[1] * src[2] + 2 * dst[0] * dst[1] * dst[2]) * 4
    report(name, timed(lambda: ops.resize_add_act(a, dst, sk, "SiLU")), n)
lg = torch.randn(B, 9, 544, 960, device=dev)
dp = torch.rand(B, 1, 136, 240, device=dev) * 40
report("unet_upsample 544x960", timed(lambda: ops.unet_upsample(lg, dp)), 4 * B * (10 * 544 * 960 + 136 * 240))
for lvl, C, D, h, w in (("fine", 16, 7, 68, 120), ("coarse", 32, 14, 34, 60)):
    x = torch.randn(B, C, D, h, w, device=dev)
    av, mx = torch.empty_like(x), torch.empty_like(x)
    report(f"pool5 {lvl}", timed(lambda: ops.pool5(x, av, mx)), 4 * B * C * D * h * w * 3)
for lvl, C, S, h, w in (("precise", 128, 5, 136, 240), ("fine", 128, 5, 68, 120), ("coarse", 256, 12, 34, 60)):
    L, R = torch.randn(B, C, h, w, device=dev), torch.randn(B, C, h, w, device=dev)
    if lvl == "coarse":
        smp = S
    else:
        yy, xx = torch.meshgrid(torch.arange(h, device=dev), torch.arange(w, device=dev), indexing="ij")
        base = 0.06 * w * (1.2 + torch.sin(xx / w * 6.0) * torch.cos(yy / h * 4.0))
        smp = (base[None, None] + torch.tensor([-4.0, -1.0, 0.0, 1.0, 4.0], device=dev).view(1, 5, 1, 1)).expand(B, 5, h, w).contiguous()
    nS = S
    report(f"group_cost {lvl}", timed(lambda: ops.group_cost(L, R, smp)), 4 * B * (2 * C * h * w + (nS * h * w if lvl != "coarse" else 0) + 3 * (C // 8) * nS * h * w))
    planes = (2 * C if lvl != "coarse" else C) + 3 * (C // 8)
    report(f"block_cost {lvl} (materialised)", timed(lambda: ops.block_cost(L, R, smp)), 4 * B * (2 * C * h * w + planes * nS * h * w))
