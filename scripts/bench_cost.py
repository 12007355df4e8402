"""Times the cost-volume path of each level (group terms, left-half conv, fused cost -> first conv; the materialising
operator + conv beside it) at the C2 B=8 shapes, CUDA-graph replays, for each TSTEREO_TC2_MT setting."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from temporalstereo_b200 import ops, synth
from temporalstereo_b200.aggregation import TEMPORALSTEREO

B, H, W = int(os.environ.get("BATCH", "8")), 544, 960
dev = "cuda"
eng = TEMPORALSTEREO()
eng.load_state_dict(synth.synthetic_state_dict(seed=0), strict=True)
eng = eng.cuda().eval()
lf, rf, li, ri = synth.synthetic_frame(H, W, B=B, seed=1)
eng([t.cuda() for t in lf], [t.cuda() for t in rf], li.cuda(), ri.cuda(), {})


def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for lvl, C, cout, S, s in (("precise", 128, 8, 5, 4), ("fine", 128, 16, 5, 8)):
    h, w = H // s, W // s
    L, R = torch.randn(B, C, h, w, device=dev), torch.randn(B, C, h, w, device=dev)
    yy, xx = torch.meshgrid(torch.arange(h, device=dev), torch.arange(w, device=dev), indexing="ij")
    base = 0.06 * w * (1.2 + torch.sin(xx / w * 6.0) * torch.cos(yy / h * 4.0)) + 0.3 * torch.rand(h, w, device=dev)
    smp = (base[None, None] + torch.tensor([-4.0, -1.0, 0.0, 1.0, 4.0], device=dev).view(1, 5, 1, 1)).expand(B, 5, h, w).contiguous()
    k = eng._pk[f"{lvl}.init3d.0.conv.0"]
    g = ops.group_cost(L, R, smp)
    al = ops.conv_hw3_tc2(L, k.tc["left"], None, cout, 1, None, half=True, oscale=k.osc)
    t_g = timed(lambda: ops.group_cost(L, R, smp))
    t_l = timed(lambda: ops.conv_hw3_tc2(L, k.tc["left"], None, cout, 1, None, half=True, oscale=k.osc))
    row = []
    for mt in ("", "2", "4"):
        os.environ.pop("TSTEREO_TC2_MT", None)
        if mt:
            os.environ["TSTEREO_TC2_MT"] = mt
        try:
            row.append(timed(lambda: ops.cost_conv_warp(R, smp, g, al, k.tc["cost"], k.b, cout, "SiLU", half=True, oscale=k.osc)))
        except Exception as e:
            row.append(float("nan"))
    os.environ.pop("TSTEREO_TC2_MT", None)
    print(f"{lvl:8s} group_cost {t_g:7.1f}  left conv {t_l:7.1f}  fused conv default/mt2/mt4 " + " / ".join(f"{t:7.1f}" for t in row), flush=True)
