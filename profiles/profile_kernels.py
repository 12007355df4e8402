"""The kernels the roofline / traffic numbers are quoted on, at the C2 B=8 shapes, between cudaProfilerStart/Stop:

    ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/r02_kernels \
        python profiles/profile_kernels.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from temporalstereo_b200 import ops, synth, temporal  # noqa: E402
from temporalstereo_b200.aggregation import TEMPORALSTEREO  # noqa: E402

B, H, W = int(os.environ.get("BATCH", "8")), 544, 960
dev = "cuda"
eng = TEMPORALSTEREO()
eng.load_state_dict(synth.synthetic_state_dict(seed=0), strict=True)
eng = eng.cuda().eval()
lf, rf, li, ri = synth.synthetic_frame(H, W, B=B, seed=1)
eng([t.cuda() for t in lf], [t.cuda() for t in rf], li.cuda(), ri.cuda(), {})          # packs the weights
hs = eng.half_split
h4, w4 = H // 4, W // 4
L, R = torch.randn(B, 128, h4, w4, device=dev), torch.randn(B, 128, h4, w4, device=dev)
yy, xx = torch.meshgrid(torch.arange(h4, device=dev), torch.arange(w4, device=dev), indexing="ij")
base = 0.06 * w4 * (1.2 + torch.sin(xx / w4 * 6.0) * torch.cos(yy / h4 * 4.0)) + 0.3 * torch.rand(h4, w4, device=dev)
smp = (base[None, None] + torch.tensor([-4.0, -1.0, 0.0, 1.0, 4.0], device=dev).view(1, 5, 1, 1)).expand(B, 5, h4, w4).contiguous()
first = eng._pk["precise.init3d.0.conv.0"]
feat_p = torch.randn(B, 16, 5, h4, w4, device=dev)
w_p = torch.randn(2, 8, 9, device=dev) * 0.1
a6 = torch.randn(B, 8, 6, h4, w4, device=dev)
sk6 = torch.randn(B, 8, 5, h4, w4, device=dev)
lg = torch.randn(B, 9, H, W, device=dev)
dp = torch.rand(B, 1, h4, w4, device=dev) * 40
x2 = torch.randn(B, 64, H // 2, W // 2, device=dev)
cc = eng._pk["precise.refinement.concat"]
vol = torch.randn(B, 16, 7, 68, 120, device=dev)
av, mx = torch.empty_like(vol), torch.empty_like(vol)
st = synth.synthetic_temporal_state(384, 1248, B=1)
state = {"prev_disp": st["prev_disp"].cuda(), "cost_memory": {k: v.cuda() for k, v in st["cost_memory"].items()}, "local_map": st["local_map"].cuda()}
pose = [st[k].cuda() for k in ("K", "T_now", "inv_T_prev", "baseline")]


sl, sr = ops.split_pack(L), ops.split_pack(R)
sr5 = ops.Split(B, 128, 1, h4, w4, 2, t=sr.t, five=True)
x2s = ops.split_pack(torch.randn(2 * B, 32, H // 2, W // 2, device=dev))                   # UNet conv2.1 input (both images)
x2o = ops.Split(2 * B, 32, 1, H // 2, W // 2, 2, device=dev, five=False)
c21 = eng._pk["precise.refinement.conv2.1"]
xcs = ops.split_pack(x2, 1)                                                               # decoder concat input, hi half only
xco = ops.Split(B, 32, 1, H // 2, W // 2, 1, device=dev, five=False)


def run():
    # ---- the engine's first-conv path at the precise level (eng._first_conv, "taps" form)
    g = ops.group_cost(L, R, smp)                                                         # block_cost_main (group terms) + resize
    ops.split_pack(L, out=sl)
    ops.split_pack(R, out=sr)
    al, _ = ops.conv_hw3_s(sl, first.tc["left"], None, 8, 1, None, half=1, oscale=first.osc)          # left half, once per frame
    T, _ = ops.conv_d_s(sr5, first.tc["taps"], None, 72, 1, 1, 1, False, None, half=1, oscale=first.tc["taps_osc"])   # 3 launches
    gc = ops.conv_hw3_tc2(g, first.tc["gconv"], None, 8, 1, None, half=True, oscale=first.osc)
    ops.cost_taps(T.view(B, 72, h4, w4), smp, gc, al, first.b, 8, "SiLU", sout=ops.Split(B, 8, 5, h4, w4, 2, device=dev))
    # ---- the producer form it replaced
    alp = ops.conv_hw3_tc2(L, first.tc["left"], None, 8, 1, None, half=hs, oscale=first.osc)
    ops.cost_conv_warp(R, smp, g, alp, first.tc["cost"], first.b, 8, "SiLU", half=hs, oscale=first.osc)
    # ---- the materialising operator + conv over the volume
    v = ops.block_cost(L, R, smp)
    ops.conv_hw3_tc2(v, first.tc["hw3"], first.b, 8, 1, "SiLU", half=hs, oscale=first.osc)
    del v
    # ---- S-format (TMA-fed) convolutions: UNet 32 -> 32 at 1/2 scale (3-term, S in / S out), decoder concat 64 -> 32 (1-term)
    ops.conv_hw3_s(x2s, c21.tc["hw3"], c21.b, 32, 1, "ReLU", half=1, oscale=c21.osc, sout=x2o)
    ops.conv_hw3_s(xcs, cc.tc["hw3"], cc.b, 32, 1, "ReLU", half=2, oscale=cc.osc, sout=xco)
    ops.conv_hw3_tc2(x2, cc.tc["hw3"], cc.b, 32, 1, "ReLU", half=2, oscale=cc.osc)        # the same layer, fp32 in / out (register producer)
    # ---- streaming kernels
    ops.heads(feat_p, w_p, 1.0)
    ops.resize_add_act_s(a6, (5, h4, w4), sk6, "SiLU")
    ops.unet_upsample(lg, dp)
    ops.pool5(vol, av, mx)
    temporal.update_map({k: (dict(v) if isinstance(v, dict) else v) for k, v in state.items()}, *pose, 384, 1248, True, 3)


run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
