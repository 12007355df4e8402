"""Round-2 GPU probe: (1) the timed plan (tensor-core vs fp32-FMA kernel per layer shape) at C2, B = 1 and 8 — the data
behind TEMPORALSTEREO._rule; (2) the fused cost -> first-conv path against the materialised one, per level; (3) whole-frame
time with / without fusion.  Writes gpurun_out/probe_r02.md."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from temporalstereo_b200 import ops, synth
from temporalstereo_b200.aggregation import TEMPORALSTEREO

H, W = 544, 960
out_lines = []


def say(s):
    print(s, flush=True)
    out_lines.append(s)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def engine(**kw):
    m = TEMPORALSTEREO()
    m.load_state_dict(synth.synthetic_state_dict(seed=0), strict=True)
    m = m.cuda().eval()
    for k, v in kw.items():
        setattr(m, k, v)
    return m


for B in (8, 1):
    lf, rf, li, ri = synth.synthetic_frame(H, W, B=B, seed=1)
    inp = [t.cuda() for t in lf + rf + [li, ri]]
    fw = lambda m: m(inp[0:3], inp[3:6], inp[6], inp[7], {})
    m = engine(plan_mode="timed")
    fw(m)
    say(f"\n## timed plan, C2 {H}x{W}, B={B}\n\n| kind | input shape | cout | extra | tc2 us | simt us | pick | rule |\n|---|---|---|---|---|---|---|---|")
    for key, t in m._plan_times.items():
        rule = TEMPORALSTEREO._rule(key[0], key[1], key[2])
        pick = min(t, key=t.get)
        say(f"| {key[0]} | {key[1]} | {key[2]} | {key[3:]} | {t.get('tc2', float('nan')):.1f} | {t.get('simt', float('nan')):.1f} | {pick} | {rule}{'' if rule == pick else '  <-- differs'} |")
    say(f"\n## whole frame, B={B} (us per step, 20 steps)\n")
    for name, kw in (("timed plan, fused cost", dict(plan_mode="timed")), ("rule plan, fused cost", dict()),
                     ("rule plan, materialised cost", dict(fuse_cost=False)), ("all tc2, fused", dict(plan_mode="tc2")),
                     ("all simt (convs on fp32 FMA), fused off", dict(plan_mode="simt", fuse_cost=False))):
        mm = engine(**kw)
        fw(mm)
        t_eager = timed(lambda: fw(mm))
        from temporalstereo_b200.graph import CapturedStep
        step = CapturedStep(lambda: fw(mm))
        say(f"* {name}: eager {t_eager:.0f} us, CUDA graph {timed(step.replay):.0f} us ({step.launches} kernels)")
        del step
    del m

# ---- fused vs materialised first conv per level, B = 8
B = 8
m = engine()
lf, rf, li, ri = synth.synthetic_frame(H, W, B=B, seed=1)
inp = [t.cuda() for t in lf + rf + [li, ri]]
m(inp[0:3], inp[3:6], inp[6], inp[7], {})
say(f"\n## cost volume -> first conv per level, B={B} (us)\n\n| level | block_cost | conv over volume | sum | group_cost | left conv | cost_conv | sum fused |\n|---|---|---|---|---|---|---|---|")
h = m.half_split
for lvl, C, cout, S, sc in (("precise", 128, 8, 5, 4), ("fine", 128, 16, 5, 8), ("coarse", 256, 32, 12, 16)):
    hh, ww = H // sc, W // sc
    L = torch.randn(B, C, hh, ww, device="cuda")
    R = torch.randn(B, C, hh, ww, device="cuda")
    a = m._pk[f"{lvl}.init3d.0.conv.0"]
    if lvl == "coarse":
        smp = S
    else:
        yy, xx = torch.meshgrid(torch.arange(hh, device="cuda"), torch.arange(ww, device="cuda"), indexing="ij")
        base = 0.06 * ww * (1.2 + torch.sin(xx / ww * 6.0) * torch.cos(yy / hh * 4.0))
        offs = torch.tensor([-4.0, -1.0, 0.0, 1.0, 4.0], device="cuda")
        smp = (base[None, None] + offs.view(1, 5, 1, 1)).expand(B, 5, hh, ww).contiguous()
    vol = ops.block_cost(L, R, smp)
    t_bc = timed(lambda: ops.block_cost(L, R, smp))
    t_cv = timed(lambda: ops.conv_hw3_tc2(vol, a.tc["hw3"], a.b, cout, 1, "SiLU", half=h))
    g = ops.group_cost(L, R, smp)
    t_g = timed(lambda: ops.group_cost(L, R, smp))
    if lvl == "coarse":
        t_l = 0.0
        t_f = timed(lambda: ops.cost_conv_shift(L, R, g, a.tc["cost"], a.b, cout, "SiLU", half=h))
    else:
        addl = ops.conv_hw3_tc2(L, a.tc["left"], None, cout, 1, None, half=h)
        t_l = timed(lambda: ops.conv_hw3_tc2(L, a.tc["left"], None, cout, 1, None, half=h))
        t_f = timed(lambda: ops.cost_conv_warp(R, smp, g, addl, a.tc["cost"], a.b, cout, "SiLU", half=h))
        for mt in ("2", "4"):
            os.environ["TSTEREO_TC2_MT"] = mt
            say(f"  ({lvl} cost_conv with MT={mt}: {timed(lambda: ops.cost_conv_warp(R, smp, g, addl, a.tc['cost'], a.b, cout, 'SiLU', half=h)):.1f} us)")
        os.environ.pop("TSTEREO_TC2_MT")
    say(f"| {lvl} | {t_bc:.1f} | {t_cv:.1f} | {t_bc + t_cv:.1f} | {t_g:.1f} | {t_l:.1f} | {t_f:.1f} | {t_g + t_l + t_f:.1f} |")
    del vol, g

os.makedirs("gpurun_out", exist_ok=True)
open("gpurun_out/probe_r02.md", "w").write("\n".join(out_lines) + "\n")
