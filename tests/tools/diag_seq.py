"""Diagnostic: two-frame sequence, engine vs oracle, same-state and carried-state."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import oracle as O
from temporalstereo_b200 import synth, temporal
from temporalstereo_b200.aggregation import TEMPORALSTEREO

H, W, B = [int(x) for x in sys.argv[1:4]] if len(sys.argv) > 3 else (96, 128, 2)
sd = synth.synthetic_state_dict(seed=0)
eng = TEMPORALSTEREO(); eng.load_state_dict(sd, strict=True); eng = eng.cuda().eval()
st = synth.synthetic_temporal_state(H, W, B=B)
pose = (st["K"], st["T_now"], st["inv_T_prev"], st["baseline"])
def cu(x):
    if torch.is_tensor(x):
        return x.cuda()
    if isinstance(x, dict):
        return {k: cu(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [cu(v) for v in x]
    return x

def rep(tag, out, want):
    names = ["disp", "cost", "sample", "off"]
    s = []
    for n, a, b in zip(names, out[:4], want[:4]):
        for i, (x, y) in enumerate(zip(a, b)):
            d = (x.cpu() - y).abs()
            s.append(f"{n}{i}:{d.mean().item():.1e}/{d.max().item():.1e}")
    print(tag, " ".join(s))

lf, rf, li, ri = synth.synthetic_frame(H, W, B=B, seed=20)
with torch.no_grad():
    want0 = O.aggregation_forward(sd, lf, rf, li, ri, {})
out0 = eng(cu(lf), cu(rf), cu(li), cu(ri), {})
rep("frame0", out0, want0)
for k in ("disp_sample", "cost_volume"):
    d = (out0[5]["cost_memory"][k].cpu() - want0[5]["cost_memory"][k]).abs()
    print("  mem", k, d.mean().item(), d.max().item())
state = want0[5]
ref_state = O.update_map({k: (dict(v) if isinstance(v, dict) else v) for k, v in state.items()}, *pose, H, W, True, 3)
dev_state = temporal.update_map(cu({k: (dict(v) if isinstance(v, dict) else v) for k, v in state.items()}), *[p.cuda() for p in pose], H, W, True, 3)
for k in ("disp_sample", "cost_volume"):
    d = (dev_state["cost_memory"][k].cpu() - ref_state["cost_memory"][k]).abs()
    print("  warped mem", k, d.mean().item(), d.max().item(), "ref absmax", ref_state["cost_memory"][k].abs().max().item())
d = (dev_state["local_map"].cpu() - ref_state["local_map"]).abs()
print("  warped local_map", d.mean().item(), d.max().item(), ref_state["local_map"].shape)
lf, rf, li, ri = synth.synthetic_frame(H, W, B=B, seed=21)
with torch.no_grad():
    want1 = O.aggregation_forward(sd, lf, rf, li, ri, ref_state)
out1 = eng(cu(lf), cu(rf), cu(li), cu(ri), dev_state)
rep("frame1 same-state", out1, want1)
# feed the oracle-warped state directly (isolates the aggregation from update_map)
out1b = eng(cu(lf), cu(rf), cu(li), cu(ri), cu({k: (dict(v) if isinstance(v, dict) else v) for k, v in ref_state.items()}))
rep("frame1 oracle-warped-state", out1b, want1)

# carried state: engine and oracle each carry their own frame-0 state
print("---- carried")
lf, rf, li, ri = synth.synthetic_frame(H, W, B=B, seed=20)
ref_prev, dev_prev = {}, {}
with torch.no_grad():
    w0 = O.aggregation_forward(sd, lf, rf, li, ri, ref_prev)
o0 = eng(cu(lf), cu(rf), cu(li), cu(ri), dev_prev)
rep("frame0", o0, w0)
ref_prev, dev_prev = w0[5], o0[5]
print("keys", sorted(ref_prev.keys()), sorted(dev_prev.keys()))
d = (dev_prev["prev_disp"].cpu() - ref_prev["prev_disp"]).abs(); print("prev_disp", d.mean().item(), d.max().item())
ref_prev = O.update_map(ref_prev, *pose, H, W, True, 3)
dev_prev = temporal.update_map(dev_prev, *[p.cuda() for p in pose], H, W, True, 3)
for k in ("disp_sample", "cost_volume"):
    d = (dev_prev["cost_memory"][k].cpu() - ref_prev["cost_memory"][k]).abs()
    print("  warped mem", k, d.mean().item(), d.max().item())
d = (dev_prev["local_map"].cpu() - ref_prev["local_map"]).abs()
print("  warped local_map", d.mean().item(), d.max().item(), ref_prev["local_map"].shape, dev_prev["local_map"].shape)
lf, rf, li, ri = synth.synthetic_frame(H, W, B=B, seed=21)
with torch.no_grad():
    w1 = O.aggregation_forward(sd, lf, rf, li, ri, ref_prev)
o1 = eng(cu(lf), cu(rf), cu(li), cu(ri), dev_prev)
rep("frame1 carried", o1, w1)
