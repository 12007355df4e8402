"""Diagnostic: replay chain A of tests/test_gpu_sequences.py at C4 and dump the pixels whose fine-level costs differ."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.set_num_threads(1)
from oracle import oracle as O
from temporalstereo_b200 import synth, temporal
from temporalstereo_b200.aggregation import TEMPORALSTEREO

H, W, B, ns, T = 480, 640, 2, 20, 5
cu = lambda x: x.cuda() if torch.is_tensor(x) else ({k: cu(v) for k, v in x.items()} if isinstance(x, dict) else ([cu(v) for v in x] if isinstance(x, (list, tuple)) else x))
cp = lambda st: {k: (dict(v) if isinstance(v, dict) else v) for k, v in st.items()}
sd = synth.synthetic_state_dict(seed=0)
eng = TEMPORALSTEREO(coarse=dict(num_sample=ns)); eng.load_state_dict(sd, strict=True); eng = eng.cuda().eval()
st = synth.synthetic_temporal_state(H, W, B=B)
pose = (st["K"], st["T_now"], st["inv_T_prev"], st["baseline"])
ref_state = {}
for t in range(T):
    lf, rf, li, ri = synth.synthetic_frame(H, W, B=B, seed=40 + t)
    if t:
        with torch.no_grad():
            ref_state = O.update_map(ref_state, *pose, H, W, True, 3)
    dev_in = cu(cp(ref_state))
    with torch.no_grad():
        want = O.aggregation_forward(sd, lf, rf, li, ri, cp(ref_state), num_sample=ns)
    out = eng(cu(lf), cu(rf), li.cuda(), ri.cuda(), dev_in)
    dc = (out[1][1].cpu() - want[1][1]).abs()
    print(f"frame {t}: fine max|dcost| {dc.max():.3e}; local_map range", None if "local_map" not in ref_state else (ref_state["local_map"].min().item(), ref_state["local_map"].max().item()),
          "mem sample range", None if ref_state.get("cost_memory") is None else (ref_state["cost_memory"]["disp_sample"].min().item(), ref_state["cost_memory"]["disp_sample"].max().item()))
    if dc.max() > 1e-3:
        am = (dc.amax(1) == dc.amax(1).max()).nonzero()[:1]
        for b, y, x in am.tolist():
            print(f" pixel b={b} y={y} x={x}")
            print("   samples oracle", [f"{v:.9g}" for v in want[2][1][b, :, y, x].tolist()])
            print("   samples engine", [f"{v:.9g}" for v in out[2][1][b, :, y, x].cpu().tolist()])
            print("   cost oracle   ", [f"{v:.5f}" for v in want[1][1][b, :, y, x].tolist()])
            print("   cost engine   ", [f"{v:.5f}" for v in out[1][1][b, :, y, x].cpu().tolist()])
            cm = ref_state["cost_memory"]
            print("   mem_cost@px   ", [f"{v:.6g}" for v in cm["cost_volume"][b, :, y, x].tolist()], " mem_sample@px", [f"{v:.9g}" for v in cm["disp_sample"][b, :, y, x].tolist()],
                  " |mem_cost| max in 5x5:", cm["cost_volume"][b, :, max(y-2,0):y+3, max(x-2,0):x+3].abs().max().item(), " global max", cm["cost_volume"].abs().max().item())
            lm = ref_state["local_map"]
            print("   local_map@px  ", [f"{v:.9g}" for v in lm[b, :, y, x].tolist()], " coarse disp (up)", want[0][3][b, 0, y, x].item())
        # which neighbours have extreme candidates?
        s = want[2][1]
        print("   fine candidates overall range", s.min().item(), s.max().item(), " nonfinite:", (~torch.isfinite(s)).sum().item())
    ref_state = want[5]
