"""Stage-by-stage comparison of the fine level at the failing frame of chain A (C4, frame 4)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
torch.set_num_threads(1)
from oracle import oracle as O
from temporalstereo_b200 import synth, ops
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
    if t == T - 1:
        break
    with torch.no_grad():
        want = O.aggregation_forward(sd, lf, rf, li, ri, cp(ref_state), num_sample=ns)
    ref_state = want[5]

def rep(name, got, want):
    d = (got.cpu() - want).abs()
    idx = (d == d.max()).nonzero()[0].tolist()
    print(f"{name:28s} max|d| {d.max().item():.3e} at {idx}  (|ref| max {want.abs().max().item():.3e}; #>1e-3: {(d > 1e-3).sum().item()})", flush=True)

with torch.no_grad():
    l4, l8, l16 = lf; r4, r8, r16 = rf
    d_c, c_c, o_c, s_c = O.coarse_level(l16, r16, sd, ref_state, ns)
    low, high = d_c - 4.0, d_c + 4.0
    samples = O.range_samples(low, high)
    lm = ref_state["local_map"]
    Hh, Ww = l8.shape[-2:]
    lmr = F.interpolate(lm * Ww / lm.shape[-1], size=(Hh, Ww), mode="bilinear", align_corners=True)
    samples = torch.cat([lmr, samples], 1)
    raw = O.block_cost(l8, r8, samples, 3)
    vol = O.init3d(raw, sd, "fine.init3d")
    volm, smp_sorted = O.merge_memory(vol, samples, sd, "fine", ref_state, 2, coarse=False)
    fused = O.pyramid_fusion(volm, sd, "fine.fuse")
    cost, off = O.prediction_heads(fused, sd, "fine.pred_heads", 1.0)
dl8, dr8, dsmp = l8.cuda(), r8.cuda(), samples.cuda().contiguous()
eng._pk = eng._pk or eng._pack(dl8.device)
graw = ops.block_cost(dl8, dr8, dsmp)
rep("block_cost (materialised)", graw, raw)
C = 128
rep("  L half", graw[:, :C], raw[:, :C]); rep("  R warp half", graw[:, C:2 * C], raw[:, C:2 * C]); rep("  group terms", graw[:, 2 * C:], raw[:, 2 * C:])
for mode in (("fine", "precise"), False):
    eng.fuse_cost = mode
    gvol = eng._init3d(dl8, dr8, dsmp, "fine.init3d")
    rep(f"init3d (fuse={bool(mode)})", gvol, vol)
# merge / fuse / heads on the ORACLE's init3d volume (isolates each stage)
dvol = vol.cuda().contiguous()
cm = ref_state["cost_memory"]
pc = eng._pk["fine.past_conv"]
gvolm, gsmp = ops.merge_memory(dvol, dsmp, cm["disp_sample"].cuda().contiguous(), cm["cost_volume"].cuda().contiguous(), pc.w, pc.b, 2)
rep("merge: sorted samples", gsmp, smp_sorted)
rep("merge: gathered volume", gvolm, volm)
d = (gvolm.cpu() - volm).abs().amax(1)          # [B, D, H, W]
bad = (d > 1e-3).nonzero()
print("pixels with a differing gathered plane:", bad.shape[0])
for b, dd, y, x in bad[:8].tolist():
    print(f"  b={b} plane={dd} y={y} x={x}: unsorted candidates", [f"{v:.7g}" for v in samples[b, :, y, x].tolist()],
          "memory", [f"{v:.7g}" for v in cm["disp_sample"][b, :, y, x].tolist()],
          "| sorted oracle", [f"{v:.7g}" for v in smp_sorted[b, :, y, x].tolist()])

# ---- pyramid fusion and heads, each fed the ORACLE's input
Cc = 16
Bq, _, Dq, Hq, Wq = volm.shape
cat = torch.empty((Bq, 4 * Cc, Dq, Hq, Wq), device="cuda")
cat[:, :Cc].copy_(volm.cuda())
c5 = eng._pk["fine.fuse.conv_5x5"]
with torch.no_grad():
    o5 = O.conv3d_bn_act(volm, sd, "fine.fuse.conv_5x5", padding=(2, 0, 0), act="SiLU") if True else None
eng._d(cat[:, :Cc], c5, 5, 1, 1, False, "SiLU", out=cat[:, Cc:2 * Cc])
try:
    rep("fuse.conv_5x5", cat[:, Cc:2 * Cc], o5)
except Exception as e:  # noqa
    print("conv_5x5 compare skipped:", e)
ops.pool5(cat[:, :Cc], cat[:, 2 * Cc:3 * Cc], cat[:, 3 * Cc:])
rep("pool5 avg", cat[:, 2 * Cc:3 * Cc], F.avg_pool3d(volm, 5, 1, 2))
rep("pool5 max", cat[:, 3 * Cc:], F.max_pool3d(volm, 5, 1, 2))
gf = eng._sep(cat, "fine.fuse.conv_fuse", act0=None, act1=None)
rep("pyramid_fusion out", gf, fused)
st_, fin = eng._pk["fine.pred_heads.stem"], eng._pk["fine.pred_heads.final"]
feat = eng._d(fused.cuda().contiguous(), st_, 3, 1, 1, False, "SiLU")
gc, go = ops.heads(feat, fin.w, 1.0)
rep("heads cost (oracle input)", gc, cost)
rep("heads off  (oracle input)", go, off)
for mode in ("simt", "tc2"):
    eng.plan_mode = mode
    feat = eng._d(fused.cuda().contiguous(), st_, 3, 1, 1, False, "SiLU")
    gc, go = ops.heads(feat, fin.w, 1.0)
    rep(f"heads cost, stem plan={mode}", gc, cost)
    gf = eng._sep(cat, "fine.fuse.conv_fuse", act0=None, act1=None)
    rep(f"pyramid_fusion, plan={mode}", gf, fused)
