"""Generate tests/golden/*.npz by running the REAL reference code (imported from
/root/reference through oracle/refstubs) on the seeded synthetic inputs of
temporalstereo_b200.synth.  Run in the build container only:

    python oracle/make_golden.py

Inputs are not stored (they are re-derived from the seeds); only reference outputs are.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_import  # noqa: E402
from oracle import oracle as O  # noqa: E402
from temporalstereo_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
torch.set_num_threads(1)            # the reference is thread-count sensitive


def save(name, **arrs):
    np.savez_compressed(os.path.join(OUT, name), **{k: np.asarray(v, dtype=np.float32) for k, v in arrs.items()})
    print("wrote", name, {k: tuple(np.shape(v)) for k, v in arrs.items()})


def op_inputs(seed, B, C, H, W, S):
    rng = np.random.RandomState(seed)
    L = torch.from_numpy(rng.standard_normal((B, C, H, W)).astype(np.float32))
    R = torch.from_numpy(rng.standard_normal((B, C, H, W)).astype(np.float32))
    smp = torch.from_numpy(rng.uniform(-3.0, W / 2.0, (B, S, H, W)).astype(np.float32))
    return L, R, smp


def main():
    os.makedirs(OUT, exist_ok=True)
    ref_import.setup()
    from architecture.modeling.aggregation.utils.block_cost import block_cost as ref_block_cost
    from architecture.modeling.layers.inverse_warp import project_to_3d as ref_project

    # ---- a1/a2/a3: block_cost, both branches, odd sizes exercise floor pooling ----
    with torch.no_grad():
        for tag, (B, C, H, W, S) in {"a": (2, 16, 10, 14, 5), "b": (1, 8, 9, 13, 8)}.items():
            L, R, smp = op_inputs(10, B, C, H, W, S)
            save(f"block_cost_warp_{tag}.npz", out=ref_block_cost(L, R, smp, 3).numpy(), shape=[B, C, H, W, S])
            save(f"block_cost_shift_{tag}.npz", out=ref_block_cost(L, R, S, 3).numpy(), shape=[B, C, H, W, S])

        # ---- a18: project_to_3d ----
        st = synth.synthetic_temporal_state(64, 96, B=2)
        K8 = st["K"].clone(); K8[:, :2] /= 8.0
        T = torch.bmm(st["T_now"], st["inv_T_prev"])
        depth = 0.54 * K8[:, 0, 0].view(-1, 1, 1, 1) / (st["cost_memory"]["disp_sample"] + 1e-5)
        o = ref_project(depth, K8, torch.inverse(K8), T)
        save("project_to_3d.npz", flow=o["optical_flow"].numpy(), tri=o["triangular_depth"].numpy())

        # ---- a15: whole aggregation, single frame ----
        sd = synth.synthetic_state_dict(seed=0)
        agg = ref_import.build_reference_aggregation()
        agg.load_state_dict(sd, strict=True)
        H, W = 96, 160
        lf, rf, li, ri = synth.synthetic_frame(H, W, B=1, seed=1)
        disps, costs, samples, offs, ranges, info = agg(lf, rf, li, ri, prev_info={})
        arrs = {}
        for i, t in enumerate(disps): arrs[f"disp{i}"] = t.numpy()
        for i, t in enumerate(costs): arrs[f"cost{i}"] = t.numpy()
        for i, t in enumerate(samples): arrs[f"sample{i}"] = t.numpy()
        for i, t in enumerate(offs): arrs[f"off{i}"] = t.numpy()
        arrs["mem_sample"] = info["cost_memory"]["disp_sample"].numpy()
        arrs["mem_cost"] = info["cost_memory"]["cost_volume"].numpy()
        save("agg_single_96x160.npz", **arrs)

        # ---- a17 + a15: temporal (update_map through the reference with the CPU splat
        #      restatement, then aggregation with memory + local map) ----
        tm = ref_import.build_reference_temporal(O.softsplat_softmax)
        st = synth.synthetic_temporal_state(H, W, B=1)
        prev = dict(prev_disp=st["prev_disp"], cost_memory=st["cost_memory"], local_map=st["local_map"])
        batch = {"baseline": st["baseline"], ("color_aug", 0, "l"): li, ("K", 0): st["K"],
                 ("inv_T", -1, "l"): st["inv_T_prev"], ("T", 0, "l"): st["T_now"]}
        _, prev = tm.update_map(batch, prev, 0)
        save("update_map_96x160.npz", mem_sample=prev["cost_memory"]["disp_sample"].numpy(),
             mem_cost=prev["cost_memory"]["cost_volume"].numpy(), local_map=prev["local_map"].numpy())
        disps, costs, samples, offs, ranges, info = agg(lf, rf, li, ri, prev_info=prev)
        arrs = {}
        for i, t in enumerate(disps): arrs[f"disp{i}"] = t.numpy()
        for i, t in enumerate(costs): arrs[f"cost{i}"] = t.numpy()
        for i, t in enumerate(samples): arrs[f"sample{i}"] = t.numpy()
        for i, t in enumerate(offs): arrs[f"off{i}"] = t.numpy()
        save("agg_temporal_96x160.npz", **arrs)

        # ---- other disparity ranges: D = 256 / 320 -> 16 / 20 coarse candidates (BASELINE configs C5 / C4), temporal
        #      mode, batch 2.  The width keeps W/16 > num_sample (narrower images leave the far candidates without any
        #      right-image support and their costs tie).
        from temporalstereo_b200.synth import DEFAULT_LEVELS
        for ns, (H, W) in {16: (96, 288), 20: (96, 352)}.items():
            lv = {k: dict(v) for k, v in DEFAULT_LEVELS.items()}
            lv["coarse"]["num_sample"] = ns
            agg = ref_import.build_reference_aggregation(lv)
            agg.load_state_dict(sd, strict=True)
            lf, rf, li, ri = synth.synthetic_frame(H, W, B=2, seed=12)
            st = synth.synthetic_temporal_state(H, W, B=2)
            prev = dict(prev_disp=st["prev_disp"], cost_memory=st["cost_memory"], local_map=st["local_map"])
            batch = {"baseline": st["baseline"], ("color_aug", 0, "l"): li, ("K", 0): st["K"],
                     ("inv_T", -1, "l"): st["inv_T_prev"], ("T", 0, "l"): st["T_now"]}
            _, prev = tm.update_map(batch, prev, 0)
            arrs = {"warp_mem_sample": prev["cost_memory"]["disp_sample"].numpy(),
                    "warp_mem_cost": prev["cost_memory"]["cost_volume"].numpy(), "warp_local_map": prev["local_map"].numpy()}
            disps, costs, samples, offs, ranges, info = agg(lf, rf, li, ri, prev_info=prev)
            for i, t in enumerate(disps): arrs[f"disp{i}"] = t.numpy()
            for i, t in enumerate(costs): arrs[f"cost{i}"] = t.numpy()
            for i, t in enumerate(samples): arrs[f"sample{i}"] = t.numpy()
            for i, t in enumerate(offs): arrs[f"off{i}"] = t.numpy()
            save(f"agg_temporal_ns{ns}_{H}x{W}.npz", **arrs)

        # ---- f2: the two loss terms, forward, from the reference's own classes (dense and sparse ground truth, the three
        #      level resolutions of a 96x160 frame, a case without any valid pixel) ----
        from architecture.modeling.losses.smooth_l1_loss import DispSmoothL1Loss
        from architecture.modeling.losses.warsserstein_distance_loss import WarssersteinDistanceLoss
        arrs = {}
        for tag, sparse in (("dense", False), ("sparse", True)):
            est, costs, offs, smps, gt = loss_inputs(sparse)
            l1 = DispSmoothL1Loss(max_disp=192, start_disp=0, global_weight=1.0, weights=None, sparse=sparse)(est, gt)
            wa = WarssersteinDistanceLoss(max_disp=192, start_disp=0, global_weight=1.0, weights=None, sparse=sparse)(costs, offs, smps, gt)
            for k, v in {**l1, **wa}.items():
                arrs[f"{tag}_{k}"] = v.numpy()
            none = torch.zeros_like(gt)                       # no valid pixel: both fall back to a masked mean (= 0)
            arrs[f"{tag}_l1_none"] = DispSmoothL1Loss(max_disp=192, sparse=sparse)(est[1], none)["l1_loss_lvl0"].numpy()
        save("losses_96x160.npz", **arrs)


def loss_inputs(sparse: bool, H: int = 96, W: int = 160, B: int = 2, seed: int = 30):
    """Seeded inputs of the loss goldens (regenerated by the tests): disparities / costs / offsets / samples at full, 1/4,
    1/8 and 1/16 resolution and a ground truth with out-of-range and (sparse: 70 % zero) invalid pixels."""
    rng = np.random.RandomState(seed)
    f32 = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float32))
    gt = rng.uniform(-5.0, 230.0, (B, 1, H, W))
    if sparse:
        gt = gt * (rng.uniform(0, 1, gt.shape) > 0.7)
    gt = f32(gt)
    est = [f32(rng.uniform(0.0, 200.0 / s, (B, 1, H // s, W // s))) for s in (1, 4, 4, 8)]
    costs, offs, smps = [], [], []
    for s, D in ((4, 5), (8, 10), (16, 14)):
        shp = (B, D, H // s, W // s)
        costs.append(f32(rng.standard_normal(shp) * 2.0))
        offs.append(f32(rng.uniform(-1.0, 1.0, shp)))
        smps.append(f32(np.sort(rng.uniform(0.0, 192.0 / s, shp), axis=1)))
    return est, costs, offs, smps, gt


if __name__ == "__main__":
    main()
