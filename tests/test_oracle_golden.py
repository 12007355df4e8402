"""Pins the CPU oracle (oracle/oracle.py) to outputs of the REAL reference code
(tests/golden/*.npz, produced by oracle/make_golden.py in the build container)."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as O
from temporalstereo_b200 import synth

# 1 thread: the oracle is then bit-identical to the reference outputs in the build container
# (multi-threaded mkldnn convs alone move disparities by ~1e-4 px)
torch.set_num_threads(1)


def _load(golden_dir, name):
    return {k: torch.from_numpy(v) for k, v in np.load(os.path.join(golden_dir, name)).items()}


def _op_inputs(seed, B, C, H, W, S):
    rng = np.random.RandomState(seed)
    L = torch.from_numpy(rng.standard_normal((B, C, H, W)).astype(np.float32))
    R = torch.from_numpy(rng.standard_normal((B, C, H, W)).astype(np.float32))
    smp = torch.from_numpy(rng.uniform(-3.0, W / 2.0, (B, S, H, W)).astype(np.float32))
    return L, R, smp


@pytest.mark.parametrize("tag", ["a", "b"])
def test_block_cost_matches_reference(golden_dir, tag):
    g = _load(golden_dir, f"block_cost_warp_{tag}.npz")
    B, C, H, W, S = [int(v) for v in g["shape"]]
    L, R, smp = _op_inputs(10, B, C, H, W, S)
    torch.testing.assert_close(O.block_cost(L, R, smp, 3), g["out"], rtol=0, atol=1e-6)
    g = _load(golden_dir, f"block_cost_shift_{tag}.npz")
    torch.testing.assert_close(O.block_cost(L, R, S, 3), g["out"], rtol=0, atol=1e-6)


def test_project_to_3d_matches_reference(golden_dir):
    g = _load(golden_dir, "project_to_3d.npz")
    st = synth.synthetic_temporal_state(64, 96, B=2)
    K8 = st["K"].clone(); K8[:, :2] /= 8.0
    T = torch.bmm(st["T_now"], st["inv_T_prev"])
    depth = 0.54 * K8[:, 0, 0].view(-1, 1, 1, 1) / (st["cost_memory"]["disp_sample"] + 1e-5)
    flow, tri = O.project_to_3d(depth, K8, torch.inverse(K8), T)
    torch.testing.assert_close(flow, g["flow"], rtol=1e-5, atol=1e-4)
    torch.testing.assert_close(tri, g["tri"], rtol=1e-6, atol=1e-6)


def _check_agg(out, g, atol_disp=2e-4, atol_cost=2e-4):
    disps, costs, samples, offs, ranges, info = out
    for i, t in enumerate(disps):
        torch.testing.assert_close(t, g[f"disp{i}"], rtol=0, atol=atol_disp, msg=lambda m, i=i: f"disp{i}: {m}")
    for i, t in enumerate(costs):
        torch.testing.assert_close(t, g[f"cost{i}"], rtol=0, atol=atol_cost, msg=lambda m, i=i: f"cost{i}: {m}")
    for i, t in enumerate(samples):
        torch.testing.assert_close(t, g[f"sample{i}"], rtol=0, atol=atol_disp)
    for i, t in enumerate(offs):
        torch.testing.assert_close(t, g[f"off{i}"], rtol=0, atol=1e-5)


def test_aggregation_single_frame_matches_reference(golden_dir):
    g = _load(golden_dir, "agg_single_96x160.npz")
    sd = synth.synthetic_state_dict(seed=0)
    lf, rf, li, ri = synth.synthetic_frame(96, 160, B=1, seed=1)
    with torch.no_grad():
        out = O.aggregation_forward(sd, lf, rf, li, ri, {})
    _check_agg(out, g)
    torch.testing.assert_close(out[5]["cost_memory"]["disp_sample"], g["mem_sample"], rtol=0, atol=2e-4)
    torch.testing.assert_close(out[5]["cost_memory"]["cost_volume"], g["mem_cost"], rtol=0, atol=2e-4)
    epe = (out[0][0] - g["disp0"]).abs().mean().item()
    assert epe < 1e-5, epe


def test_temporal_matches_reference(golden_dir):
    gm = _load(golden_dir, "update_map_96x160.npz")
    g = _load(golden_dir, "agg_temporal_96x160.npz")
    H, W = 96, 160
    sd = synth.synthetic_state_dict(seed=0)
    lf, rf, li, ri = synth.synthetic_frame(H, W, B=1, seed=1)
    st = synth.synthetic_temporal_state(H, W, B=1)
    prev = dict(prev_disp=st["prev_disp"], cost_memory=st["cost_memory"], local_map=st["local_map"])
    with torch.no_grad():
        prev = O.update_map(prev, st["K"], st["T_now"], st["inv_T_prev"], st["baseline"], H, W, True, 3)
        torch.testing.assert_close(prev["cost_memory"]["disp_sample"], gm["mem_sample"], rtol=1e-5, atol=1e-4)
        torch.testing.assert_close(prev["cost_memory"]["cost_volume"], gm["mem_cost"], rtol=1e-5, atol=1e-4)
        torch.testing.assert_close(prev["local_map"], gm["local_map"], rtol=1e-5, atol=1e-4)
        out = O.aggregation_forward(sd, lf, rf, li, ri, prev)
    _check_agg(out, g)


@pytest.mark.parametrize("ns,H,W", [(16, 96, 288), (20, 96, 352)])
def test_other_disparity_ranges_match_reference(golden_dir, ns, H, W):
    """D = 256 / 320 (16 / 20 coarse candidates: BASELINE configs C5 / C4), temporal mode, batch 2."""
    g = _load(golden_dir, f"agg_temporal_ns{ns}_{H}x{W}.npz")
    sd = synth.synthetic_state_dict(seed=0)
    lf, rf, li, ri = synth.synthetic_frame(H, W, B=2, seed=12)
    st = synth.synthetic_temporal_state(H, W, B=2)
    prev = dict(prev_disp=st["prev_disp"], cost_memory=dict(st["cost_memory"]), local_map=st["local_map"])
    with torch.no_grad():
        prev = O.update_map(prev, st["K"], st["T_now"], st["inv_T_prev"], st["baseline"], H, W, True, 3)
        torch.testing.assert_close(prev["cost_memory"]["disp_sample"], g["warp_mem_sample"], rtol=1e-5, atol=1e-4)
        torch.testing.assert_close(prev["cost_memory"]["cost_volume"], g["warp_mem_cost"], rtol=1e-5, atol=1e-4)
        torch.testing.assert_close(prev["local_map"], g["warp_local_map"], rtol=1e-5, atol=1e-4)
        out = O.aggregation_forward(sd, lf, rf, li, ri, prev, num_sample=ns)
    _check_agg(out, g)
    assert out[2][2].shape[1] == ns + 2


def test_losses_match_reference_golden(golden_dir):
    """f2 (forward): the oracle's restatement of the two loss terms against values computed by the reference's own classes
    (oracle/make_golden.py, inputs regenerated from the same seeds)."""
    from oracle.make_golden import loss_inputs
    g = np.load(os.path.join(golden_dir, "losses_96x160.npz"))
    for tag, sparse in (("dense", False), ("sparse", True)):
        est, costs, offs, smps, gt = loss_inputs(sparse)
        for i, e in enumerate(est):
            got = O.smooth_l1_loss_level(e, gt, 192, 0, sparse)
            np.testing.assert_allclose(got.numpy(), g[f"{tag}_l1_loss_lvl{i}"], rtol=1e-6, atol=0)
        for i, (c, o, s) in enumerate(zip(costs, offs, smps)):
            got = O.wasserstein_loss_level(c, o, s, gt, 192, 0, sparse)
            np.testing.assert_allclose(got.numpy(), g[f"{tag}_wars_loss_lvl{i}"], rtol=1e-6, atol=0)
        assert float(O.smooth_l1_loss_level(est[1], torch.zeros_like(gt), 192, 0, sparse)) == float(g[f"{tag}_l1_none"]) == 0.0
