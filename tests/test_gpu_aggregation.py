"""End-to-end parity of the drop-in TEMPORALSTEREO module (CUDA, through the C ABI) against
 (1) golden outputs of the real reference (tests/golden, made by oracle/make_golden.py) and
 (2) the CPU oracle on the same seeded inputs at the BASELINE.json sizes.

Tolerances (north_star): regressed disparity within 1e-3 px EPE of the reference fp32 path;
index work (sorted candidate lists, top-2 selection) exact.
"""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as O
from temporalstereo_b200 import synth

pytestmark = pytest.mark.gpu

EPE_TOL = 1e-3      # px, mean abs difference of the full-resolution disparity (north_star)


@pytest.fixture(scope="module")
def engine():
    assert torch.cuda.is_available()
    from temporalstereo_b200.aggregation import TEMPORALSTEREO
    m = TEMPORALSTEREO()
    m.load_state_dict(synth.synthetic_state_dict(seed=0), strict=True)
    return m.cuda().eval()


def _load(golden_dir, name):
    return {k: torch.from_numpy(v) for k, v in np.load(os.path.join(golden_dir, name)).items()}


def _cuda(x):
    if torch.is_tensor(x):
        return x.cuda()
    if isinstance(x, dict):
        return {k: _cuda(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [_cuda(v) for v in x]
    return x


def _report(out, ref_disps, ref_costs, ref_samples, ref_offs):
    disps, costs, samples, offs = out[:4]
    rep = {}
    for i, (a, b) in enumerate(zip(disps, ref_disps)):
        d = (a.cpu() - b).abs()
        rep[f"disp{i}"] = (d.mean().item(), d.max().item())
    for i, (a, b) in enumerate(zip(costs, ref_costs)):
        rep[f"cost{i}"] = ((a.cpu() - b).abs().max().item(),)
    for i, (a, b) in enumerate(zip(samples, ref_samples)):
        rep[f"sample{i}"] = ((a.cpu() - b).abs().max().item(),)
    for i, (a, b) in enumerate(zip(offs, ref_offs)):
        rep[f"off{i}"] = ((a.cpu() - b).abs().max().item(),)
    return rep


def _check(out, ref, what):
    ref_disps, ref_costs, ref_samples, ref_offs = ref
    rep = _report(out, ref_disps, ref_costs, ref_samples, ref_offs)
    print(what, {k: tuple(f"{x:.2e}" for x in v) for k, v in rep.items()})
    for t in list(out[0]) + list(out[1]) + list(out[2]) + list(out[3]):
        assert torch.isfinite(t).all(), f"{what}: non-finite output"
    # shapes of the 6-tuple
    for a, b in zip(list(out[0]) + list(out[1]) + list(out[2]) + list(out[3]),
                    list(ref_disps) + list(ref_costs) + list(ref_samples) + list(ref_offs)):
        assert tuple(a.shape) == tuple(b.shape), (what, a.shape, b.shape)
    assert rep["disp0"][0] < EPE_TOL, f"{what}: full-res EPE {rep['disp0'][0]:.3e} px >= {EPE_TOL}"
    for i in (1, 2, 3):
        assert rep[f"disp{i}"][0] < EPE_TOL, f"{what}: disp{i} EPE {rep[f'disp{i}'][0]:.3e}"
    # the coarse candidate list is pure index work (sorted integers + zero memory): exact
    assert rep["sample2"][0] < 2e-5, f"{what}: coarse samples differ {rep['sample2']}"      # exact in single-frame mode (asserted there)
    return rep


def test_single_frame_matches_reference_golden(engine, golden_dir):
    g = _load(golden_dir, "agg_single_96x160.npz")
    lf, rf, li, ri = synth.synthetic_frame(96, 160, B=1, seed=1)
    out = engine(_cuda(lf), _cuda(rf), li.cuda(), ri.cuda(), {})
    ref = ([g[f"disp{i}"] for i in range(4)], [g[f"cost{i}"] for i in range(3)],
           [g[f"sample{i}"] for i in range(3)], [g[f"off{i}"] for i in range(3)])
    rep = _check(out, ref, "golden single 96x160")
    assert rep["sample2"][0] == 0.0, "coarse sorted candidates must be bit-exact"
    for i in range(3):
        assert rep[f"cost{i}"][0] < 2e-3, rep
        assert rep[f"off{i}"][0] < 1e-4, rep
    info = out[5]
    d = (info["cost_memory"]["disp_sample"].cpu() - g["mem_sample"]).abs()
    assert d.mean() < EPE_TOL, d.mean()
    d = (info["cost_memory"]["cost_volume"].cpu() - g["mem_cost"]).abs()
    assert d.max() < 2e-3, d.max()
    assert info["prev_disp"].data_ptr() == out[0][0].data_ptr() or torch.equal(info["prev_disp"], out[0][0])


def test_temporal_frame_matches_reference_golden(engine, golden_dir):
    """update_map (CUDA) -> aggregation with cost memory + local map, against the reference's outputs."""
    from temporalstereo_b200 import temporal
    g = _load(golden_dir, "agg_temporal_96x160.npz")
    H, W = 96, 160
    lf, rf, li, ri = synth.synthetic_frame(H, W, B=1, seed=1)
    st = synth.synthetic_temporal_state(H, W, B=1)
    prev = _cuda(dict(prev_disp=st["prev_disp"], cost_memory=st["cost_memory"], local_map=st["local_map"]))
    prev = temporal.update_map(prev, st["K"].cuda(), st["T_now"].cuda(), st["inv_T_prev"].cuda(),
                               st["baseline"].cuda(), H, W, True, 3)
    out = engine(_cuda(lf), _cuda(rf), li.cuda(), ri.cuda(), prev)
    ref = ([g[f"disp{i}"] for i in range(4)], [g[f"cost{i}"] for i in range(3)],
           [g[f"sample{i}"] for i in range(3)], [g[f"off{i}"] for i in range(3)])
    assert out[2][1].shape[1] == 10 and out[2][2].shape[1] == 14      # 3 local-map + 5 + 2 memory; 12 + 2
    _check(out, ref, "golden temporal 96x160")


@pytest.mark.parametrize("ns,H,W", [(16, 96, 288), (20, 96, 352)])
def test_other_disparity_ranges_match_reference_golden(golden_dir, ns, H, W):
    """D = 256 / 320 (16 / 20 coarse candidates, BASELINE configs C5 / C4), temporal mode, B=2, against outputs of the REAL
    reference.  The aggregation starts from the reference's own warped state (stored in the golden), so the discontinuous
    splat normalisation cannot leak into the comparison; the engine's warp is compared with it in the bulk."""
    from temporalstereo_b200 import temporal
    from temporalstereo_b200.aggregation import TEMPORALSTEREO
    g = _load(golden_dir, f"agg_temporal_ns{ns}_{H}x{W}.npz")
    eng = TEMPORALSTEREO(coarse=dict(num_sample=ns))
    eng.load_state_dict(synth.synthetic_state_dict(seed=0), strict=True)
    eng = eng.cuda().eval()
    lf, rf, li, ri = synth.synthetic_frame(H, W, B=2, seed=12)
    st = synth.synthetic_temporal_state(H, W, B=2)
    own = temporal.update_map(_cuda(dict(prev_disp=st["prev_disp"], cost_memory=dict(st["cost_memory"]), local_map=st["local_map"])),
                              st["K"].cuda(), st["T_now"].cuda(), st["inv_T_prev"].cuda(), st["baseline"].cuda(), H, W, True, 3)
    for got, key in ((own["cost_memory"]["disp_sample"], "warp_mem_sample"), (own["cost_memory"]["cost_volume"], "warp_mem_cost"),
                     (own["local_map"], "warp_local_map")):
        d = (got.cpu() - g[key]).abs()
        assert d.median() < 1e-5 and (d > 1e-3).float().mean() < 0.01, (key, d.median().item())
    prev = _cuda(dict(prev_disp=st["prev_disp"], use_past_cost=True, local_map_size=3, local_map=g["warp_local_map"],
                      cost_memory=dict(disp_sample=g["warp_mem_sample"], cost_volume=g["warp_mem_cost"])))
    out = eng(_cuda(lf), _cuda(rf), li.cuda(), ri.cuda(), prev)
    ref = ([g[f"disp{i}"] for i in range(4)], [g[f"cost{i}"] for i in range(3)],
           [g[f"sample{i}"] for i in range(3)], [g[f"off{i}"] for i in range(3)])
    assert out[2][2].shape[1] == ns + 2 and out[2][1].shape[1] == 10
    _check(out, ref, f"golden temporal D={16 * ns} {H}x{W} B=2")


@pytest.mark.parametrize("H,W,B", [(320, 576, 1), (544, 960, 1), (96, 112, 3)])
def test_single_frame_vs_oracle(engine, H, W, B):
    """BASELINE configs C1 (320x576) and C2 (540x960 -> 544x960), plus a batched ragged case."""
    sd = synth.synthetic_state_dict(seed=0)
    lf, rf, li, ri = synth.synthetic_frame(H, W, B=B, seed=3)
    with torch.no_grad():
        want = O.aggregation_forward(sd, lf, rf, li, ri, {})
    out = engine(_cuda(lf), _cuda(rf), li.cuda(), ri.cuda(), {})
    rep = _check(out, want[:4], f"oracle single {H}x{W} B={B}")
    assert rep["sample2"][0] == 0.0, "single-frame coarse candidates are pure index work: exact"
    # top-2 index work: the stored memory is top-2 (sample+offset)/2 resized; compare as EPE
    d = (out[5]["cost_memory"]["disp_sample"].cpu() - want[5]["cost_memory"]["disp_sample"]).abs().mean()
    assert d < EPE_TOL, d


def test_sequence_vs_oracle(engine):
    """BASELINE config C3 shape (KITTI 384x1248, T=2, pose warp on): two frames carrying prev_info."""
    from temporalstereo_b200 import temporal
    H, W = 384, 1248
    sd = synth.synthetic_state_dict(seed=0)
    st = synth.synthetic_temporal_state(H, W, B=1)
    pose = (st["K"], st["T_now"], st["inv_T_prev"], st["baseline"])
    # frame 0
    lf, rf, li, ri = synth.synthetic_frame(H, W, B=1, seed=10)
    with torch.no_grad():
        want0 = O.aggregation_forward(sd, lf, rf, li, ri, {})
    out0 = engine(_cuda(lf), _cuda(rf), li.cuda(), ri.cuda(), {})
    _check(out0, want0[:4], "oracle sequence frame 0")
    carried = {k: (dict(v) if isinstance(v, dict) else v) for k, v in out0[5].items()}

    # frame 1, both sides starting from the SAME recurrent state (the oracle's frame-0 state): the
    # top-2 selection is discontinuous, so a handful of frame-0 pixels whose two best costs are
    # within rounding of each other would otherwise dominate the comparison (SURVEY.md §7 hard part 3)
    lf, rf, li, ri = synth.synthetic_frame(H, W, B=1, seed=11)
    state = want0[5]
    dev_state = temporal.update_map(_cuda({k: (dict(v) if isinstance(v, dict) else v) for k, v in state.items()}),
                                    *[p.cuda() for p in pose], H, W, True, 3)
    with torch.no_grad():
        ref_state = O.update_map(dict(state), *pose, H, W, True, 3)
        want1 = O.aggregation_forward(sd, lf, rf, li, ri, ref_state)
    out1 = engine(_cuda(lf), _cuda(rf), li.cuda(), ri.cuda(), dev_state)
    assert out1[2][1].shape[1] == 8, "fine level must see 1 local-map (first warp) + 5 range + 2 memory candidates"
    _check(out1, want1[:4], "oracle sequence frame 1 (same state)")

    # frame 1 again with the engine's OWN carried state: the bulk of the image must still agree
    dev_state = temporal.update_map(carried, *[p.cuda() for p in pose], H, W, True, 3)
    out1c = engine(_cuda(lf), _cuda(rf), li.cuda(), ri.cuda(), dev_state)
    d = (out1c[0][0].cpu() - want1[0][0]).abs()
    frac = (d > 1e-2).float().mean().item()
    print(f"carried state: median |d| {d.median().item():.2e} px, mean {d.mean().item():.2e}, >0.01px: {100 * frac:.2f} %")
    assert d.median() < 1e-3 and frac < 0.2, (d.median().item(), frac)


def test_tf32_split_engine_vs_oracle():
    """The tf32 hi+lo operand split (half_split = False, no fp16 range limit) through the whole engine."""
    from temporalstereo_b200.aggregation import TEMPORALSTEREO
    sd = synth.synthetic_state_dict(seed=0)
    eng = TEMPORALSTEREO()
    eng.half_split = False
    eng.load_state_dict(sd, strict=True)
    eng = eng.cuda().eval()
    lf, rf, li, ri = synth.synthetic_frame(96, 160, B=2, seed=4)
    with torch.no_grad():
        want = O.aggregation_forward(sd, lf, rf, li, ri, {})
    out = eng(_cuda(lf), _cuda(rf), li.cuda(), ri.cuda(), {})
    _check(out, want[:4], "oracle single 96x160 B=2, tf32 split")


@pytest.mark.parametrize("num_sample,H,W,B", [(20, 96, 352, 2), (16, 96, 288, 1)])
def test_other_disparity_ranges_temporal(num_sample, H, W, B):
    """BASELINE configs C4 (D=320 -> 20 coarse candidates) and C5 (D=256 -> 16) at reduced resolution, temporal
    mode on (cost memory + local map: 8 fine candidates), batch > 1.  The width keeps W/16 > num_sample: narrower
    images leave the far candidates without any right-image support, their costs tie, and the top-2 selection
    (discontinuous) flips on rounding noise — in the oracle as much as in the engine."""
    from temporalstereo_b200 import temporal
    from temporalstereo_b200.aggregation import TEMPORALSTEREO
    sd = synth.synthetic_state_dict(seed=0)
    eng = TEMPORALSTEREO(coarse=dict(num_sample=num_sample))
    eng.load_state_dict(sd, strict=True)
    eng = eng.cuda().eval()
    st = synth.synthetic_temporal_state(H, W, B=B)
    pose = (st["K"], st["T_now"], st["inv_T_prev"], st["baseline"])
    state = dict(prev_disp=st["prev_disp"], cost_memory=dict(st["cost_memory"]), local_map=st["local_map"])
    with torch.no_grad():
        ref_state = O.update_map({k: (dict(v) if isinstance(v, dict) else v) for k, v in state.items()}, *pose, H, W, True, 3)
    dev_state = temporal.update_map(_cuda(state), *[p.cuda() for p in pose], H, W, True, 3)
    lf, rf, li, ri = synth.synthetic_frame(H, W, B=B, seed=12)
    # both sides aggregate from the oracle's warped state (the splat normalisation is discontinuous); the forward
    # updates prev_info in place, so each side gets its own copy
    copy = lambda st_: {k: (dict(v) if isinstance(v, dict) else v) for k, v in st_.items()}
    dev_in = _cuda(copy(ref_state))
    with torch.no_grad():
        want = O.aggregation_forward(sd, lf, rf, li, ri, copy(ref_state), num_sample=num_sample)
    out = eng(_cuda(lf), _cuda(rf), li.cuda(), ri.cuda(), dev_in)
    assert out[2][2].shape[1] == num_sample + 2 and out[2][1].shape[1] == want[2][1].shape[1]
    _check(out, want[:4], f"oracle D={16 * num_sample} temporal {H}x{W} B={B}")
    # and the CUDA warp itself against the oracle's: the splat's x / (norm + 1e-22) is discontinuous where almost
    # nothing lands, so compare the bulk (the B=1 golden test pins the per-pixel values)
    for k in ("disp_sample", "cost_volume"):
        d = (dev_state["cost_memory"][k].cpu() - ref_state["cost_memory"][k]).abs()
        assert d.median() < 1e-5 and (d > 1e-3).float().mean() < 0.01, (k, d.median(), (d > 1e-3).float().mean())


@pytest.mark.parametrize("plan", ["tc2", "simt"])
def test_idempotent_and_batch_independent(engine, plan):
    """Size-independent properties: same input -> bit-identical output (no atomics on the aggregation
    path); a batch of 2 equals the two frames run separately."""
    lf, rf, li, ri = synth.synthetic_frame(128, 192, B=2, seed=5)
    engine.plan_mode = plan          # fixed kernel choice: the auto plan may pick different kernels per batch size
    a = engine(_cuda(lf), _cuda(rf), li.cuda(), ri.cuda(), {})
    b = engine(_cuda(lf), _cuda(rf), li.cuda(), ri.cuda(), {})
    for x, y in zip(a[0] + a[1], b[0] + b[1]):
        assert torch.equal(x, y)
    for i in range(2):
        one = engine([t[i:i + 1].cuda() for t in lf], [t[i:i + 1].cuda() for t in rf], li[i:i + 1].cuda(),
                     ri[i:i + 1].cuda(), {})
        for x, y in zip(a[0] + a[1], one[0] + one[1]):
            assert torch.equal(x[i:i + 1], y)
    engine.plan_mode = "auto"


def test_materialised_cost_path_agrees_with_the_fused_one(engine):
    """fuse_cost=False runs ops.block_cost (the drop-in operator, raw volume in HBM) + the plain first conv at every level;
    True never writes a raw volume; the default fuses the two warp levels.  Same math, other summation order: all within
    tolerance of the oracle and of each other."""
    sd = synth.synthetic_state_dict(seed=0)
    lf, rf, li, ri = synth.synthetic_frame(128, 192, B=2, seed=6)
    with torch.no_grad():
        want = O.aggregation_forward(sd, lf, rf, li, ri, {})
    default = engine.fuse_cost
    assert set(default) == {"fine", "precise"}
    outs = {}
    try:
        for mode in (default, True, False):
            engine.fuse_cost = mode
            outs[str(mode)] = engine(_cuda(lf), _cuda(rf), li.cuda(), ri.cuda(), {})
            _check(outs[str(mode)], want[:4], f"fuse_cost={mode}")
    finally:
        engine.fuse_cost = default
    assert (outs["True"][0][0] - outs["False"][0][0]).abs().mean() < 1e-4


def test_decoder_single_term_precision(engine):
    """The UNet decoder runs one fp16 MMA term (engine.decoder_single_term): measured on the oracle at 7.6e-5 px EPE
    (tests/tools/precision_probe_decoder.py).  On the GPU: both settings within 1e-3 px of the oracle, every output other
    than the full-resolution disparity bit-identical between them."""
    sd = synth.synthetic_state_dict(seed=0)
    lf, rf, li, ri = synth.synthetic_frame(128, 192, B=2, seed=9)
    with torch.no_grad():
        want = O.aggregation_forward(sd, lf, rf, li, ri, {})
    assert engine.decoder_single_term
    fast = engine(_cuda(lf), _cuda(rf), li.cuda(), ri.cuda(), {})
    engine.decoder_single_term = False
    try:
        exact = engine(_cuda(lf), _cuda(rf), li.cuda(), ri.cuda(), {})
    finally:
        engine.decoder_single_term = True
    _check(fast, want[:4], "decoder single-term")
    _check(exact, want[:4], "decoder 3-term")
    for a, b in zip(fast[0][1:] + fast[1] + fast[2] + fast[3], exact[0][1:] + exact[1] + exact[2] + exact[3]):
        assert torch.equal(a, b)
    d = (fast[0][0] - exact[0][0]).abs()
    e3 = (exact[0][0].cpu() - want[0][0]).abs().mean().item()
    print(f"decoder single-term vs 3-term: EPE {d.mean().item():.2e} max {d.max().item():.2e}; 3-term vs oracle {e3:.2e}")
    assert d.mean() < 5e-4


def test_streams_and_forms_agree(engine):
    """Engine switches that must not change results: the side streams (UNet encoder / decoder, the independent branches of a
    level) only reorder launches -> bit-identical outputs, also under repetition (a stream race would show up as a
    difference); the S-format (`split_format`) and the two forms of the warp levels' first conv (`cost_form`) change the
    summation order only -> every variant within the oracle tolerances."""
    sd = synth.synthetic_state_dict(seed=0)
    lf, rf, li, ri = synth.synthetic_frame(128, 192, B=2, seed=11)
    with torch.no_grad():
        want = O.aggregation_forward(sd, lf, rf, li, ri, {})
    args = (_cuda(lf), _cuda(rf), li.cuda(), ri.cuda())
    base = engine(*args, {})
    _check(base, want[:4], "default")
    assert engine.overlap_encoder and engine.overlap_branches and engine.split_format and engine.cost_form == "taps"
    try:
        for enc, br in ((False, False), (True, False), (False, True), (True, True), (True, True)):
            engine.overlap_encoder, engine.overlap_branches = enc, br
            out = engine(*args, {})
            for x, y in zip(base[0] + base[1] + base[2] + base[3], out[0] + out[1] + out[2] + out[3]):
                assert torch.equal(x, y), (enc, br)
        engine.cost_form = "producer"
        _check(engine(*args, {}), want[:4], "cost_form=producer")
        engine.cost_form = "taps"
        engine.split_format = False
        _check(engine(*args, {}), want[:4], "split_format=False")
    finally:
        engine.overlap_encoder = engine.overlap_branches = engine.split_format = True
        engine.cost_form = "taps"
