"""BASELINE.json configs at their STATED shapes, and multi-frame sequences with carried state.

  C3  KITTI 384x1248, D=192, T=2, pose warp on          (test_gpu_aggregation.py::test_sequence_vs_oracle)
  C4  TartanAir 480x640, D=320 (20 coarse candidates), T=5 sequence, B>=2
  C5  1088x1920 (1080x1920 padded), D=256 (16 candidates), temporal frame, B=2

Index work is checked as index work: the coarse candidate list must be exact, and the top-2 selection of every level
must pick the same two candidates as the oracle wherever the oracle's decision is not within rounding of a tie (margin
between the 2nd and 3rd best cost > MARGIN); regressed disparity within 1e-3 px EPE (north_star).  Pixels whose candidate
SORT is itself decided by < 1e-4 px between a memory plane and a regular plane (see `sort_tie_masks`) are excluded with
their receptive field and their share of the image is reported (a fraction of a per cent).  In the end-to-end runs the
same holds for a top-2 selection that flips INSIDE the oracle's own tie margin at the coarse / fine level: it is counted
as a raw (undecided) mismatch at its level, and the finer levels — whose candidates it moves — are not compared around it.
"""
import pytest
import torch
import torch.nn.functional as F

from oracle import oracle as O
from temporalstereo_b200 import synth

pytestmark = pytest.mark.gpu

EPE_TOL = 1e-3
MARGIN = 1e-4          # cost margin below which a top-2 decision counts as a tie: a selection can only flip if the 2nd / 3rd
                       # best costs move against each other by more than their margin, i.e. 2 x the largest cost difference —
                       # 5e-5 on the oracle's candidates (strict checks), up to 8.5e-5 end to end (MARGIN_E2E)
MARGIN_E2E = 2e-4


def _cuda(x):
    if torch.is_tensor(x):
        return x.cuda()
    if isinstance(x, dict):
        return {k: _cuda(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [_cuda(v) for v in x]
    return x


def _copy(st):
    return {k: (dict(v) if isinstance(v, dict) else v) for k, v in st.items()}


def _engine(num_sample):
    from temporalstereo_b200.aggregation import TEMPORALSTEREO
    eng = TEMPORALSTEREO(coarse=dict(num_sample=num_sample))
    eng.load_state_dict(synth.synthetic_state_dict(seed=0), strict=True)
    return eng.cuda().eval()


def top2_mismatches(cost, ref_cost, keep, margin=MARGIN):
    """Pixels (inside `keep`) whose top-2 candidate SET differs from the oracle's, counted only where the oracle's 2nd and
    3rd best costs are more than `margin` apart (otherwise the choice is a rounding-level tie in the reference itself)."""
    ref_top = torch.topk(ref_cost, k=min(3, ref_cost.shape[1]), dim=1)
    got_top = torch.topk(cost.cpu(), k=2, dim=1)
    a = torch.sort(ref_top.indices[:, :2], dim=1).values
    b = torch.sort(got_top.indices, dim=1).values
    differs = (a != b).any(1) & keep
    if ref_cost.shape[1] > 2:
        decided = (ref_top.values[:, 1] - ref_top.values[:, 2]) > margin
    else:
        decided = torch.ones_like(differs)
    return int((differs & decided).sum()), int(differs.sum()), int(keep.sum())


def top2_tie_flips(cost, ref_cost, margin=MARGIN):
    """Pixels whose top-2 candidate set differs from the oracle's WHERE the oracle's own 2nd / 3rd best costs are within
    `margin` (a rounding-level tie of the reference): the level's disparity jumps there, and so does everything the next
    level computes around it.  Same class of discontinuity as a sort tie; such pixels are counted as `raw` mismatches at
    their own level and their receptive field is excluded at the finer levels."""
    if ref_cost.shape[1] <= 2:
        return torch.zeros(ref_cost.shape[0], *ref_cost.shape[-2:], dtype=torch.bool)
    ref_top = torch.topk(ref_cost, k=3, dim=1)
    got_top = torch.topk(cost.cpu(), k=2, dim=1)
    a = torch.sort(ref_top.indices[:, :2], dim=1).values
    b = torch.sort(got_top.indices, dim=1).values
    return (a != b).any(1) & ((ref_top.values[:, 1] - ref_top.values[:, 2]) <= margin)


def _dilate(m, r):
    return F.max_pool2d(m.float().unsqueeze(1), 2 * r + 1, 1, r).squeeze(1) > 0


def sort_tie_masks(ref_samples, state, tol=1e-4, flip_c=None, flip_f=None):
    """Where the reference's candidate SORT is decided by less than `tol` between a memory plane and a regular plane.

    merge_memory concatenates the level's candidates with the two memory samples and sorts them (coarse.py:100-104,
    fine.py:118-122); the volume planes follow the permutation.  A memory plane (past_conv of a stored cost) and a regular
    plane have unrelated contents, so when a memory sample lands within rounding of another candidate (the engine's
    up-sampled disparities differ from the oracle's by up to ~3e-5 px) the order of two DIFFERENT planes hangs on the last
    bits of the previous level's output — a discontinuity of the reference algorithm (the fp64 oracle flips there against
    the fp32 one as well), not something a kernel can match.  Returns keep masks (precise, fine, coarse, full) that
    exclude those pixels and the receptive field they feed (pool5 + two 3x3 convs + convex up-sampling: radius 6 at the
    level, carried through each x2 up-sampling; the precise hourglass adds its own).  `flip_c` / `flip_f`: pixels of the
    coarse / fine level whose top-2 selection flipped inside the reference's own tie margin (`top2_tie_flips`): their
    disparity is discontinuous too, so the finer levels are not compared around them."""
    s_p, s_f, s_c = ref_samples
    B, _, Hc, Wc = s_c.shape
    memory = state.get("cost_memory") if state.get("use_past_cost", False) else None

    def ties(sorted_s, ms):
        """ms [B,2,h,w]: a memory sample within tol of a candidate that is not (one of) the memory samples themselves"""
        t = torch.zeros(sorted_s.shape[0], *sorted_s.shape[-2:], dtype=torch.bool)
        for k in range(ms.shape[1]):
            v = ms[:, k:k + 1]
            near_all = ((sorted_s - v).abs() < tol).sum(1)
            near_mem = ((ms - v).abs() < tol).sum(1)
            t |= near_all > near_mem
        return t
    if memory is None:
        ms_f = torch.zeros(B, 2, *s_f.shape[-2:])
        tie_c = torch.zeros(B, Hc, Wc, dtype=torch.bool)          # zeros against the integer candidate 0: exact on both sides
    else:
        ms_f = memory["disp_sample"]
        ms_c = F.interpolate(ms_f * Wc / ms_f.shape[-1], size=(Hc, Wc), mode="bilinear", align_corners=True)
        tie_c = ties(s_c, ms_c)
    tie_c = _dilate(tie_c, 6)
    up = lambda m: F.interpolate(m.float().unsqueeze(1), scale_factor=2, mode="nearest").squeeze(1) > 0
    tie_f = _dilate(ties(s_f, ms_f) | _dilate(up(tie_c), 2), 6)
    # whatever moves the coarse DISPARITY — a flipped selection, or a coarse sort tie (its cost differences can flip the
    # selection anywhere in its neighbourhood) — moves the fine level's CANDIDATES (3x3 convex up-sampling around it): the
    # difference enters the fine cost volume itself and spreads through the whole fine init3d (hourglass + dilated convs)
    moved = tie_c if flip_c is None else (tie_c | flip_c)
    if moved.any():
        tie_f = tie_f | _dilate(_dilate(up(moved), 2), 46)
    src_f = tie_f if flip_f is None else (tie_f | flip_f)
    tie_p = _dilate(up(src_f), 40)            # + the precise hourglass (two stride-2 stages) and the dilated convs
    tie_full = _dilate(F.interpolate(tie_p.float().unsqueeze(1), scale_factor=4, mode="nearest").squeeze(1) > 0, 4)
    return ~tie_p, ~tie_f, ~tie_c, ~tie_full


def check_frame(out, want, what, state, strict=False):
    """`state`: the recurrent state the frame was computed from.  strict: nothing excluded (used when the levels run on
    the oracle's candidates: the sorts cannot flip)."""
    disps, costs, samples, offs = out[:4]
    rd, rc, rs, ro = want[:4]
    has_memory = state.get("cost_memory") is not None and state.get("use_past_cost", False)
    flip_c = flip_f = None
    if not strict:      # end to end: a rounding-level top-2 tie at one level moves the candidates of the next
        flip_c, flip_f = top2_tie_flips(costs[2], rc[2], MARGIN_E2E), top2_tie_flips(costs[1], rc[1], MARGIN_E2E)
    keep_p, keep_f, keep_c, keep_full = sort_tie_masks(rs, state, tol=-1.0 if strict else 1e-4, flip_c=flip_c, flip_f=flip_f)
    keeps = [keep_p, keep_f, keep_c]
    dkeep = [keep_full, keep_p, keep_p, keep_f]          # disps: full, precise (1/4), fine up-sampled (1/4), coarse up-sampled (1/8)
    if not strict:
        # a tie-margin flip of the precise level's own selection moves that pixel's precise disparity and the 3x3 (x4)
        # neighbourhood of the full-resolution one; with ~1 % of the image left after the sort-tie exclusion a single such
        # pixel would carry the mean
        flip_p = top2_tie_flips(costs[0], rc[0], MARGIN_E2E)
        if flip_p.any():
            dkeep[1] = keep_p & ~flip_p
            dkeep[0] = keep_full & ~_dilate(F.interpolate(_dilate(flip_p, 1).float().unsqueeze(1), scale_factor=4, mode="nearest").squeeze(1) > 0, 4)
    excluded = 1.0 - keep_full.float().mean().item()
    epes = [((a.cpu() - b).abs()[:, 0][k]).mean().item() if k.any() else 0.0 for a, b, k in zip(disps, rd, dkeep)]
    rep, bad_total = [], 0
    for i, lvl in enumerate(("precise", "fine", "coarse")):
        bad, raw, n = top2_mismatches(costs[i], rc[i], keeps[i], MARGIN if strict else MARGIN_E2E)
        bad_total += bad
        dc = (costs[i].cpu() - rc[i]).abs().amax(1)[keeps[i]].max().item() if keeps[i].any() else 0.0
        rep.append(f"{lvl} {bad}/{raw}/{n} (max |dcost| {dc:.1e})")
        assert dc < 2e-3, f"{what}: {lvl} costs differ by {dc:.2e} away from any sort tie"
    print(what, "EPE full/precise/fine/coarse", " ".join(f"{e:.2e}" for e in epes),
          "| top-2 mismatches (decided/raw/pixels):", ", ".join(rep), f"| sort-tie neighbourhoods excluded: {100 * excluded:.2f} % of the image",
          flush=True)
    for a, b in zip(disps, rd):
        assert tuple(a.shape) == tuple(b.shape)
    # coarse candidates = sorted [integers 0..Dc-1 | two memory samples]: the integer entries (index work) must be exact
    # everywhere and sit at the same sorted positions; the memory entries are bilinear re-samplings (fp32 rounding)
    got_s, ref_s = samples[2].cpu(), rs[2]
    integral = (ref_s == ref_s.round()) & keep_c.unsqueeze(1)
    assert torch.equal(got_s[integral], ref_s[integral]), f"{what}: coarse candidate list differs"
    assert (got_s - ref_s).abs().amax(1)[keep_c].max() < 2e-5, f"{what}: coarse memory candidates differ"
    assert bad_total == 0, f"{what}: top-2 index mismatches outside the tie margin: {rep}"
    if strict:
        assert excluded == 0.0
        assert torch.equal(samples[2].cpu(), rs[2]) or has_memory, f"{what}: single-frame coarse candidates must be bit-exact"
    for i, e in enumerate(epes):
        assert e < EPE_TOL, f"{what}: disp{i} EPE {e:.3e} px"


def _cpu(x):
    if torch.is_tensor(x):
        return x.detach().cpu()
    if isinstance(x, dict):
        return {k: _cpu(v) for k, v in x.items()}
    return x


def _sequence(H, W, B, num_sample, T, seed0=40, engine_chain=True):
    """Two chains over T frames.

      A  the ORACLE carries its state; per frame the engine starts from the oracle's previous state (same-state comparison:
         EPE, exact index work), so one discontinuous flip cannot hide later frames.
      B  the ENGINE carries its OWN state through all T frames (update_map + forward on the device, nothing reset); per
         frame the oracle is evaluated from a CPU copy of that same engine state and must agree to the same bar.  Every step
         of the engine's trajectory is thereby a step the reference would have taken from the same state.

    The two trajectories themselves drift apart (printed, not asserted): the reference algorithm amplifies 1e-5-level state
    differences through its discontinuities (top-2 selection, the splat's x / (norm + 1e-22)) — the oracle re-run from its
    own state plus 1e-5 noise moves a few per cent of the pixels by more than 0.01 px as well."""
    from temporalstereo_b200 import temporal
    sd = synth.synthetic_state_dict(seed=0)
    eng = _engine(num_sample)
    st = synth.synthetic_temporal_state(H, W, B=B)
    pose = (st["K"], st["T_now"], st["inv_T_prev"], st["baseline"])
    dpose = [p.cuda() for p in pose]
    ref_state, own_state = {}, {}
    for t in range(T):
        lf, rf, li, ri = synth.synthetic_frame(H, W, B=B, seed=seed0 + t)
        dl, dr, dli, dri = _cuda(lf), _cuda(rf), li.cuda(), ri.cuda()
        # ---- chain A
        dev_in = _cuda(_copy(ref_state))
        if t:
            with torch.no_grad():
                ref_state = O.update_map(ref_state, *pose, H, W, True, 3)
            warped = temporal.update_map(dev_in, *dpose, H, W, True, 3)
            # the engine's warp against the oracle's, in the bulk (the splat's normalisation is discontinuous where almost
            # nothing lands; tests/test_gpu_ops.py pins the per-pixel values against the reference's own kernel)
            for got, want_ in ((warped["cost_memory"]["disp_sample"], ref_state["cost_memory"]["disp_sample"]),
                               (warped["cost_memory"]["cost_volume"], ref_state["cost_memory"]["cost_volume"]),
                               (warped["local_map"], ref_state["local_map"])):
                d = (got.cpu() - want_).abs()
                assert d.median() < 1e-5 and (d > 1e-3).float().mean() < 0.01, (t, d.median().item())
            dev_in = _cuda(_copy(ref_state))
        ref_in = _copy(ref_state)                                   # the state this frame is computed from (read-only copy)
        with torch.no_grad():
            want = O.aggregation_forward(sd, lf, rf, li, ri, _copy(ref_state), num_sample=num_sample)
        # (1) level by level on bit-identical candidates: the engine's fine / precise levels are centred on the ORACLE's
        #     up-sampled coarse / fine disparities, so no candidate sort can flip: strict everywhere, nothing excluded
        eng._inject = {"coarse_disp": want[0][3].cuda(), "fine_disp": want[0][2].cuda()}
        try:
            forced = eng(dl, dr, dli, dri, _cuda(_copy(ref_state)))
        finally:
            eng._inject = None
        check_frame(forced, want, f"A {H}x{W} D={16 * num_sample} B={B} frame {t} (levels on the oracle's candidates)", ref_in, strict=True)
        # (2) end to end, each level centred on the engine's own previous level
        out = eng(dl, dr, dli, dri, dev_in)
        n_fine = (min(t, 3) if t else 0) + 5 + 2
        assert out[2][1].shape[1] == n_fine == want[2][1].shape[1], "fine candidates: local map + 5 range + 2 memory"
        check_frame(out, want, f"A {H}x{W} D={16 * num_sample} B={B} frame {t} (end to end)", ref_in)
        ref_state = want[5]
        # ---- chain B
        if engine_chain:
            if t:
                own_state = temporal.update_map(own_state, *dpose, H, W, True, 3)
            snapshot = _cpu(_copy(own_state))                     # the state the engine aggregates from, for the oracle
            with torch.no_grad():
                want_b = O.aggregation_forward(sd, lf, rf, li, ri, _copy(snapshot), num_sample=num_sample)
            eng._inject = {"coarse_disp": want_b[0][3].cuda(), "fine_disp": want_b[0][2].cuda()}
            try:
                forced = eng(dl, dr, dli, dri, _cuda(_copy(snapshot)))
            finally:
                eng._inject = None
            check_frame(forced, want_b, f"B {H}x{W} D={16 * num_sample} B={B} frame {t} (engine-carried state, levels on the oracle's candidates)",
                        snapshot, strict=True)
            own = eng(dl, dr, dli, dri, own_state)                # the natural run: this is what is carried forward
            own_state = own[5]
            check_frame(own, want_b, f"B {H}x{W} D={16 * num_sample} B={B} frame {t} (engine-carried state, end to end)", snapshot)
            d = (own[0][0].cpu() - want[0][0]).abs()
            print(f"  chains A/B apart at frame {t}: median {d.median().item():.2e} px, > 0.01 px: {100 * (d > 1e-2).float().mean().item():.2f} %")


def test_c4_tartanair_sequence_t5():
    """BASELINE config C4 at its stated shape: 480x640, D=320, T=5, B=2 per rank."""
    _sequence(480, 640, 2, 20, 5)


def test_c5_1080p_temporal():
    """BASELINE config C5 at its stated shape: 1088x1920 (1080 padded to x16), D=256, T=3 -> the three distinct frame kinds
    (no state / first warp / local map growing), B=2."""
    _sequence(1088, 1920, 2, 16, 3, engine_chain=False)


def test_captured_step_replays_bit_identically():
    """CUDA-graph capture of update_map + forward: a replay equals the eager call bit for bit (no atomics on the
    aggregation path; the splat's float atomics make the temporal half order-dependent, so it is compared in the bulk)."""
    from temporalstereo_b200.graph import CapturedStep
    eng = _engine(12)
    H, W, B = 128, 192, 2
    lf, rf, li, ri = synth.synthetic_frame(H, W, B=B, seed=7)
    dl, dr, dli, dri = _cuda(lf), _cuda(rf), li.cuda(), ri.cuda()
    eager = eng(dl, dr, dli, dri, {})
    step = CapturedStep(lambda: eng(dl, dr, dli, dri, {}))
    assert step.launches > 50
    out = step.replay()
    torch.cuda.synchronize()
    for a, b in zip(eager[0] + eager[1] + eager[2] + eager[3], out[0] + out[1] + out[2] + out[3]):
        assert torch.equal(a, b)
    # refill the input buffers in place, replay: equals an eager call on the new data
    lf2, rf2, li2, ri2 = synth.synthetic_frame(H, W, B=B, seed=8)
    for d, s in zip(dl + dr + [dli, dri], lf2 + rf2 + [li2, ri2]):
        d.copy_(s)
    out = step.replay()
    torch.cuda.synchronize()
    again = eng(dl, dr, dli, dri, {})
    for a, b in zip(again[0] + again[1], out[0] + out[1]):
        assert torch.equal(a, b)


def test_stereo_engine_forward_left_right_prev_state():
    """forward(left, right, prev_state) -> full-resolution disparity over a 3-frame sequence with the pose warp, against the
    oracle chain (update_map -> aggregation -> upsample_disps, projects/TemporalStereo/TemporalStereo.py:292-309)."""
    from temporalstereo_b200.stereo import StereoEngine
    H, W, B = 96, 160, 1
    sd = synth.synthetic_state_dict(seed=0)
    model = StereoEngine(_engine(12))
    st = synth.synthetic_temporal_state(H, W, B=B)
    pose = dict(K=st["K"].cuda(), T=st["T_now"].cuda(), inv_T_prev=st["inv_T_prev"].cuda(), baseline=st["baseline"].cuda())
    ref_state = {}
    for t in range(3):
        lf, rf, li, ri = synth.synthetic_frame(H, W, B=B, seed=50 + t)
        state = _cuda(_copy(ref_state))              # same-state comparison (see _sequence)
        if t:
            with torch.no_grad():
                ref_state = O.update_map(ref_state, st["K"], st["T_now"], st["inv_T_prev"], st["baseline"], H, W, True, 3)
        with torch.no_grad():
            want = O.aggregation_forward(sd, lf, rf, li, ri, ref_state)
            want_full = O.upsample_disps(want[0], H, W)
        disp = model(li.cuda(), ri.cuda(), state, feats=(_cuda(lf), _cuda(rf)), pose=pose)
        assert disp.shape == (B, 1, H, W) and state["prev_disp"] is not None and "cost_memory" in state
        for a, b in zip(model.last["disps"], want_full):
            assert a.shape == b.shape == (B, 1, H, W)
            assert (a.cpu() - b).abs().mean() < EPE_TOL * (W / 20.0), "up-sampled disparity (values scale with full_w / dw)"
        assert (disp.cpu() - want_full[0]).abs().mean() < EPE_TOL
        ref_state = want[5]
