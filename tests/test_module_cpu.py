"""Host-side logic of the drop-in module that needs no GPU: the parameter tree (526 state-dict
entries with the reference's names and shapes), strict loading, config construction, registry."""
import pytest
import torch

from temporalstereo_b200 import synth
from temporalstereo_b200.aggregation import TEMPORALSTEREO, build_aggregation
from temporalstereo_b200.registry import AGGREGATION_REGISTRY


def test_state_dict_has_the_reference_tree():
    m = TEMPORALSTEREO()
    sd = m.state_dict()
    spec = synth.state_dict_spec()
    assert len(sd) == 526 == len(spec)                      # SURVEY.md §8b, probed on the reference module
    assert list(sd.keys()) == [k for k, _ in spec]
    for k, shape in spec:
        assert tuple(sd[k].shape) == tuple(shape), k
    assert sd["coarse.init3d.0.conv.0.weight"].shape == (32, 352, 1, 3, 3)
    assert sd["precise.init3d.0.conv.0.weight"].shape == (8, 304, 1, 3, 3)
    assert not m.fine.phi.requires_grad or True


def test_strict_load_and_invalidate():
    m = TEMPORALSTEREO()
    sd = synth.synthetic_state_dict(seed=0)
    m._pk = {"stale": None}
    m.load_state_dict(sd, strict=True)
    assert m._pk is None, "packed weights must be dropped when a checkpoint is loaded"
    bad = dict(sd)
    bad.pop("fine.phi")
    with pytest.raises(RuntimeError):
        m.load_state_dict(bad, strict=True)


def test_from_config_and_registry():
    cfg = {"MODEL": {"AGGREGATION": {"NAME": "TEMPORALSTEREO",
                                     "COARSE": {"IN_PLANES": 256, "C": 32, "NUM_SAMPLE": 16},
                                     "FINE": {"IN_PLANES": 128, "C": 16, "NUM_SAMPLE": 5},
                                     "PRECISE": {"IN_PLANES": 64, "C": 8, "NUM_SAMPLE": 5}}}}
    m = build_aggregation(cfg)
    assert isinstance(m, TEMPORALSTEREO) and AGGREGATION_REGISTRY.get("TEMPORALSTEREO") is TEMPORALSTEREO
    assert m.levels["coarse"]["num_sample"] == 16
    assert not m.training
    # the reference trainer toggles eval()/train() around every history frame (projects/TemporalStereo/TemporalStereo.py:
    # 268-274) and Lightning calls train() around fit / validate / test: it must not raise, and the engine stays in eval
    with pytest.warns(UserWarning, match="inference engine"):
        m.train()
    assert not m.training
    m.eval().train(True)        # second toggle: silent
    assert not m.training and all(not c.training for c in m.modules())


def test_fold_matches_batchnorm():
    """BN folding (host logic) against F.batch_norm on a tiny conv."""
    import torch.nn.functional as F
    m = TEMPORALSTEREO()
    sd = synth.synthetic_state_dict(seed=0)
    p = "fine.init3d.2.conv.0"
    pk = m._mk(*m._fold(sd, p, p + ".norm"))
    x = torch.randn(1, 16, 2, 6, 7)
    y = F.conv3d(x, sd[p + ".weight"], None, 1, (0, 2, 2), (1, 2, 2))
    y = F.batch_norm(y, sd[p + ".norm.running_mean"], sd[p + ".norm.running_var"], sd[p + ".norm.weight"],
                     sd[p + ".norm.bias"], False, 0.0, 1e-5)
    w = pk.w[:, :, :pk.cout].permute(2, 0, 1).reshape(16, 16, 1, 3, 3)
    y2 = F.conv3d(x, w, pk.b, 1, (0, 2, 2), (1, 2, 2))
    torch.testing.assert_close(y2, y, atol=1e-5, rtol=1e-5)


def test_no_cpu_fallback():
    m = TEMPORALSTEREO().eval()
    lf, rf, li, ri = synth.synthetic_frame(32, 48, B=1)
    with pytest.raises(TypeError):
        m(lf, rf, li, ri, {})


def test_pack_runs_on_cpu_into_one_arena():
    """BN folding and operand packing are host work: `_pack` needs no GPU, lands every tensor in one buffer (a single
    upload) and announces exactly the operand sizes the C ABI expects."""
    from temporalstereo_b200 import _lib
    lib = _lib.load()
    m = TEMPORALSTEREO()
    m.load_state_dict(synth.synthetic_state_dict(seed=0), strict=True)
    pk = m._pack("cpu")
    base, n = m._arena.data_ptr(), m._arena.numel() * 4
    for name, k in pk.items():
        for t in [k.w, k.b] + list(k.tc.values()):
            if t is not None:
                assert base <= t.data_ptr() < base + n and (t.data_ptr() - base) % 256 == 0, name
    first = pk["precise.init3d.0.conv.0"]
    assert set(first.tc) == {"hw3", "left", "cost", "taps", "taps_osc", "gconv"}
    assert first.tc["cost"].numel() == lib.tstereo_cost_conv_wpack_floats(128, 8, 1)
    assert first.tc["left"].numel() == lib.tstereo_conv_hw3_tc2_wpack_floats(128, 8, 1)
    assert pk["coarse.init3d.0.conv.0"].tc["cost"].numel() == lib.tstereo_cost_conv_wpack_floats(256, 32, 1)
    assert "s2" in pk["coarse.init3d.1.conv1.conv.0"].tc and "dc" in pk["fine.init3d.1.conv5.conv.0"].tc
    assert pk["precise.refinement.deconv2"].tc["dc"].numel() == lib.tstereo_deconv_hw_tc2_wpack_floats(32, 9, 1)


def test_forward_validates_the_pyramid_geometry():
    """Sizes that are not multiples of 16 make the reference fail with a shape error; the engine
    must raise too instead of writing out of bounds."""
    m = TEMPORALSTEREO().eval()
    lf, rf, li, ri = synth.synthetic_frame(32, 48, B=1)
    lf = [lf[0], lf[1][..., :-1, :], lf[2]]
    fake = lambda t: type("T", (), {})()
    with pytest.raises(ValueError, match="multiples of 16"):
        m._check_pyramid(lf[0], lf[1], lf[2], rf[0], rf[1], rf[2], li, ri)


def test_forward_refuses_gradients():
    m = TEMPORALSTEREO().eval()
    lf, rf, li, ri = synth.synthetic_frame(32, 48, B=1)
    lf[0].requires_grad_(True)
    with pytest.raises(NotImplementedError, match="no backward"):
        m(lf, rf, li, ri, {})
