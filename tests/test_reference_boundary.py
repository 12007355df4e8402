"""The drop-in boundary exercised from the REFERENCE's side, in the build container where
/root/reference exists: the reference's own `build_aggregation` + its shipped YAML must hand back the B200 engine after
`install_into_reference()`, a state dict of the reference's module must load with strict=True, and the engine must survive
the eval()/train() toggling of the reference trainer (projects/TemporalStereo/TemporalStereo.py:268-274).  CPU only — no
kernel is launched; skipped on the GPU box (no reference tree there)."""
import os

import pytest
import torch

from oracle import ref_import

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference tree not present (GPU box)")


def _cfg(name="sceneflow.yaml"):
    import yaml
    ref_import.setup()
    from fvcore.common.config import CfgNode          # oracle/refstubs stand-in with attribute access + .get
    path = os.path.join(ref_import.REF, "projects", "TemporalStereo", "configs", name)
    return CfgNode(yaml.safe_load(open(path)))


def test_reference_builder_returns_the_engine_and_loads_reference_weights():
    ref_import.setup()
    from architecture.modeling.aggregation import builder as ref_builder
    from architecture.modeling.aggregation.TemporalStereo.TemporalStereo import TEMPORALSTEREO as RefAgg
    from temporalstereo_b200.aggregation import TEMPORALSTEREO
    from temporalstereo_b200.registry import install_into_reference, registry_map

    cfg = _cfg()
    ref_module = RefAgg(cfg)                            # the reference's own class, its own from_config
    table = registry_map(ref_builder.AGGREGATION_REGISTRY)
    saved = table["TEMPORALSTEREO"]
    try:
        install_into_reference()
        eng = ref_builder.build_aggregation(cfg)        # reference aggregation/builder.py:12-20, unchanged
    finally:
        table["TEMPORALSTEREO"] = saved
    assert isinstance(eng, TEMPORALSTEREO) and not isinstance(eng, RefAgg)
    assert eng.levels["coarse"]["num_sample"] == cfg.MODEL.AGGREGATION.COARSE.NUM_SAMPLE
    assert eng.levels["precise"]["in_planes"] == cfg.MODEL.AGGREGATION.PRECISE.IN_PLANES

    # reference checkpoints load with strict=True (projects/TemporalStereo/demo.py:250-251): same keys, same shapes
    sd = ref_module.state_dict()
    assert list(sd) == list(eng.state_dict()), "state-dict keys / order differ from the reference module"
    missing, unexpected = eng.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    for k, v in eng.state_dict().items():
        assert torch.equal(v, sd[k]), k
    # ... and with the Lightning checkpoint prefix stripped the way demo.py does it
    ckpt = {"state_dict": {"aggregation." + k: v for k, v in sd.items()}}
    eng.load_state_dict({k[len("aggregation."):]: v for k, v in ckpt["state_dict"].items()}, strict=True)

    # the trainer's mode toggling (TemporalStereo.py:268-274; Lightning's on_*_model_train): no exception, stays in eval
    eng.eval()
    with pytest.warns(UserWarning):
        eng.train()
    eng.eval()
    eng.train()
    assert not eng.training
    # fine.phi never receives a gradient in the reference either (fine.py:33): kept out of the trainable set
    assert not dict(eng.named_parameters())["fine.phi"].requires_grad


@pytest.mark.parametrize("yaml_name", ["sceneflow.yaml", "kitti2015.yaml", "tartanair.yaml"])
def test_every_shipped_config_builds(yaml_name):
    path = os.path.join(ref_import.REF, "projects", "TemporalStereo", "configs", yaml_name)
    if not os.path.exists(path):
        pytest.skip(f"{yaml_name} not shipped")
    from temporalstereo_b200.aggregation import build_aggregation
    eng = build_aggregation(_cfg(yaml_name))
    assert len(eng.state_dict()) == 526


def test_layers_level_replacement_keeps_the_reference_signatures():
    """FunctionSoftsplat / project_to_3d are replaced by name in the reference's LightningModule module
    (projects/TemporalStereo/TemporalStereo.py:21); the signatures must agree."""
    import inspect
    ref_import.setup()
    import importlib
    ref_splat = importlib.import_module('architecture.modeling.layers.softsplat')
    ref_warp = importlib.import_module('architecture.modeling.layers.inverse_warp')
    from temporalstereo_b200 import temporal
    assert list(inspect.signature(temporal.FunctionSoftsplat).parameters) == list(inspect.signature(ref_splat.FunctionSoftsplat).parameters)
    ours = list(inspect.signature(temporal.project_to_3d).parameters)
    theirs = list(inspect.signature(ref_warp.project_to_3d).parameters)
    assert ours[:len(theirs)] == theirs or theirs[:len(ours)] == ours, (ours, theirs)
