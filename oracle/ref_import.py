"""Import the UNMODIFIED reference from /root/reference through the stub packages in
oracle/refstubs.  Build-container only: /root/reference does not
exist on the GPU box, so nothing under tests -m gpu / smoke() / bench.py calls this."""
import os
import sys
import warnings

REF = os.environ.get("TSTEREO_REFERENCE", "/root/reference")
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refstubs")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "architecture"))


def setup():
    if not available():
        raise RuntimeError(f"reference tree not found at {REF}")
    for p in (os.path.join(REF, "projects", "TemporalStereo"), REF, _STUBS):
        if p not in sys.path:
            sys.path.insert(0, p)
    warnings.filterwarnings("ignore", category=SyntaxWarning)


def build_reference_aggregation(levels=None):
    """Reference TEMPORALSTEREO aggregation with the shipped hyper-parameters
    (reference configs/sceneflow.yaml:38-60)."""
    setup()
    from architecture.modeling.aggregation.TemporalStereo.coarse import CoarseAggregation
    from architecture.modeling.aggregation.TemporalStereo.fine import FineAggregation
    from architecture.modeling.aggregation.TemporalStereo.precise import PreciseAggregation
    from architecture.modeling.aggregation.TemporalStereo.TemporalStereo import TEMPORALSTEREO
    from temporalstereo_b200.synth import DEFAULT_LEVELS
    lv = levels or DEFAULT_LEVELS
    c, f, p = lv["coarse"], lv["fine"], lv["precise"]
    return TEMPORALSTEREO(
        coarse=CoarseAggregation(c["in_planes"], c["C"], c["num_sample"], 1.0, 3, 2, True),
        fine=FineAggregation(f["in_planes"], f["C"], f["num_sample"], 1.0, 3, 2, True),
        precise=PreciseAggregation(p["in_planes"], p["C"], p["num_sample"], 1.0, 3, 2),
    ).eval()


def build_reference_temporal(cpu_splat, use_past_cost=True, local_map_size=3):
    """The reference LightningModule shell, only far enough to call update_map
    (projects/TemporalStereo/TemporalStereo.py:326-461).  The CuPy splat cannot run on CPU
    (softsplat.py:269-270), so the module-level name is replaced by `cpu_splat`."""
    setup()
    import torch.nn as nn
    import TemporalStereo as TS          # projects/TemporalStereo/TemporalStereo.py
    m = TS.TemporalStereo.__new__(TS.TemporalStereo)
    nn.Module.__init__(m)
    m.with_previous = True
    m.use_past_cost = use_past_cost
    m.local_map_size = local_map_size
    TS.FunctionSoftsplat = lambda tenInput, tenFlow, tenMetric, strType: cpu_splat(tenInput, tenFlow, tenMetric)
    return m
