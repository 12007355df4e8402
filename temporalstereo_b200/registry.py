"""`AGGREGATION_REGISTRY` — the registry the reference builds its aggregation module from
(reference architecture/modeling/aggregation/builder.py:1-20).

When detectron2 is importable its `Registry` class is used, so the object is interchangeable with
the reference's; otherwise a minimal class with the same `register()` / `get()` behaviour stands in
(detectron2 is not installable in the build image).
"""
from __future__ import annotations

try:  # pragma: no cover - detectron2 is absent in the build image
    from detectron2.utils.registry import Registry
except Exception:  # noqa: BLE001
    class Registry:
        def __init__(self, name: str):
            self._name = name
            self._obj_map = {}

        def _do_register(self, name, obj):
            if name in self._obj_map:
                raise AssertionError(f"An object named '{name}' was already registered in '{self._name}' registry!")
            self._obj_map[name] = obj

        def register(self, obj=None):
            if obj is None:
                def deco(func_or_class):
                    self._do_register(func_or_class.__name__, func_or_class)
                    return func_or_class
                return deco
            self._do_register(obj.__name__, obj)
            return obj

        def get(self, name):
            ret = self._obj_map.get(name)
            if ret is None:
                raise KeyError(f"No object named '{name}' found in '{self._name}' registry!")
            return ret

        def __contains__(self, name):
            return name in self._obj_map


AGGREGATION_REGISTRY = Registry("AGGREGATION")
AGGREGATION_REGISTRY.__doc__ = "Registry for cost aggregation modules (callables taking a config, returning nn.Module)."


def install_into_reference() -> None:
    """Swap the engine in for the reference's own class: after this call the reference's
    `build_aggregation(cfg)` with `MODEL.AGGREGATION.NAME == 'TEMPORALSTEREO'` returns the B200
    engine.  Requires the reference package to be importable."""
    from architecture.modeling.aggregation import builder as ref_builder  # type: ignore
    import architecture.modeling.aggregation  # noqa: F401  (registers the reference's own class first)
    from .aggregation import TEMPORALSTEREO
    registry_map(ref_builder.AGGREGATION_REGISTRY)["TEMPORALSTEREO"] = TEMPORALSTEREO


def registry_map(reg) -> dict:
    """The name -> class table of a detectron2 / fvcore `Registry` (`_obj_map`) or of a stand-in with the same
    `register()` / `get()` behaviour."""
    for attr in ("_obj_map", "_map"):
        m = getattr(reg, attr, None)
        if isinstance(m, dict):
            return m
    raise TypeError(f"{type(reg).__name__} exposes no name table")
