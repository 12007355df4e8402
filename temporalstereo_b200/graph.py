"""CUDA-graph capture of the hot path.

A frame is ~110 small launches; issued one by one from Python they cost ~20 us each on the host, which is more than most
of the kernels run for (at B=1 the eager frame is entirely launch-bound).  `CapturedStep` records any callable made of
libtstereo ops — one aggregation forward, or update_map + forward — into a CUDA graph once and replays it with a single
launch.  Static-shape contract of CUDA graphs: the callable's INPUT tensors are the graph's input buffers (refill them in
place, e.g. with `copy_` from pinned host memory, then `replay()`), and the returned tensors are the graph's output buffers,
overwritten by the next replay.

The reference has no counterpart (it issues ~580 eager ATen ops per frame); SURVEY.md §7 step 10.
"""
from __future__ import annotations

from typing import Any, Callable

import torch

from . import _lib


class CapturedStep:
    """graph = CapturedStep(fn); out = graph.replay() — `fn()` must be repeatable (its warm-up runs it eagerly first:
    weight packing, kernel attribute set-up and the plan all happen outside the capture)."""

    def __init__(self, fn: Callable[[], Any], warmup: int = 1, device=None):
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        lib = _lib.load()
        # the graph's kernels hold raw pointers into every tensor `fn` closes over (its input buffers): keep the callable —
        # and with it those tensors — alive as long as the graph.  (A caller that dropped them got replays over recycled
        # memory, and an illegal address once torch.cuda.empty_cache() unmapped the segment.)
        self._fn = fn
        with torch.cuda.device(dev):
            cur = torch.cuda.current_stream(dev)
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                for _ in range(max(warmup, 1)):
                    fn()
            cur.wait_stream(side)
            torch.cuda.synchronize(dev)
            self.graph = torch.cuda.CUDAGraph()
            n0 = lib.tstereo_launch_count()
            with torch.cuda.graph(self.graph):
                self.outputs = fn()
            # kernels of libtstereo inside one replay (the launch counter ticks at capture time only)
            self.launches = int(lib.tstereo_launch_count() - n0)
        self.device = dev
        self.replays = 0

    def replay(self):
        self.graph.replay()
        self.replays += 1
        return self.outputs
