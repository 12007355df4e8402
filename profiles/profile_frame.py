"""One frame of the bench workload between cudaProfilerStart/Stop, for ncu --profile-from-start off.

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python profiles/profile_frame.py [--batch B] [--temporal]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from temporalstereo_b200 import synth, temporal  # noqa: E402
from temporalstereo_b200.aggregation import TEMPORALSTEREO  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--height", type=int, default=544)
ap.add_argument("--width", type=int, default=960)
ap.add_argument("--temporal", action="store_true")
a = ap.parse_args()

eng = TEMPORALSTEREO()
eng.load_state_dict(synth.synthetic_state_dict(seed=0), strict=True)
eng = eng.cuda().eval()
lf, rf, li, ri = synth.synthetic_frame(a.height, a.width, B=a.batch, seed=1)
lf, rf, li, ri = [t.cuda() for t in lf], [t.cuda() for t in rf], li.cuda(), ri.cuda()
st = synth.synthetic_temporal_state(a.height, a.width, B=a.batch)
pose = [st[k].cuda() for k in ("K", "T_now", "inv_T_prev", "baseline")]


def frame(prev):
    if a.temporal and "prev_disp" in prev:
        prev = temporal.update_map(prev, *pose, a.height, a.width, True, 3)
    return eng(lf, rf, li, ri, prev)[5]


prev = {}
for _ in range(5 if a.temporal else 2):   # temporal: the local map grows to 3 planes over the first frames (new layer shapes)
    prev = frame(prev)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
frame(prev)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
