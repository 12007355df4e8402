#!/usr/bin/env python
"""bench.py — stereo frames/sec of the cost-volume hot path at BASELINE.json's headline config.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

One "step" = one pass of the per-frame hot path (3 cost volumes -> 3-level separable 3-D aggregation
-> top-2 soft-argmin -> convex / UNet up-sampling) over one batch of B
synthetic 544x960 (540x960 padded to a multiple of 16) D=192 stereo frames per GPU (B = 8 by default).  Prints ONE JSON
line on rank 0 (see DESIGN.md §measurement for every field).

  value      frames/s, inputs resident in HBM, whole job (all ranks), CUDA-event timed, max over ranks
  e2e        the same through the public nn.Module call with HOST (pinned) inputs: H2D of the feature
             pyramids + images and D2H of the full-resolution disparity inside the timed region
  roofline   the cost-volume kernels (block_cost, the path north_star sets the HBM target on) timed
             alone with CUDA events: algorithmic bytes / time vs MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the CPU oracle port of the reference forward on the box's host cores (bounded sample)
  --impl reference   times that CPU path as the reference arm (rank 0 only)
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "stereo frames/sec @540x960 D=192"
UNIT = "frames/s"
L2_BYTES = 126 * 1024 * 1024


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="frames per GPU per step (SURVEY.md §8d: C2 at B=8)")
    ap.add_argument("--height", type=int, default=544)
    ap.add_argument("--width", type=int, default=960)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the cpu_baseline sample")
    return ap.parse_args()


def workload_name(a):
    return (f"C2: synthetic {a.height}x{a.width} (540x960 padded to x16) D=192 single-frame "
            f"cost volume + 3-D aggregation, B={a.batch}/GPU")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi sampled every 200 ms while the timed region runs (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None
        return self

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:  # noqa: BLE001
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nme, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU path (oracle port of the reference)
def cpu_reference_frames(a, seconds: float, max_frames: int = 8):
    """Times the CPU oracle (a functional restatement of the reference's own PyTorch forward, pinned to
    the real reference's outputs by tests/test_oracle_golden.py) on the host cores.  B=1 frames of the
    bench workload, 1 warm-up + up to `max_frames` timed within `seconds`."""
    from oracle import oracle as O
    from temporalstereo_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = synth.synthetic_state_dict(seed=0)
    lf, rf, li, ri = synth.synthetic_frame(a.height, a.width, B=1, seed=1)
    times = []
    with torch.no_grad():
        O.aggregation_forward(sd, lf, rf, li, ri, {})
        t_end = time.perf_counter() + seconds
        while len(times) < max_frames and (not times or time.perf_counter() < t_end):
            t0 = time.perf_counter()
            O.aggregation_forward(sd, lf, rf, li, ri, {})
            times.append(time.perf_counter() - t0)
    med = statistics.median(times)
    return 1.0 / med, med, len(times), torch.get_num_threads()


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    from temporalstereo_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = synth.synthetic_state_dict(seed=0)
    lf, rf, li, ri = synth.synthetic_frame(a.height, a.width, B=1, seed=1)
    # a step of the reference arm = ONE frame of the workload (bounded sample of the B-frame batch)
    steps = min(a.steps, 10)
    warm = min(a.warmup, 2)
    with torch.no_grad():
        for _ in range(max(warm, 1)):
            O.aggregation_forward(sd, lf, rf, li, ri, {})
        t0 = time.perf_counter()
        for _ in range(steps):
            O.aggregation_forward(sd, lf, rf, li, ri, {})
        dt = time.perf_counter() - t0
    fps = steps / dt
    sample = f"{steps} single frames (B=1) of the workload, fp32, torch CPU ops, {torch.get_num_threads()} threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": a.gpus, "steps": steps,
        "warmup": max(warm, 1), "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "note": "CPU path: oracle port of the reference forward "
                   "(the Python reference tree is not present on the GPU box); one frame per step"},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ----------------------------------------------------------------------------- GPU arm
def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
        return

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback for the hot path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=dev)

    from temporalstereo_b200 import _lib, ops, shard, synth
    from temporalstereo_b200.aggregation import TEMPORALSTEREO
    lib = _lib.load()

    H, W, B = a.height, a.width, a.batch
    eng = TEMPORALSTEREO()
    eng.load_state_dict(synth.synthetic_state_dict(seed=0), strict=True)
    eng = eng.to(dev).eval()

    # ---- inputs: host (pinned) master copy + NSETS device-resident sets rotated so that consecutive
    #      steps never re-read inputs from L2 (sets differ by a horizontal roll)
    lf, rf, li, ri = synth.synthetic_frame(H, W, B=B, seed=1 + rank)
    host = [t.pin_memory() for t in (lf + rf + [li, ri])]
    in_bytes = sum(t.numel() * 4 for t in host)
    nsets = max(2, -(-2 * L2_BYTES // in_bytes))
    sets = []
    for s in range(nsets):
        sets.append([torch.roll(t.to(dev), shifts=s, dims=-1).contiguous() for t in host])

    def forward(inp):
        return eng(inp[0:3], inp[3:6], inp[6], inp[7], {})

    def barrier():
        shard.barrier(dist)

    # ---- device-resident throughput
    for i in range(max(a.warmup, 3)):
        out = forward(sets[i % nsets])
    barrier()
    n0 = lib.tstereo_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        barrier()
        e0.record()
        for i in range(a.steps):
            out = forward(sets[i % nsets])
        e1.record()
        barrier()
    launches = lib.tstereo_launch_count() - n0
    # whole-job throughput: frames of all ranks / slowest rank's device time
    fps, ms, _ = shard.aggregate_throughput(B * a.steps, e0.elapsed_time(e1), dist, dev)

    # ---- end to end through the public module call with host buffers: every step uploads its own inputs from
    #      pinned host memory and downloads its full-resolution disparity.  Two staging sets: the upload of
    #      step i+1 runs on a copy stream while step i computes (the steady state of a streaming caller).
    stages = [[torch.empty_like(t, device=dev) for t in host] for _ in range(2)]
    full_host = [torch.empty((B, 1, H, W), dtype=torch.float32).pin_memory() for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream(dev)
    uploaded = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def upload(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i % 2])          # the set is free once the step that read it is done
            for d, h in zip(stages[i % 2], host):
                d.copy_(h, non_blocking=True)
            uploaded[i % 2].record(copy_stream)

    def e2e_run(n):
        for ev in consumed:
            ev.record(main_stream)
        upload(0)
        for i in range(n):
            if i + 1 < n:
                upload(i + 1)
            main_stream.wait_event(uploaded[i % 2])
            o = forward(stages[i % 2])
            consumed[i % 2].record(main_stream)
            full_host[i % 2].copy_(o[0][0], non_blocking=True)

    e2e_run(3)
    barrier()
    e0.record()
    e2e_run(a.steps)
    e1.record()
    barrier()
    fps_e2e, ms_e2e, _ = shard.aggregate_throughput(B * a.steps, e0.elapsed_time(e1), dist, dev)

    # ---- roofline of the cost-volume operator (block_cost: the path north_star sets the HBM target on).
    #      Dominant launch = the precise-level volume (198 of the 329 MB/frame); the three-level total
    #      is reported beside it.  Candidates are the engine's own pattern: a piecewise-smooth disparity
    #      +- {4,1,0} (a trained network regresses smooth maps; random-init weights do not).
    peak, peak_src = peaks()

    def smooth_samples(h, w, S):
        yy, xx = torch.meshgrid(torch.arange(h, device=dev), torch.arange(w, device=dev), indexing="ij")
        base = 0.06 * w * (1.2 + torch.sin(xx / w * 6.0) * torch.cos(yy / h * 4.0)) + 0.3 * torch.rand(h, w, device=dev)
        offs = torch.tensor([-4.0, -1.0, 0.0, 1.0, 4.0], device=dev)[:S]
        return (base[None, None] + offs.view(1, S, 1, 1)).expand(B, S, h, w).contiguous()

    nrot = max(2, -(-2 * L2_BYTES // (8 * B * 128 * (H // 4) * (W // 4))))
    lcat = [torch.randn(B, 128, H // 4, W // 4, device=dev) for _ in range(nrot)]
    rcat = [torch.randn(B, 128, H // 4, W // 4, device=dev) for _ in range(nrot)]
    smp4, smp8 = smooth_samples(H // 4, W // 4, 5), smooth_samples(H // 8, W // 8, 5)

    def timed(fn, reps):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        e0.record()
        for i in range(reps):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    def vol_bytes(C, h, w, S, planes, warp=True):
        return 4 * B * (2 * C * h * w + (S * h * w if warp else 0) + planes * S * h * w)

    reps = max(10, min(a.steps, 50))
    b_p = vol_bytes(128, H // 4, W // 4, 5, 304)
    b_f = vol_bytes(128, H // 8, W // 8, 5, 304)
    b_c = vol_bytes(256, H // 16, W // 16, 12, 352, warp=False)
    ms_p = timed(lambda i: ops.block_cost(lcat[i % nrot], rcat[i % nrot], smp4), reps)
    ms_f = timed(lambda i: ops.block_cost(sets[i % nsets][1], sets[i % nsets][4], smp8), reps)
    ms_c = timed(lambda i: ops.block_cost(sets[i % nsets][2], sets[i % nsets][5], 12), reps)
    achieved = b_p / (ms_p * 1e-3) / 1e9
    cv_ms, alg_bytes = ms_p + ms_f + ms_c, b_p + b_f + b_c

    # ---- the heaviest single launch of the step: the first (1,3,3) conv of the precise level (304 -> 8 channels over
    #      the raw volume) on the tensor-core kernel.  It is HBM-bound (36 FLOP/B): algorithmic bytes =
    #      volume read once + output written once.
    pk_first = eng._pk["precise.init3d.0.conv.0"]
    vols = [torch.randn(B, 304, 5, H // 4, W // 4, device=dev) for _ in range(2)]
    out8 = torch.empty(B, 8, 5, H // 4, W // 4, device=dev)
    ms_conv = timed(lambda i: ops.conv_hw3_tc2(vols[i % 2], pk_first.tc["hw3"], pk_first.b, 8, 1, "SiLU", out=out8, half=eng.half_split), reps)
    b_conv = 4 * B * (304 + 8) * 5 * (H // 4) * (W // 4)
    del vols

    # DRAM traffic of the roofline kernel from the committed ncu capture (profiles/r01_traffic.json, per launch at the
    # batch size it names); null when the bench runs at another batch size
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        if int(tj.get("batch", -1)) == B and (H, W) == (544, 960):
            traffic = tj["block_cost_precise_dram_bytes"]
    except (OSError, ValueError, KeyError):
        pass

    result = {
        "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
        "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "frames_per_step_per_gpu": B,
                   "l2": f"inputs rotate over {nsets} device-resident sets ({nsets * in_bytes / 1e6:.0f} MB > 126 MB L2); "
                         "every intermediate cost volume alone exceeds L2",
                   "parallelism": f"batch-sharded replicas x{world}, no data-path collective"},
        "clocks": clk.summary(),
        "e2e": {"value": fps_e2e, "unit": UNIT, "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": B * H * W * 4,
                "ms_per_step": ms_e2e / a.steps},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm",
                     "kernel": "block_cost_warp at the precise level (block_cost_main_kernel<1,1> + block_cost_resize_kernel)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": "ncu --set full dram__bytes_read.sum + dram__bytes_write.sum of both "
                     "kernels, profiles/r01_traffic.json", "peak_source": peak_src, "algorithmic_bytes": b_p, "ms": ms_p,
                     "all_three_levels": {"algorithmic_bytes": alg_bytes, "ms": cv_ms,
                                          "achieved": alg_bytes / (cv_ms * 1e-3) / 1e9,
                                          "frac": alg_bytes / (cv_ms * 1e-3) / 1e9 / peak},
                     "share_of_step": cv_ms / (ms / a.steps)},
        "roofline_conv": {"bound": "hbm", "kernel": "conv_tc2_kernel<8,4,ACC>: first (1,3,3) conv of the precise level, "
                          "304 -> 8 channels over the raw cost volume (tcgen05, 3xTF32)",
                          "achieved": b_conv / (ms_conv * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                          "frac": b_conv / (ms_conv * 1e-3) / 1e9 / peak, "algorithmic_bytes": b_conv, "ms": ms_conv,
                          "tflops_fp32_equiv": 2.0 * B * 304 * 8 * 9 * 5 * (H // 4) * (W // 4) / (ms_conv * 1e-3) / 1e12,
                          "share_of_step": ms_conv / (ms / a.steps)},
    }
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cfps, med, n, thr = cpu_reference_frames(a, a.cpu_seconds)
        result["cpu_baseline"] = {"value": cfps, "unit": UNIT, "cores": thr, "kind": "port",
                                  "sample": f"median of {n} single frames (B=1) of the workload after 1 warm-up, "
                                            f"{med:.3f} s/frame, CPU oracle port of the reference forward"}
    if rank == 0:
        print(json.dumps(result))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
