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
    ap.add_argument("--eager", action="store_true", help="launch kernel by kernel instead of replaying a CUDA graph per step")
    ap.add_argument("--no-extras", action="store_true", help="skip the B=1 latency and C3 temporal side measurements")
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

    from temporalstereo_b200 import _lib, ops, shard, synth, temporal
    from temporalstereo_b200.aggregation import TEMPORALSTEREO
    from temporalstereo_b200.graph import CapturedStep
    # pin the process next to its GPU BEFORE any pinned staging buffer is allocated (first touch decides the NUMA node)
    numa_cpus = shard.bind_to_gpu_numa(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    H, W, B = a.height, a.width, a.batch
    eng = TEMPORALSTEREO()
    eng.load_state_dict(synth.synthetic_state_dict(seed=0), strict=True)
    eng = eng.to(dev).eval()

    # ---- inputs.  Host side (pinned): the fp32 feature pyramids as the backbone hands them over, and the two images in
    #      their wire format: uint8 HWC, normalised on the device (reference data/datasets/base.py:120-127).
    lf, rf, li, ri = synth.synthetic_frame(H, W, B=B, seed=1 + rank)
    mean = torch.tensor(ops.IMAGENET_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(ops.IMAGENET_STD).view(1, 3, 1, 1)
    to_u8 = lambda t: ((t * std + mean).clamp(0, 1) * 255).round().to(torch.uint8).permute(0, 2, 3, 1).contiguous()
    host_feats = [t.pin_memory() for t in lf + rf]
    host_imgs = [to_u8(li).pin_memory(), to_u8(ri).pin_memory()]
    h2d_bytes = sum(t.numel() * t.element_size() for t in host_feats + host_imgs)

    # ---- device-resident sets for `value` (inputs already in HBM when the timed region starts): NSETS sets rotated so
    #      that consecutive steps never re-read inputs from L2 (sets differ by a horizontal roll)
    in_bytes = sum(t.numel() * 4 for t in lf + rf + [li, ri])
    nsets = max(2, -(-2 * L2_BYTES // in_bytes))
    sets = []
    for s in range(nsets):
        feats = [torch.roll(t.to(dev), shifts=s, dims=-1).contiguous() for t in lf + rf]
        imgs = [ops.normalize_u8(torch.roll(t.to(dev), shifts=s, dims=2).contiguous()) for t in host_imgs]
        sets.append(feats + imgs)

    def forward(inp):
        return eng(inp[0:3], inp[3:6], inp[6], inp[7], {})

    def barrier():
        shard.barrier(dist)

    use_graph = not a.eager
    if use_graph:
        steps_dev = [CapturedStep(lambda inp=inp: forward(inp), device=dev) for inp in sets]
        run_dev = lambda i: steps_dev[i % nsets].replay()
        launches_per_step = steps_dev[0].launches
    else:
        run_dev = lambda i: forward(sets[i % nsets])
        launches_per_step = None

    # ---- device-resident throughput
    for i in range(max(a.warmup, 3)):
        run_dev(i)
    barrier()
    n0 = lib.tstereo_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        barrier()
        e0.record()
        for i in range(a.steps):
            run_dev(i)
        e1.record()
        barrier()
    launches = launches_per_step * a.steps if use_graph else lib.tstereo_launch_count() - n0
    # whole-job throughput: frames of all ranks / slowest rank's device time
    fps, ms, _ = shard.aggregate_throughput(B * a.steps, e0.elapsed_time(e1), dist, dev)

    # ---- end to end through the public call with HOST buffers.  Every step uploads its own inputs from pinned host
    #      memory (features fp32, images uint8), normalises the images on the device, runs the frame and downloads its
    #      full-resolution disparity.  Two staging sets: the upload of step i+1 (copy stream) and the download of step i-1
    #      (its own stream) overlap the compute of step i — the steady state of a streaming caller.
    stage_f = [[torch.empty_like(t, device=dev) for t in host_feats] for _ in range(2)]
    stage_i = [[torch.empty_like(t, device=dev) for t in host_imgs] for _ in range(2)]
    norm_i = [[torch.empty((B, 3, H, W), device=dev) for _ in range(2)] for _ in range(2)]
    full_host = [torch.empty((B, 1, H, W), dtype=torch.float32).pin_memory() for _ in range(2)]
    copy_stream, d2h_stream = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream(dev)
    uploaded = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    computed = [torch.cuda.Event() for _ in range(2)]
    downloaded = [torch.cuda.Event() for _ in range(2)]

    def frame_from_wire(k):
        ops.normalize_u8(stage_i[k][0], out=norm_i[k][0])
        ops.normalize_u8(stage_i[k][1], out=norm_i[k][1])
        return eng(stage_f[k][0:3], stage_f[k][3:6], norm_i[k][0], norm_i[k][1], {})

    if use_graph:
        steps_e2e = [CapturedStep(lambda k=k: frame_from_wire(k), device=dev) for k in range(2)]
        run_e2e = lambda k: steps_e2e[k].replay()
    else:
        run_e2e = frame_from_wire

    def upload(i):
        k = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[k])              # the set is free once the step that read it is done
            for d, h in zip(stage_f[k] + stage_i[k], host_feats + host_imgs):
                d.copy_(h, non_blocking=True)
            uploaded[k].record(copy_stream)

    def e2e_run(n):
        for ev in consumed + downloaded:
            ev.record(main_stream)
        upload(0)
        for i in range(n):
            k = i % 2
            if i + 1 < n:
                upload(i + 1)
            main_stream.wait_event(uploaded[k])
            main_stream.wait_event(downloaded[k])            # the graph's output buffer of set k was read by its last D2H
            o = run_e2e(k)
            consumed[k].record(main_stream)
            computed[k].record(main_stream)
            with torch.cuda.stream(d2h_stream):
                d2h_stream.wait_event(computed[k])
                full_host[k].copy_(o[0][0], non_blocking=True)
                downloaded[k].record(d2h_stream)
        main_stream.wait_stream(d2h_stream)

    e2e_run(3)
    barrier()
    e0.record()
    e2e_run(a.steps)
    e1.record()
    barrier()
    fps_e2e, ms_e2e, _ = shard.aggregate_throughput(B * a.steps, e0.elapsed_time(e1), dist, dev)

    # ---- what the upload alone can do on this box: every rank copies the same bytes from the same pinned buffers at the
    #      same time (nothing else running).  e2e cannot beat bytes / this rate; at N > 1 the ranks share the host's
    #      memory and PCIe fabric, which is what the N-GPU e2e scaling is bound by.
    def h2d_only(n):
        for _ in range(n):
            for d, h in zip(stage_f[0] + stage_i[0], host_feats + host_imgs):
                d.copy_(h, non_blocking=True)
    h2d_only(2)
    barrier()
    e0.record()
    h2d_only(10)
    e1.record()
    barrier()
    ms_h2d = shard.max_over_ranks(e0.elapsed_time(e1), dist, dev) / 10
    h2d_ceiling = h2d_bytes / (ms_h2d * 1e-3) / 1e9

    peak, peak_src = peaks()

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

    reps = max(10, min(a.steps, 50))
    result_extra = {}
    if rank == 0:
        # ---- roofline of the cost-volume path.  Candidates are the engine's own pattern: a piecewise-smooth disparity
        #      +- {4,1,0} (a trained network regresses smooth maps; random-init weights do not).
        def smooth_samples(h, w, S):
            yy, xx = torch.meshgrid(torch.arange(h, device=dev), torch.arange(w, device=dev), indexing="ij")
            base = 0.06 * w * (1.2 + torch.sin(xx / w * 6.0) * torch.cos(yy / h * 4.0)) + 0.3 * torch.rand(h, w, device=dev)
            offs = torch.tensor([-4.0, -1.0, 0.0, 1.0, 4.0], device=dev)[:S]
            return (base[None, None] + offs.view(1, S, 1, 1)).expand(B, S, h, w).contiguous()

        h4, w4 = H // 4, W // 4
        nrot = max(2, -(-2 * L2_BYTES // (8 * B * 128 * h4 * w4)))
        lcat = [torch.randn(B, 128, h4, w4, device=dev) for _ in range(nrot)]
        rcat = [torch.randn(B, 128, h4, w4, device=dev) for _ in range(nrot)]
        smp4, smp8 = smooth_samples(h4, w4, 5), smooth_samples(H // 8, W // 8, 5)
        first = eng._pk["precise.init3d.0.conv.0"]
        hs = eng.half_split

        # (1) the path the engine runs at the precise level (eng._first_conv): group terms, left-half conv, the 1x1
        #     "tap projection" of the right features (once per frame), the 3x3 conv over the group-wise channels, and the
        #     per-candidate gather / lerp kernel.  The raw volume never exists; algorithmic bytes = SURVEY.md 8d "fused
        #     path": features + candidates + conv output + weights.
        key = "precise.init3d"
        def fused(i):
            return eng._first_conv(lcat[i % nrot], rcat[i % nrot], smp4, key)
        gsteps = [CapturedStep(lambda i=i: fused(i), device=dev) for i in range(nrot)]
        ms_fused = timed(lambda i: gsteps[i % nrot].replay(), reps)
        sf = eng._sfmt() and eng.cost_form == "taps" and "taps" in first.tc
        part = {}
        if sf:
            sl0, sr0 = ops.split_pack(lcat[0]), ops.split_pack(rcat[0])
            sr5 = ops.Split(B, 128, 1, h4, w4, 2, t=sr0.t, five=True)
            g0 = ops.group_cost(lcat[0], rcat[0], smp4)
            al0, _ = ops.conv_hw3_s(sl0, first.tc["left"], None, 8, 1, None, half=1, oscale=first.osc)
            T0, _ = ops.conv_d_s(sr5, first.tc["taps"], None, 72, 1, 1, 1, False, None, half=1, oscale=first.tc["taps_osc"])
            T0 = T0.view(B, 72, h4, w4)
            gc0 = ops.conv_hw3_tc2(g0, first.tc["gconv"], None, 8, 1, None, half=True, oscale=first.osc)
            so0 = ops.Split(B, 8, 5, h4, w4, 2, device=dev)
            pieces = (("group_cost", lambda i: ops.group_cost(lcat[i % nrot], rcat[i % nrot], smp4)),
                      ("split_pack_LR", lambda i: (ops.split_pack(lcat[i % nrot], out=sl0), ops.split_pack(rcat[i % nrot], out=sr0))),
                      ("left_conv", lambda i: ops.conv_hw3_s(sl0, first.tc["left"], None, 8, 1, None, half=1, oscale=first.osc)),
                      ("tap_projection", lambda i: ops.conv_d_s(sr5, first.tc["taps"], None, 72, 1, 1, 1, False, None, half=1,
                                                                oscale=first.tc["taps_osc"])),
                      ("group_conv", lambda i: ops.conv_hw3_tc2(g0, first.tc["gconv"], None, 8, 1, None, half=True, oscale=first.osc)),
                      ("cost_taps", lambda i: ops.cost_taps(T0, smp4, gc0, al0, first.b, 8, "SiLU", sout=so0)))
            for name, fn in pieces:
                gs = CapturedStep(lambda fn=fn: [fn(i) for i in range(nrot)], device=dev)
                part[name] = timed(lambda i: gs.replay(), max(reps // nrot, 5)) / nrot
                del gs
            del sl0, sr0, sr5, T0, gc0, so0, al0
        # the producer form it replaced (ops.cost_conv_warp: the conv's producer warps rebuild the warped right features
        # per candidate), timed beside it
        g0 = ops.group_cost(lcat[0], rcat[0], smp4)
        alp = ops.conv_hw3_tc2(lcat[0], first.tc["left"], None, 8, 1, None, half=hs, oscale=first.osc)
        gs = CapturedStep(lambda: [ops.cost_conv_warp(rcat[i % nrot], smp4, g0, alp, first.tc["cost"], first.b, 8, "SiLU", half=hs,
                                                      oscale=first.osc) for i in range(nrot)], device=dev)
        part["producer_form_cost_conv (not on the default path)"] = timed(lambda i: gs.replay(), max(reps // nrot, 5)) / nrot
        del gs, g0, alp
        b_fused = 4 * B * (2 * 128 * h4 * w4 + 5 * h4 * w4 + 8 * 5 * h4 * w4) + 4 * 8 * 304 * 9
        # MACs actually issued: left 3x3 (128 -> 8) and tap projection (128 -> 72, 1x1) once per frame; group conv (48 -> 8, 3x3)
        # per candidate; the gather adds 2 * 9 * 8 fp32 FMAs per output position on the CUDA cores
        flops_fused = 2.0 * B * h4 * w4 * (128 * 9 * 8 + 128 * 72 + 5 * 48 * 9 * 8)
        del gsteps

        # (2) the materialising operator (ops.block_cost, the drop-in for the reference's block_cost): the kernel the
        #     north-star HBM target is quoted on; on the engine's path at the coarse level, and at every level with
        #     fuse_cost=False.
        def vol_bytes(C, h, w, S, planes, warp=True):
            return 4 * B * (2 * C * h * w + (S * h * w if warp else 0) + planes * S * h * w)
        b_p = vol_bytes(128, h4, w4, 5, 304)
        b_f = vol_bytes(128, H // 8, W // 8, 5, 304)
        b_c = vol_bytes(256, H // 16, W // 16, 12, 352, warp=False)
        ms_p = timed(lambda i: ops.block_cost(lcat[i % nrot], rcat[i % nrot], smp4), reps)
        ms_f = timed(lambda i: ops.block_cost(sets[i % nsets][1], sets[i % nsets][4], smp8), reps)
        ms_c = timed(lambda i: ops.block_cost(sets[i % nsets][2], sets[i % nsets][5], 12), reps)
        del lcat, rcat

        traffic = {}
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
            if int(tj.get("batch", -1)) == B and (H, W) == (544, 960):
                traffic = tj
        except (OSError, ValueError, KeyError):
            pass
        step_ms = ms / a.steps
        ach_fused = b_fused / (ms_fused * 1e-3) / 1e9
        result_extra["roofline"] = {
            "bound": "hbm", "kernel": "cost volume -> first conv at the precise level, as the engine runs it (group_cost + left-half "
            "conv + 1x1 tap projection of the right features + 3x3 conv over the group-wise channels + cost_taps gather/lerp): the raw "
            "198 MB/frame volume is never written",
            "achieved": ach_fused, "peak": peak, "unit": "GB/s", "frac": ach_fused / peak,
            "traffic": traffic.get("fused_precise_dram_bytes"), "traffic_source": traffic.get("source"),
            "peak_source": peak_src, "algorithmic_bytes": b_fused, "ms": ms_fused, "share_of_step": ms_fused / step_ms,
            "kernels_ms": part,
            "tensor_bound": {"flops_3term": 3 * flops_fused, "tflops_3term": 3 * flops_fused / (ms_fused * 1e-3) / 1e12,
                             "note": "tensor-core MACs actually issued (the channel contraction of the right half is hoisted out of the "
                                     "candidate loop), fp16 hi+lo split = 3 MMA terms per product; bound = max(bytes / HBM, flops / tensor peak)"}}
        result_extra["roofline_block_cost"] = {
            "bound": "hbm", "kernel": "block_cost_warp at the precise level (block_cost_main_kernel<1,1,0> + block_cost_resize_kernel): "
            "the materialising drop-in operator", "achieved": b_p / (ms_p * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
            "frac": b_p / (ms_p * 1e-3) / 1e9 / peak, "traffic": traffic.get("block_cost_precise_dram_bytes"),
            "algorithmic_bytes": b_p, "ms": ms_p,
            "all_three_levels": {"algorithmic_bytes": b_p + b_f + b_c, "ms": ms_p + ms_f + ms_c,
                                 "achieved": (b_p + b_f + b_c) / ((ms_p + ms_f + ms_c) * 1e-3) / 1e9,
                                 "frac": (b_p + b_f + b_c) / ((ms_p + ms_f + ms_c) * 1e-3) / 1e9 / peak}}

        # ---- the kernel most of the step runs on: the TMA-fed S-format tensor-core convolution, on its largest layer (UNet
        #      conv2.1, 32 -> 32 at 1/2 scale, both images of the B frames), S-format in and out.  HBM-bound: algorithmic
        #      bytes = 4 B per element in + out (fp16 hi + lo); ncu evidence (tensor pipe, shared-memory pipe, DRAM bytes) in
        #      profiles/r02_ncu_kernels.txt row 15.
        if eng._sfmt():
            c21 = eng._pk["precise.refinement.conv2.1"]
            h2, w2 = H // 2, W // 2
            nrc = max(2, -(-2 * L2_BYTES // (2 * B * 32 * h2 * w2 * 4)))
            xs_ = [ops.split_pack(torch.randn(2 * B, 32, h2, w2, device=dev)) for _ in range(nrc)]
            so_ = ops.Split(2 * B, 32, 1, h2, w2, 2, device=dev, five=False)
            gs = CapturedStep(lambda: [ops.conv_hw3_s(x_, c21.tc["hw3"], c21.b, 32, 1, "ReLU", half=1, oscale=c21.osc, sout=so_)
                                       for x_ in xs_], device=dev)
            ms_c = timed(lambda i: gs.replay(), max(reps // nrc, 5)) / nrc
            b_conv = 2 * (2 * B * 32 * h2 * w2 * 4)
            result_extra["roofline_conv"] = {
                "bound": "hbm", "kernel": "conv_tc2_kernel<32,2,DIRECT,RAW=3 (S-format in via TMA),F16,FOLD=3>: UNet conv2.1 32->32 3x3 at "
                "1/2 scale, 2B images, S-format in and out, 3-term fp16 split", "achieved": b_conv / (ms_c * 1e-3) / 1e9, "peak": peak,
                "unit": "GB/s", "frac": b_conv / (ms_c * 1e-3) / 1e9 / peak, "algorithmic_bytes": b_conv, "ms": ms_c,
                "traffic": traffic.get("split_conv_unet_conv21_dram_bytes"),
                "tflops_fp16_issued": 3 * 2.0 * 2 * B * h2 * w2 * 32 * 32 * 9 / (ms_c * 1e-3) / 1e12,
                "note": "ncu: tensor pipe 52 % active, shared-memory pipe ~96 % busy (tensor-core operand reads 61 % + TMA fills / "
                        "epilogue 35 %), DRAM 49 % of peak: the SWIZZLE_NONE K-major operand form reads A (4 KB) + B (3 KB) per 128x96x16 MMA"}
            del gs, xs_, so_

        # ---- streaming latency (B = 1, CUDA graph) and the temporal configuration C3 (KITTI 384x1248, pose warp on)
        if not a.no_extras:
            result_extra.update(extras(eng, dev, timed, CapturedStep, synth, temporal))

    result = {
        "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
        "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "frames_per_step_per_gpu": B,
                   "l2": f"inputs rotate over {nsets} device-resident sets ({nsets * in_bytes / 1e6:.0f} MB > 126 MB L2); "
                         "every intermediate cost volume alone exceeds L2",
                   "launch": "one CUDA-graph replay per step" if use_graph else "eager (one launch per kernel)",
                   "wire_format": "e2e uploads fp32 feature pyramids + uint8 HWC images (normalised on the device)",
                   "parallelism": f"batch-sharded replicas x{world}, no data-path collective",
                   "numa": f"rank pinned to {len(numa_cpus)} cores local to its GPU" if numa_cpus else "no NUMA pinning (NVML affinity unavailable)"},
        "clocks": clk.summary(),
        "e2e": {"value": fps_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": B * H * W * 4,
                "ms_per_step": ms_e2e / a.steps,
                "h2d_gbs_achieved": h2d_bytes / (ms_e2e / a.steps * 1e-3) / 1e9,
                "h2d_gbs_ceiling": h2d_ceiling, "h2d_ceiling_ms_per_step": ms_h2d,
                "h2d_note": f"ceiling = the same {h2d_bytes / 1e6:.0f} MB copied pinned->device by all {world} rank(s) at once with nothing "
                            "else running (max over ranks); e2e frames/s cannot exceed frames_per_step / h2d_ceiling_ms_per_step"},
        "gpu_launches": int(launches),
    }
    result.update(result_extra)
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cfps, med, n, thr = cpu_reference_frames(a, a.cpu_seconds)
        result["cpu_baseline"] = {"value": cfps, "unit": UNIT, "cores": thr, "kind": "port",
                                  "sample": f"median of {n} single frames (B=1) of the workload after 1 warm-up, "
                                            f"{med:.3f} s/frame, CPU oracle port of the reference forward"}
    if rank == 0:
        print(json.dumps(result))
    if dist is not None:
        dist.destroy_process_group()


def extras(eng, dev, timed, CapturedStep, synth, temporal):
    """Rank-0 side measurements beside the headline: B=1 streaming latency at C2 and the temporal configuration C3
    (KITTI 384x1248, D=192, pose warp on, steady state: cost memory + 3-plane local map), each as one CUDA-graph replay."""
    out = {}
    cu = lambda t: t.to(dev)
    lf, rf, li, ri = synth.synthetic_frame(544, 960, B=1, seed=3)
    inp = [cu(t) for t in lf + rf + [li, ri]]
    step = CapturedStep(lambda: eng(inp[0:3], inp[3:6], inp[6], inp[7], {}), device=dev)
    ms1 = timed(lambda i: step.replay(), 30)
    ms1_eager = timed(lambda i: eng(inp[0:3], inp[3:6], inp[6], inp[7], {}), 10)
    out["latency_b1"] = {"workload": "C2 544x960 D=192, B=1, single frame", "ms_per_frame_graph": ms1, "ms_per_frame_eager": ms1_eager,
                         "kernels": step.launches}
    del step
    H, W = 384, 1248
    lf, rf, li, ri = synth.synthetic_frame(H, W, B=1, seed=4)
    inp = [cu(t) for t in lf + rf + [li, ri]]
    st = synth.synthetic_temporal_state(H, W, B=1)
    pose = [cu(st[k]) for k in ("K", "T_now", "inv_T_prev", "baseline")]
    state = {}
    for _ in range(5):                      # the local map grows to 3 planes over the first frames
        if "prev_disp" in state:
            state = temporal.update_map(state, *pose, H, W, True, 3)
        state = eng(inp[0:3], inp[3:6], inp[6], inp[7], state)[5]
    frozen = {k: ({a_: b_.clone() for a_, b_ in v.items()} if isinstance(v, dict) else (v.clone() if torch.is_tensor(v) else v))
              for k, v in state.items()}
    copy = lambda s: {k: (dict(v) if isinstance(v, dict) else v) for k, v in s.items()}

    def frame():
        s = temporal.update_map(copy(frozen), *pose, H, W, True, 3)
        return eng(inp[0:3], inp[3:6], inp[6], inp[7], s)[0][0]
    step = CapturedStep(frame, device=dev)
    warp = CapturedStep(lambda: temporal.update_map(copy(frozen), *pose, H, W, True, 3)["local_map"], device=dev)
    out["temporal"] = {"workload": "C3 KITTI 384x1248 D=192, B=1, steady-state frame (update_map + aggregation, cost memory + local map)",
                       "ms_per_frame_graph": timed(lambda i: step.replay(), 30), "kernels_per_frame": step.launches,
                       "update_map_ms": timed(lambda i: warp.replay(), 30), "update_map_kernels": warp.launches,
                       "frames_per_s": None}
    out["temporal"]["frames_per_s"] = 1e3 / out["temporal"]["ms_per_frame_graph"]
    del step, warp

    # ---- BASELINE configs C4 / C5 at their stated shapes and per-GPU batch (B=4/GPU): a whole sequence,
    #      frame by frame with the state carried (update_map + aggregation), each distinct frame kind one CUDA-graph replay
    from temporalstereo_b200.aggregation import TEMPORALSTEREO
    for tag, (Hs, Ws, ns, Tn, Bs) in (("C4", (480, 640, 20, 5, 4)), ("C5", (1088, 1920, 16, 3, 4))):
        e2 = TEMPORALSTEREO(coarse=dict(num_sample=ns))
        e2.load_state_dict(synth.synthetic_state_dict(seed=0), strict=True)
        e2 = e2.to(dev).eval()
        lf, rf, li, ri = synth.synthetic_frame(Hs, Ws, B=Bs, seed=5)
        inp = [cu(t) for t in lf + rf + [li, ri]]
        st = synth.synthetic_temporal_state(Hs, Ws, B=Bs)
        pose = [cu(st[k]) for k in ("K", "T_now", "inv_T_prev", "baseline")]
        steps, state = [], {}
        for t in range(Tn):                  # frame kinds: no state / first warp / local map growing
            snap = {k: ({a_: b_.clone() for a_, b_ in v.items()} if isinstance(v, dict) else (v.clone() if torch.is_tensor(v) else v))
                    for k, v in state.items()}

            def frame(snap=snap, t=t):
                s = copy(snap)
                if t:
                    s = temporal.update_map(s, *pose, Hs, Ws, True, 3)
                return e2(inp[0:3], inp[3:6], inp[6], inp[7], s)
            state = frame()[5]
            steps.append(CapturedStep(lambda frame=frame: frame()[0][0], device=dev))
        ms_seq = timed(lambda i: [g.replay() for g in steps], 5)
        out["sequence_" + tag] = {"workload": f"{tag}: {Hs}x{Ws} D={16 * ns}, sequence T={Tn}, B={Bs}/GPU, pose warp on, state carried",
                                  "ms_per_sequence": ms_seq, "frames_per_s": Bs * Tn / (ms_seq * 1e-3),
                                  "kernels_per_sequence": sum(g.launches for g in steps)}
        del steps, e2, state
        torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    main()
