"""Top stall-sample instructions of one kernel from `ncu --page source --csv` (+ key raw metrics)."""
import csv, subprocess, sys
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); h, u, d = rows[0], rows[1], rows[2]
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"]
for k in keys:
    if k in h:
        print(f"{k:90s} {d[h.index(k)]} {u[h.index(k)]}")
for i, k in enumerate(h):
    if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and float(d[i]) > 0.2:
        print(f"  stall {k.split('issue_stalled_')[1].split('_per_')[0]:22s} {d[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines())); h = rows[1]; d = rows[2:]
ia, isamp, iex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
tot = sum(int(r[isamp] or 0) for r in d)
print("total samples", tot, "instructions", len(d))
for idx, r in sorted(enumerate(d), key=lambda t: -int(t[1][isamp] or 0))[:n]:
    print(f"{idx:5d} {int(r[isamp]):6d} {100*int(r[isamp])/tot:5.1f}% ex={r[iex]:>9s}  {r[ia][:100]}")
