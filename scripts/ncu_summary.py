"""Condense an `ncu --page raw --csv` export into one line per launch (the columns we read)."""
import csv, sys
path = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else ""
rows = list(csv.reader(open(path)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "t"), ("launch__grid_size", "grid"), ("launch__block_size", "blk"),
        ("launch__registers_per_thread", "regs"), ("launch__occupancy_limit_registers", "occR"),
        ("launch__occupancy_limit_shared_mem", "occS"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("dram__bytes_read.sum", "rdMB"), ("dram__bytes_write.sum", "wrMB"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
        ("smsp__inst_executed.sum", "inst")]
have = [(c, n) for c, n in want if c in col]
print("idx " + " ".join(f"{n:>8}" for _, n in have) + "  kernel")
for i, r in enumerate(data):
    name = r[col["Kernel Name"]]
    if pat and pat not in name:
        continue
    out = []
    for c, n in have:
        v = r[col[c]].replace(",", "")
        u = units[col[c]]
        try:
            f = float(v)
            if n in ("rdMB", "wrMB"):
                f = f / {"byte": 1e6, "Kbyte": 1e3, "Mbyte": 1, "Gbyte": 1e-3}.get(u, 1e6)
            if n == "t":
                f = f / {"ns": 1e3, "us": 1, "ms": 1e-3}.get(u, 1e3)
            out.append(f"{f:8.1f}" if f < 1e6 else f"{f:8.2e}")
        except ValueError:
            out.append(f"{v:>8}")
    print(f"{i:3d} " + " ".join(out) + "  " + name.split("(")[0][:60])
