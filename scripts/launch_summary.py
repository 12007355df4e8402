"""ncu `--metrics gpu__time_duration.sum --csv` launch list -> markdown summary (per kernel family + per launch)."""
import collections, csv, sys
src, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "launch list")
lines = [l for l in open(src) if not l.startswith("==")]
seq, agg, tot = [], collections.OrderedDict(), 0.0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    t = float(row["Metric Value"].replace(",", ""))
    t = t / 1e3 if row["Metric Unit"] == "ns" else (t * 1e3 if row["Metric Unit"] == "ms" else t)
    name = row["Kernel Name"].split("(")[0].replace("void ", "").replace("tstereo::", "")
    seq.append((name, t, row["Grid Size"], row["Block Size"]))
    tot += t
    a = agg.setdefault(name, [0.0, 0])
    a[0] += t
    a[1] += 1
print(f"# {title}\n\n{len(seq)} launches, {tot:.1f} us of serialised (cold-cache, `--clock-control none`) kernel time.\n")
print("| kernel | launches | us | share |\n|---|---|---|---|")
for k, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"| `{k}` | {n} | {t:.1f} | {100 * t / tot:.1f} % |")
print("\n<details><summary>every launch in order</summary>\n\n| # | us | grid | block | kernel |\n|---|---|---|---|---|")
for i, (n, t, g, b) in enumerate(seq):
    print(f"| {i} | {t:.1f} | {g} | {b} | `{n}` |")
print("\n</details>")
