"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) of `bench.py --steps 1 --warmup 1` into the
per-kernel shares of the timed step: the step starts at the first launch of the LAST run of fine-propagator launches.
usage: python scripts/ncu_step_summary.py launches.csv out.csv"""
import csv, sys, collections
rows = []
with open(sys.argv[1]) as fh:
    lines = [l for l in fh if l.startswith('"')]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "s": 1e6, "second": 1e6}[unit]
        rows.append((int(r["ID"]), r["Kernel Name"], r["Grid Size"], r["Block Size"], v))
fine = [i for i, r in enumerate(rows) if "rk_fhn_tile_kernel<11" in r[1] or ("rk_pde_kernel<FhnPde, 11" in r[1])]
start = fine[-1]
while start - 1 in fine:  # the fine step is a run of consecutive chunk launches
    start -= 1
step = rows[start:]
agg = collections.OrderedDict()
for _, name, grid, block, us in step:
    short = name.split("(")[0].replace("void ", "")
    a = agg.setdefault(short, [0, 0.0])
    a[0] += 1; a[1] += us
tot = sum(a[1] for a in agg.values())
with open(sys.argv[2], "w", newline="") as fh:
    w = csv.writer(fh)
    w.writerow(["kernel", "launches", "total_us", "avg_us", "share_of_step"])
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        w.writerow([k, n, f"{us:.1f}", f"{us/n:.2f}", f"{us/tot:.4f}"])
    w.writerow(["TOTAL (cold-cache, serialised under ncu)", len(step), f"{tot:.1f}", "", "1.0"])
print(open(sys.argv[2]).read())
