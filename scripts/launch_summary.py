"""Summarises an ncu `--metrics gpu__time_duration.sum --csv` launch list: per kernel name,
launch count, total and share of the device time (last `--steps` fraction of the run)."""
import csv, sys, collections, re
path = sys.argv[1]
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5     # keep the last fraction (skip warm-up step)
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
    rows.append((r["Kernel Name"], r.get("Grid Size", ""), ns))
rows = rows[int(len(rows) * (1 - frac)):]
tot = sum(r[2] for r in rows)
agg = collections.OrderedDict()
for name, grid, ns in rows:
    short = re.sub(r"\(.*", "", name)
    short = re.sub(r"void |\(anonymous namespace\)::|at::native::", "", short)[:90]
    a = agg.setdefault(short, [0, 0.0])
    a[0] += 1; a[1] += ns
print("launches %d, total %.3f ms" % (len(rows), tot / 1e6))
for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%6.2f%% %9.3f ms %5d  %s" % (100 * ns / tot, ns / 1e6, n, k))
