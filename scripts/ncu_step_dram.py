"""Post-processes an ncu CSV launch list (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum)
of `bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline`: takes the LAST `n` launches (= one replay of the
captured step; n = launches per capture, printed by bench as gpu_launches/steps) and prints per kernel: launches,
device time (cold, serialised), DRAM bytes read / written, achieved GB/s, and the step totals.
usage: python scripts/ncu_step_dram.py launches.csv first_id last_id [out.txt]
(the replay to sum = the launch IDs from one pack_weights_multi_kernel to the next; bench.py --steps 1 --warmup 3 runs
eager step, replay, replay, the TIMED replay, then one end-to-end step)"""
import collections
import csv
import re
import sys

path, first, last_id = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = collections.OrderedDict()
for r in csv.DictReader(lines):
    key = r["ID"]
    d = rows.setdefault(key, {"name": r["Kernel Name"]})
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "")
    m = r["Metric Name"]
    if m == "gpu__time_duration.sum":
        d["ns"] = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
    else:
        mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        d["rd" if "read" in m else "wr"] = v * mul
last = [d for k, d in rows.items() if first <= int(k) <= last_id]
agg = collections.OrderedDict()
for d in last:
    short = re.sub(r"\(.*", "", d["name"])
    short = re.sub(r"void |\(anonymous namespace\)::|at::native::", "", short)[:70]
    a = agg.setdefault(short, [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += d.get("ns", 0); a[2] += d.get("rd", 0); a[3] += d.get("wr", 0)
tot_ns = sum(a[1] for a in agg.values()); tot_rd = sum(a[2] for a in agg.values()); tot_wr = sum(a[3] for a in agg.values())
out = ["%d launches of one step replay: device time %.3f ms (cold-cache, serialised), DRAM read %.1f MB + write %.1f MB = %.1f MB"
       % (len(last), tot_ns / 1e6, tot_rd / 1e6, tot_wr / 1e6, (tot_rd + tot_wr) / 1e6),
       "%7s %9s %5s %10s %10s %8s  %s" % ("share", "time ms", "n", "read MB", "write MB", "GB/s", "kernel")]
for k, (cnt, ns, rd, wr) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append("%6.2f%% %9.3f %5d %10.1f %10.1f %8.0f  %s" % (100 * ns / tot_ns, ns / 1e6, cnt, rd / 1e6, wr / 1e6,
                                                           (rd + wr) / ns if ns else 0, k))
txt = "\n".join(out)
print(txt)
if len(sys.argv) > 4:
    open(sys.argv[4], "w").write(txt + "\n")
