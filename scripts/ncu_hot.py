"""Top stall-sample SASS instructions of an ncu report (source page).  usage: ncu_hot.py file.ncu-rep [n]"""
import csv, subprocess, sys
rep = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(lines[start:]))
tot = sum(int(r["# Samples"] or 0) for r in rows)
print("total samples", tot, "instructions", len(rows))
idx = sorted(range(len(rows)), key=lambda i: -int(rows[i]["# Samples"] or 0))[:n]
stalls = [k for k in rows[0] if k.startswith("stall_") and "Not Issued" not in k]
for i in sorted(idx):
    r = rows[i]
    top = sorted(((int(r[k] or 0), k) for k in stalls), reverse=True)[:2]
    print("%5d %6.2f%% ex=%-8s %-60s %s" % (i, 100.0 * int(r["# Samples"]) / tot, r["Instructions Executed"], r["Source"].strip()[:60],
                                       " ".join("%s:%d" % (k[6:], v) for v, k in top if v)))
