"""Per-kernel device time of one Inception pool_3 batch (torch.profiler / CUPTI), and per-unit timing of the conv units.
    python scripts/inception_profile.py [batch] [dtype] [out.txt]"""
import sys
from collections import defaultdict

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, ".")
from littlegan_b200.inception import InceptionPool3  # noqa: E402
from littlegan_b200 import kernels as K  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 100
dtype = sys.argv[2] if len(sys.argv) > 2 else "bf16"
out = open(sys.argv[3], "w") if len(sys.argv) > 3 else sys.stdout
net = InceptionPool3(seed=0, dtype=dtype)
img = torch.randint(0, 256, (B, 128, 128, 3), dtype=torch.uint8).cuda()
for _ in range(2):
    net(img)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    net(img)
    torch.cuda.synchronize()
tot = defaultdict(lambda: [0.0, 0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        tot[e.name][0] += e.device_time
        tot[e.name][1] += 1
total = sum(v[0] for v in tot.values())
print("# batch %d %s: kernel time %.2f ms" % (B, dtype, total / 1e3), file=out)
for name, (t, n) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    print("%6.2f%% %9.1f us %4d launches  %s" % (100 * t / total, t, n, name[:90]), file=out)

# per-unit timing (CUDA events around each conv unit, L2 warm)
orig = net._run_conv
rows = []


def timed(u, x, y=None, y_off=0):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = orig(u, x, y, y_off)
    e1.record()
    ho, wo = r.shape[1], r.shape[2]
    rows.append((u, x.shape[0] * ho * wo, e0, e1))
    return r


net._run_conv = timed
net(img)
torch.cuda.synchronize()
print("# per unit: name, M, K, N, us, TFLOP/s", file=out)
for u, M, e0, e1 in rows:
    us = e0.elapsed_time(e1) * 1e3
    Kd = u["k"][0] * u["k"][1] * u["cin"]
    print("%-28s M %7d K %5d N %4d  %8.1f us  %7.1f TF/s" % (u["name"], M, Kd, u["cout"], us,
                                                               2.0 * M * Kd * u["cout"] / us / 1e6), file=out)
