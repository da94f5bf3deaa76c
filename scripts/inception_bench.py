"""Throughput of the Inception-2015 pool_3 forward (config 5's feature pass): images/s at batch 100 (evaluate.py:54),
128x128 uint8 inputs from host memory and device-resident, plus the per-kernel share.
    python scripts/inception_bench.py [batch] [dtype]"""
import sys
import time

import torch

sys.path.insert(0, ".")
from littlegan_b200.inception import InceptionPool3  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 100
dtype = sys.argv[2] if len(sys.argv) > 2 else "fp32"
net = InceptionPool3(seed=0, dtype=dtype)
img = torch.randint(0, 256, (B, 128, 128, 3), dtype=torch.uint8).pin_memory()
dimg = img.cuda()
for _ in range(2):
    net(dimg)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 5
e0.record()
for _ in range(reps):
    net(dimg)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
flops = 2 * 5.713e9 * B
print("inception pool_3 %s batch %d: %.2f ms/batch device-resident = %.0f img/s, %.1f TFLOP/s" % (
    dtype, B, ms, B / ms * 1e3, flops / ms / 1e9))
t = time.perf_counter()
for _ in range(reps):
    f = net(img)
f.cpu()
dt = (time.perf_counter() - t) / reps
print("from pinned host bytes incl. feature read-back: %.2f ms/batch = %.0f img/s" % (dt * 1e3, B / dt))
