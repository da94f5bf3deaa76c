"""Throughput of the Inception-2015 pool_3 forward (config 5's feature pass): images/s at batch 100 (evaluate.py:54),
128x128 uint8 inputs from host memory and device-resident.  With `full`: the whole of config 5 on this GPU - 50 000
synthetic images -> features -> mean / covariance through fid.calculate_activation_statistics - and the CPU oracle
of the same forward on the host cores (a reported baseline only).
    python scripts/inception_bench.py [batch] [dtype] [full]"""
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

if len(sys.argv) > 3 and sys.argv[3] == "full":
    import os

    import numpy as np

    from littlegan_b200 import fid
    N = 50000
    rng = np.random.default_rng(0)
    imgs = rng.integers(0, 256, (N, 128, 128, 3), dtype=np.uint8)
    torch.cuda.synchronize()
    t = time.perf_counter()
    mu, sigma = fid.calculate_activation_statistics(imgs, net, batch_size=B, as_numpy=False)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    print("config 5 on 1 GPU: %d synthetic 128x128 images -> pool_3 -> mu / sigma in %.2f s = %.0f img/s (%s, host uint8 "
          "input, batch %d); sigma trace %.3f" % (N, dt, N / dt, dtype, B, float(sigma.trace())))
    from oracle import inception_oracle as IO
    torch.set_num_threads(os.cpu_count())
    ora = IO.InceptionOracle(IO.random_weights(0), dtype=torch.float32)
    x = torch.from_numpy(imgs[:16]).float()
    ora(x[:2])
    t = time.perf_counter()
    ora(x)
    dt = time.perf_counter() - t
    print("CPU oracle (PyTorch-CPU fp32, %d threads): 16 images in %.2f s = %.1f img/s" % (os.cpu_count(), dt, 16 / dt))
