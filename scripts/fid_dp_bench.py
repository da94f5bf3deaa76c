"""Config 5 sharded over the GPUs of one box: every rank pushes its slice of the 50 000 synthetic 128x128 images through
the Inception pool_3 forward and the streaming (n, S1, S2) accumulation; ONE sum all-reduce of the statistics at the
end (fid.FeatureStatistics.finalize), then mu / sigma.  Launch:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 \
        scripts/fid_dp_bench.py [total_images] [dtype]"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from littlegan_b200 import fid  # noqa: E402
from littlegan_b200.inception import InceptionPool3  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
total = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
dtype = sys.argv[2] if len(sys.argv) > 2 else "bf16"
per = total // world
imgs = np.random.default_rng(rank).integers(0, 256, (per, 128, 128, 3), dtype=np.uint8)
net = InceptionPool3(seed=0, dtype=dtype)
fid.calculate_activation_statistics(imgs[:400], net, batch_size=100, as_numpy=False)      # warm-up (incl. NCCL)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t = time.perf_counter()
e0.record()
mu, sigma = fid.calculate_activation_statistics(imgs, net, batch_size=100, as_numpy=False)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dist.barrier()
wall = time.perf_counter() - t
if rank == 0:
    print("config 5 on %d GPU(s): %d images (%d per rank) -> pool_3 -> all-reduced mu / sigma: %.2f s device time (max over "
          "ranks), %.2f s wall = %.0f img/s (%s); trace(sigma) %.4f" % (world, per * world, per, float(ms) / 1e3, wall,
                                                                     per * world / (float(ms) / 1e3), dtype,
                                                                     float(sigma.trace())))
if world > 1:
    dist.destroy_process_group()
