"""Per-step kernel timeline of the data-parallel train step (CUPTI via torch.profiler over CUDA-graph replays), rank 0:
every launch of the last profiled step with start offset, duration and stream - shows which ncclDevKernel is exposed
(nothing else in flight) and what runs beside NCCL.  Launch:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        scripts/dp_timeline.py [out.txt]
(N = 1 works too: the same timeline without collectives, for the diff.)"""
import collections
import os
import re
import sys
sys.path.insert(0, ".")
import torch
import torch.distributed as dist
from torch.profiler import profile, ProfilerActivity
from bench import _bench_args, PER_GPU_BATCH
from littlegan_b200 import model as M
from littlegan_b200.dataset import DevicePrefetcher, SyntheticCelebA
from littlegan_b200.eager_trainer import EagerTrainer

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
args = _bench_args(PER_GPU_BATCH)
M.set_init_seed(0)
dec, enc = M.Decoder(args), M.Encoder(args)
gen, disc = M.Generator(args, dec), M.Discriminator(args, enc)
adj = M.Adjuster(args, disc, gen)
data = SyntheticCelebA(args, batches=10 ** 9, seed=1 + rank, pool=4)
trainer = EagerTrainer(args, gen, disc, adj, data)
it = DevicePrefetcher(data.get_new_iterator(), depth=4)
for b in range(12, 17):
    trainer._train_step(b, it)
torch.cuda.synchronize()
graph = trainer._graphs[(True, None, True, True)][0]
steps = 4
if world > 1:
    dist.barrier()
# time first (no profiler), then profile
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    graph.replay()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(steps):
        graph.replay()
    torch.cuda.synchronize()
if rank == 0:
    seq = []
    for ev in prof.events():
        if ev.device_type != torch.autograd.DeviceType.CUDA:
            continue
        name = re.sub(r"void |\(anonymous namespace\)::|at::native::", "", ev.name)
        name = re.sub(r"\(.*", "", name)[:64]
        us = ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
        seq.append((ev.time_range.start, ev.time_range.end, name, us, getattr(ev, "device_resource_id", -1)))
    seq.sort()
    per = len(seq) // steps
    last = seq[-per:]
    t0 = last[0][0]
    lines = ["world %d  graph replay %.3f ms/step (CUDA events, 20 replays, no profiler)" % (world, ms),
             "last profiled step: %d launches, span %.1f us" % (len(last), last[-1][1] - t0)]
    # exposure of each NCCL kernel: time during which nothing else is in flight
    others = [(s, e) for s, e, n, _, _ in last if "nccl" not in n.lower()]
    for s, e, n, us, st in last:
        if "nccl" in n.lower():
            covered = 0.0
            cur = s
            for os_, oe in sorted(others):
                if oe <= cur or os_ >= e:
                    continue
                covered += min(oe, e) - max(os_, cur)
                cur = max(cur, min(oe, e))
            lines.append("NCCL  @%8.1f  %7.1f us  stream %s  alone for %.1f us   %s" % (s - t0, us, st, max(0.0, (e - s) - covered), n))
    lines.append("---- launches in order (> 6 us):  start  duration  stream  name")
    for s, e, n, us, st in last:
        if us > 6:
            lines.append("%8.1f %7.1f  %3s  %s" % (s - t0, us, st, n))
    txt = "\n".join(lines)
    print(txt)
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(txt + "\n")
if world > 1:
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)
