"""Warm per-kernel device times of the full train step (torch.profiler / CUPTI over eager steps at the
bench configuration): unlike the ncu launch list nothing is serialised or cache-flushed.
usage: python scripts/step_profile.py [steps] [out.txt] [graph]     (graph: profile the CUDA-graph replay)"""
import collections
import re
import sys
sys.path.insert(0, ".")
import torch
from torch.profiler import profile, ProfilerActivity
from bench import _bench_args, PER_GPU_BATCH
from littlegan_b200 import model as M
from littlegan_b200.dataset import SyntheticCelebA
from littlegan_b200.eager_trainer import EagerTrainer

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
args = _bench_args(PER_GPU_BATCH)
args.cuda_graph = len(sys.argv) > 3 and sys.argv[3] == "graph"
M.set_init_seed(0)
dec, enc = M.Decoder(args), M.Encoder(args)
gen, disc = M.Generator(args, dec), M.Discriminator(args, enc)
adj = M.Adjuster(args, disc, gen)
data = SyntheticCelebA(args, batches=10 ** 9, seed=1, pool=2)
trainer = EagerTrainer(args, gen, disc, adj, data)
it = data.get_new_iterator()
for b in range(12, 15):
    trainer._train_step(b, it)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for b in range(15, 15 + steps):
        trainer._train_step(b, it)
    torch.cuda.synchronize()
agg = collections.OrderedDict()
seq = []
tot = 0.0
for ev in prof.events():
    if ev.device_type != torch.autograd.DeviceType.CUDA:
        continue
    name = re.sub(r"void |\(anonymous namespace\)::|at::native::", "", ev.name)
    name = re.sub(r"\(.*", "", name)[:70]
    us = ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += us; tot += us
    seq.append((ev.time_range.start, name, us, ev.time_range.end))
lines = ["steps %d  kernel time per step %.3f ms" % (steps, tot / steps / 1e3)]
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append("%6.2f%% %9.1f us/step %6.1f launches/step  %8.1f us avg  %s" % (100 * us / tot, us / steps, n / steps, us / n, k))
seq.sort()
lines.append("---- launches of the last step in order (> 8 us)")
last = seq[-len(seq) // steps:]
t0 = last[0][0]
for st, name, us, _ in last:
    if us > 8:
        lines.append("%8.1f us  @%8.1f  %s" % (us, st - t0, name))
# concurrency of the last step: how much of the wall span has 1, 2, 3+ kernels in flight (the step runs its
# independent chains on parallel streams)
evs = sorted([(st, 1) for st, _, _, en in last] + [(en, -1) for st, _, _, en in last])
depth, prev, hist = 0, evs[0][0], collections.Counter()
for t, d in evs:
    hist[min(depth, 4)] += t - prev
    prev = t
    depth += d
span = evs[-1][0] - evs[0][0]
lines.append("---- last step: wall span %.1f us, sum of kernel durations %.1f us" % (span, sum(us for _, _, us, _ in last)))
lines.append("     time with k kernels in flight: " + "  ".join("k=%d%s: %.1f us" % (k, "+" if k == 4 else "", v) for k, v in sorted(hist.items())))
txt = "\n".join(lines)
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt + "\n")
