"""Where does the end-to-end step lose time against the graph replay alone?  CUPTI timeline of the public-API
loop (host batches through DevicePrefetcher, pipelined loss read-back): GPU-busy union vs wall, and the work that
runs outside the captured graph.    usage: python scripts/e2e_gaps.py [steps]"""
import collections
import re
import sys
sys.path.insert(0, ".")
import torch
from torch.profiler import profile, ProfilerActivity
from bench import _bench_args, PER_GPU_BATCH
from littlegan_b200 import model as M
from littlegan_b200.dataset import DevicePrefetcher, SyntheticCelebA
from littlegan_b200.eager_trainer import EagerTrainer

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
args = _bench_args(PER_GPU_BATCH)
M.set_init_seed(0)
dec, enc = M.Decoder(args), M.Encoder(args)
gen, disc = M.Generator(args, dec), M.Discriminator(args, enc)
adj = M.Adjuster(args, disc, gen)
data = SyntheticCelebA(args, batches=10 ** 9, seed=1, pool=8)
trainer = EagerTrainer(args, gen, disc, adj, data)
it = DevicePrefetcher(data.get_new_iterator(), depth=4)
b = 11
for _ in range(6):
    b += 1
    trainer._train_step(b, it)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    prev = None
    for _ in range(steps):
        b += 1
        res = trainer._train_step(b, it)
        if prev is not None:
            float(prev[3]); float(prev[4]); float(prev[5])
        prev = res
    float(prev[3])
    torch.cuda.synchronize()
evs = []
for ev in prof.events():
    if ev.device_type != torch.autograd.DeviceType.CUDA:
        continue
    evs.append((ev.time_range.start, ev.time_range.end, ev.name))
evs.sort()
t0, t1 = evs[0][0], max(e[1] for e in evs)
# union of busy intervals, H2D copies excluded (they run on the copy engine beside the kernels)
busy, cur_s, cur_e = 0.0, None, None
for s, e, n in evs:
    if "Memcpy HtoD" in n:
        continue
    if cur_e is None or s > cur_e:
        if cur_e is not None:
            busy += cur_e - cur_s
        cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
busy += cur_e - cur_s
print("wall %.1f us for %d steps = %.3f ms/step; SM-busy union %.3f ms/step; idle %.3f ms/step" % (
    t1 - t0, steps, (t1 - t0) / steps / 1e3, busy / steps / 1e3, (t1 - t0 - busy) / steps / 1e3))
agg = collections.Counter()
for s, e, n in evs:
    if any(k in n for k in ("Memcpy", "Memset", "elementwise", "distribution", "copy")):
        agg[re.sub(r"\(.*", "", n)[:60]] += e - s
for k, v in agg.most_common(8):
    print("%9.1f us/step  %s" % (v / steps, k))
