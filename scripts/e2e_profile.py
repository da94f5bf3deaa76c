"""Host-side profile of the end-to-end step loop (public API, host batches, loss read-back every step).
usage: python scripts/e2e_profile.py [steps]"""
import cProfile
import pstats
import sys
import time
sys.path.insert(0, ".")
import torch
from bench import _bench_args, PER_GPU_BATCH
from littlegan_b200 import model as M
from littlegan_b200.dataset import DevicePrefetcher, SyntheticCelebA
from littlegan_b200.eager_trainer import EagerTrainer

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
args = _bench_args(PER_GPU_BATCH)
M.set_init_seed(0)
dec, enc = M.Decoder(args), M.Encoder(args)
gen, disc = M.Generator(args, dec), M.Discriminator(args, enc)
adj = M.Adjuster(args, disc, gen)
data = SyntheticCelebA(args, batches=10 ** 9, seed=1, pool=8)
trainer = EagerTrainer(args, gen, disc, adj, data)
it = DevicePrefetcher(data.get_new_iterator(), depth=4)
b = 11
for _ in range(5):
    b += 1
    trainer._train_step(b, it)
torch.cuda.synchronize()


def loop(n, read):
    global b
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    prev = None
    for _ in range(n):
        b += 1
        res = trainer._train_step(b, it)
        if read == "sync":
            losses = (float(res[3]), float(res[4]), float(res[5]))
        elif read == "pipelined" and prev is not None:
            losses = (float(prev[3]), float(prev[4]), float(prev[5]))
        prev = res
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3 / n


print("e2e ms/step, losses of step i read right after step i: %.3f" % loop(steps, "sync"))
print("e2e ms/step, losses of step i read after submitting step i+1: %.3f" % loop(steps, "pipelined"))
print("e2e ms/step, no read-back (host runs ahead): %.3f" % loop(steps, None))
pr = cProfile.Profile()
pr.enable()
loop(steps, "pipelined")
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
