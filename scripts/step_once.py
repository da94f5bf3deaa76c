"""Runs a few eager (non-graph) full train steps at the bench configuration - the command that
is profiled with `ncu --metrics gpu__time_duration.sum` to get the per-kernel launch list."""
import sys
sys.path.insert(0, ".")
import torch
from bench import _bench_args, PER_GPU_BATCH
from littlegan_b200 import model as M
from littlegan_b200.dataset import SyntheticCelebA
from littlegan_b200.eager_trainer import EagerTrainer

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
args = _bench_args(PER_GPU_BATCH)
args.cuda_graph = False
M.set_init_seed(0)
dec, enc = M.Decoder(args), M.Encoder(args)
gen, disc = M.Generator(args, dec), M.Discriminator(args, enc)
adj = M.Adjuster(args, disc, gen)
data = SyntheticCelebA(args, batches=10 ** 9, seed=1, pool=2)
trainer = EagerTrainer(args, gen, disc, adj, data)
it = data.get_new_iterator()
for b in range(12, 12 + steps):
    res = trainer._train_step(b, it)
torch.cuda.synchronize()
print("losses", float(res[3]), float(res[4]), float(res[5]))
