"""The bench's HBM-roofline kernels alone (bench.hbm_rooflines): python scripts/hbm_bench.py"""
import sys
sys.path.insert(0, ".")
import torch
import bench
from littlegan_b200 import model as M
from littlegan_b200.eager_trainer import EagerTrainer
args = bench._bench_args(bench.PER_GPU_BATCH)
M.set_init_seed(0)
dec, enc = M.Decoder(args), M.Encoder(args)
gen, disc = M.Generator(args, dec), M.Discriminator(args, enc)
adj = M.Adjuster(args, disc, gen)
tr = EagerTrainer(args, gen, disc, adj, None)
for r in bench.hbm_rooflines(bench._peaks(), tr):
    print("%-70s %6.0f GB/s  %.3f  %7.1f us" % (r["kernel"][:70], r["achieved"], r["frac"], r["ms_per_launch"] * 1e3))
