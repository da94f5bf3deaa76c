"""Config 4 of BASELINE.json: EagerTrainer.predict (eager_trainer.py:265-298: 1 G + 2 D + 2 A forwards and the
four MSE scalars), batch 256, through the public call with HOST inputs (numpy), wall clock per call.
usage: python scripts/predict_bench.py [batch] [reps]"""
import sys
import time
sys.path.insert(0, ".")
import numpy as np
import torch
from bench import _bench_args, E_MAC, DC_MAC
from littlegan_b200 import model as M
from littlegan_b200.eager_trainer import EagerTrainer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
args = _bench_args(B)
M.set_init_seed(0)
dec, enc = M.Decoder(args), M.Encoder(args)
gen, disc = M.Generator(args, dec), M.Discriminator(args, enc)
adj = M.Adjuster(args, disc, gen)
trainer = EagerTrainer(args, gen, disc, adj, None)
rng = np.random.default_rng(0)
noise = rng.standard_normal((B, args.noise_dim)).astype(np.float32)
cond = (0.96 * rng.choice([-1.0, 1.0], size=(B, args.cond_dim)) + 0.02).astype(np.float32)
image = rng.uniform(-1, 1, size=(B, 128, 128, 3)).astype(np.float32)
for _ in range(3):
    trainer.predict(noise, cond, image)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(reps):
    out = trainer.predict(noise, cond, image)
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) * 1e3 / reps
flop = 2 * (4 * E_MAC + 3 * DC_MAC) * B
print("predict batch %d, host numpy inputs : %.2f ms per call, %.0f images/sec, %.1f TFLOP/s of the reference's conv "
      "work (1 G + 2 D + 2 A passes)" % (B, ms, B / ms * 1e3, flop / ms / 1e9))
dn, dc, di = torch.from_numpy(noise).cuda(), torch.from_numpy(cond).cuda(), torch.from_numpy(image).cuda()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(reps):
    out = trainer.predict(dn, dc, di)
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) * 1e3 / reps
print("predict batch %d, device inputs     : %.2f ms per call, %.0f images/sec, %.1f TFLOP/s" % (
    B, ms, B / ms * 1e3, flop / ms / 1e9))
