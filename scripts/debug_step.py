import sys, torch
sys.path.insert(0, ".")
from oracle import littlegan_oracle as O
from tests.util import *
from tests.test_train_step_gpu import _setup, _ListIterator
batch_no = int(sys.argv[1]) if len(sys.argv) > 1 else 3
oargs = small_args(use_partition=True)
if len(sys.argv) > 2 and sys.argv[2] == "full":
    oargs = O.make_args(cond_dim=40, batch_size=2, use_partition=False)
pargs, gen, disc, adj, trainer, W = _setup(oargs, "fp32")
ot = O.OracleTrainer(oargs, W, dtype=torch.float64)
i1, c1, i2, c2, noise = O.synthetic_batch(oargs, oargs.batch_size, seed=5)
ref = ot.train_step(batch_no, i1, c1, i2, c2, noise, return_grads=True)
res = trainer._train_step(batch_no, _ListIterator([(i1, c1), (i2, c2)]), noise=noise)
names = {"D": disc.weights, "G": gen.weights, "A": adj.weights[16:20]}
for key, ws in names.items():
    if ref["grads"][key] is None: continue
    for idx, gref in ref["grads"][key].items():
        got = ws[idx].lg_grad
        if key == "D": got = got.clamp(-0.5, 0.5)
        e = rel_err(got, gref)
        extra = ""
        if gref.numel() == 1: extra = " got %.6g ref %.6g raw %.6g" % (float(got), float(gref), float(ws[idx].lg_grad))
        print(key, idx, tuple(gref.shape), "%.3e" % e, extra)
