"""Per-step deviation of the bf16 / fp32 loss trajectory from the oracle's golden fp32 trajectory (first 14 steps)."""
import json, os, sys
sys.path.insert(0, ".")
import numpy as np
from oracle import littlegan_oracle as O
from tests.test_train_step_gpu import _setup, _ListIterator, GOLD
gold = json.load(open(os.path.join(GOLD, "trajectory_full.json")))
g64 = json.load(open(os.path.join(GOLD, "trajectory_full_fp64.json")))
oargs = O.make_args(**gold["args"])
B = gold["args"]["batch_size"]
for dtype in ("fp32", "bf16"):
    pargs, gen, disc, adj, trainer, W = _setup(oargs, dtype, cuda_graph=True, seed=gold["seed"])
    row = []
    for b in range(1, 15):
        i1, c1, i2, c2, noise = O.synthetic_batch(oargs, B, seed=gold["data_seed"] + b)
        res = trainer._train_step(b, _ListIterator([(i1, c1), (i2, c2)]), noise=noise)
        row.append(max(abs(float(g) - w) / abs(w) for g, w in zip(res[3:5], (gold["gen"][b - 1], gold["disc"][b - 1]))))
    print(dtype, " ".join("%.4f" % v for v in row))
print("fp64-vs-fp32 oracle", " ".join("%.4f" % max(abs(a - b) / abs(b) for a, b in ((gold["gen"][i], g64["gen"][i]), (gold["disc"][i], g64["disc"][i]))) for i in range(14)))
