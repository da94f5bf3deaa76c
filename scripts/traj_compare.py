"""Per-step deviation of the GPU loss trajectory from the committed oracle trajectory."""
import json, sys, torch
sys.path.insert(0, ".")
from oracle import littlegan_oracle as O
from tests.test_train_step_gpu import _setup, _ListIterator
gold = json.load(open("tests/golden/trajectory_full.json"))
oargs = O.make_args(**gold["args"])
for dtype in sys.argv[1:] or ["fp32", "bf16"]:
    pargs, gen, disc, adj, trainer, W = _setup(oargs, dtype, cuda_graph=True, seed=gold["seed"])
    B = gold["args"]["batch_size"]
    worst = [0, 0, 0]
    for b in range(1, len(gold["gen"]) + 1):
        i1, c1, i2, c2, noise = O.synthetic_batch(oargs, B, seed=gold["data_seed"] + b)
        res = trainer._train_step(b, _ListIterator([(i1, c1), (i2, c2)]), noise=noise)
        want = (gold["gen"][b - 1], gold["disc"][b - 1], gold["adj"][b - 1])
        devs = []
        for k, (got, w) in enumerate(zip(res[3:6], want)):
            if w is None: devs.append(None); continue
            d = abs(float(got) - w) / abs(w); devs.append(d); worst[k] = max(worst[k], d)
        print(dtype, b, " ".join("%.4f" % w if w is not None else "None" for w in want), "|", " ".join("%.2e" % d if d is not None else "None" for d in devs), flush=True)
    print(dtype, "WORST", worst)
