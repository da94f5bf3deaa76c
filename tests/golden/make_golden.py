"""Generates the committed golden fixtures from the CPU oracle (run in the build container):

    python tests/golden/make_golden.py [--steps 100]

trajectory_full_fp64.json : the same run with the oracle in fp64 (the fp32-vs-fp64 gap is the noise
                       floor of this chaotic, Adam-normalised system).
trajectory_full.json : 100-step G/D/A loss trajectory of the real architecture (cond 40, batch 4),
                       initial weights = the product builders' seeded Glorot init (seed 0),
                       inputs = oracle.synthetic_batch(seed = data_seed + step).
small_step.npz       : one full step (batch_no 11) of the reduced architecture: outputs, losses and
                       per-tensor gradient / updated-weight checksums.
The reference itself (TF 1.15) cannot run here - these vectors pin the ORACLE, not the reference.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import littlegan_oracle as O  # noqa: E402
from tests.util import build_product, product_args, product_weights_to_oracle, small_args  # noqa: E402


def trajectory(steps, dtype=torch.float32, name="trajectory_full.json"):
    cfg = dict(cond_dim=40, batch_size=4, use_partition=True)
    oargs = O.make_args(**cfg)
    gen, disc, adj = build_product(product_args(oargs, dtype="fp32"), seed=0)
    ot = O.OracleTrainer(oargs, product_weights_to_oracle(gen, disc, adj), dtype=dtype)
    out = dict(args=cfg, seed=0, data_seed=1000, gen=[], disc=[], adj=[])
    for b in range(1, steps + 1):
        r = ot.train_step(b, *O.synthetic_batch(oargs, cfg["batch_size"], seed=1000 + b))
        out["gen"].append(float(r["gen_loss"]))
        out["disc"].append(float(r["disc_loss"]))
        out["adj"].append(None if r["adj_loss"] is None else float(r["adj_loss"]))
        print(b, out["gen"][-1], out["disc"][-1], out["adj"][-1], flush=True)
    with open(os.path.join(HERE, name), "w") as f:
        json.dump(out, f)


def small_step():
    oargs = small_args(use_partition=True)
    gen, disc, adj = build_product(product_args(oargs, dtype="fp32"), seed=0)
    ot = O.OracleTrainer(oargs, product_weights_to_oracle(gen, disc, adj), dtype=torch.float64)
    r = ot.train_step(11, *O.synthetic_batch(oargs, 4, seed=5), return_grads=True)
    save = dict(fake_image=r["fake_image"].numpy().astype(np.float32),
                adj_image=r["adj_image"].numpy().astype(np.float32),
                losses=np.array([float(r["gen_loss"]), float(r["disc_loss"]), float(r["adj_loss"])]))
    for key in "DGA":
        save["grad_sum_" + key] = np.array([float(g.sum()) for g in r["grads"][key].values()])
        save["grad_abs_" + key] = np.array([float(g.abs().sum()) for g in r["grads"][key].values()])
        save["w_sum_" + key] = np.array([float(w.detach().sum()) for w in ot.W[key]])
    np.savez_compressed(os.path.join(HERE, "small_step.npz"), **save)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--fp64", action="store_true", help="also write the fp64 oracle trajectory (noise floor)")
    a = ap.parse_args()
    small_step()
    trajectory(a.steps)
    if a.fp64:
        trajectory(a.steps, torch.float64, "trajectory_full_fp64.json")
