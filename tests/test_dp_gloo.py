"""World-size-2 `gloo` tests (CPU) of the data-parallel protocol the product uses on NCCL:
  - train step: each rank takes its slice of the global batch, gradients are all-reduced with
    AVERAGE over the flat per-optimiser ranges, the D value-clip is applied AFTER the reduce
    (eager_trainer.py `_step_body`), which must equal the single-process global-batch step;
  - FID: (n, S1, S2) partial sums with a rank-0-broadcast shift are SUM-reduced (fid.py
    `FeatureStatistics.finalize`), which must equal np.mean / np.cov of the concatenated shards.
The arithmetic on each rank is the CPU oracle (the checker); what is under test is the PRODUCT's
host code: the flat gradient arena and its lg_grad views, `_range` / `_bucket` (partition groups,
the generator's two buckets) and `EagerTrainer._reduce_async` itself, called on CPU arenas over
gloo exactly as `_step_body` calls it over NCCL."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import fid_oracle
from oracle import littlegan_oracle as O
from tests.util import build_product, product_args, product_weights_to_oracle, small_args


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        from littlegan_b200.eager_trainer import EagerTrainer, _dist
        assert _dist() is not None and _dist().get_world_size() == world
        oargs = small_args(use_partition=True, batch_size=2)          # per-rank batch 2, global 4
        pargs = product_args(oargs, dtype="fp32")
        gen, disc, adj = build_product(pargs, seed=0)
        W = product_weights_to_oracle(gen, disc, adj)
        trainer = EagerTrainer(pargs, gen, disc, adj, None)          # host arenas (CPU tensors): bookkeeping only
        i1, c1, i2, c2, noise = O.synthetic_batch(small_args(batch_size=4), 4, seed=9)
        sl = slice(rank * 2, rank * 2 + 2)
        batch_no = 15                                                 # partition step: group 0 of D and G
        ot = O.OracleTrainer(oargs, W, dtype=torch.float64)
        ot.args.use_clip = False                                      # local grads unclipped; clip after reduce
        r = ot.train_step(batch_no, i1[sl], c1[sl], i2[sl], c2[sl], noise[sl], return_grads=True)
        # this rank's local gradients go where the kernels would have put them - the lg_grad views into the flat
        # gradient arena - and everything outside the active group holds a per-rank marker that must survive
        trainer.Gd.fill_(100.0 + rank)
        for key, ws in (("D", disc.weights), ("G", gen.weights), ("A", adj.weights[16:20])):
            for i, g in r["grads"][key].items():
                ws[i].lg_grad.copy_(g)
        # the PRODUCT's reduction calls, exactly as _step_body / _adjuster_chain issue them
        trainer._reduce_async("Adjuster", batch_no)
        trainer._reduce_async("Discriminator", batch_no)
        trainer._reduce_async("Generator", batch_no, part=(4, 22))
        trainer._reduce_async("Generator", batch_no, part=(0, 4))
        res = {}
        for name, key, ws in (("Discriminator", "D", disc.weights), ("Generator", "G", gen.weights),
                              ("Adjuster", "A", adj.weights[16:20])):
            lo, hi = trainer._range(name, batch_no)
            grads = {i: ws[i].lg_grad.double().clone() for i in r["grads"][key]}
            if key == "D":
                grads = {i: g.clamp(-0.5, 0.5) for i, g in grads.items()}      # the value-clip follows the reduce
            offs = trainer._offsets[name]
            outside = torch.cat([trainer.Gd[offs[0]:lo], trainer.Gd[hi:offs[-1]]])
            res[key] = (grads, bool((outside == 100.0 + rank).all()), (lo - offs[0], hi - offs[0]))
        loss = torch.tensor([float(r["gen_loss"]), float(r["disc_loss"]), float(r["adj_loss"])], dtype=torch.float64)
        dist.all_reduce(loss)
        loss /= world

        # FID partial sums
        rng = np.random.RandomState(5)
        feats = rng.randn(400, 24) @ rng.randn(24, 24) + 1.5
        mine = feats[rank * 200:(rank + 1) * 200]
        shift = torch.from_numpy(mine.mean(0))
        dist.broadcast(shift, src=0)
        c = mine - shift.numpy()
        S1, S2 = torch.from_numpy(c.sum(0)), torch.from_numpy(c.T @ c)
        n = torch.tensor([200.0], dtype=torch.float64)
        for t in (S1, S2, n):
            dist.all_reduce(t)
        if rank == 0:
            q.put(dict(res={k: ({i: g.numpy() for i, g in v[0].items()}, v[1], v[2]) for k, v in res.items()},
                       loss=loss.numpy(), S1=S1.numpy(), S2=S2.numpy(), n=float(n), shift=shift.numpy(),
                       feats=feats))
    finally:
        dist.destroy_process_group()


def test_data_parallel_protocol_world2():
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0

    # single-process, global batch 4
    oargs = small_args(use_partition=True, batch_size=4)
    gen, disc, adj = build_product(product_args(oargs, dtype="fp32"), seed=0)
    ot = O.OracleTrainer(oargs, product_weights_to_oracle(gen, disc, adj), dtype=torch.float64)
    i1, c1, i2, c2, noise = O.synthetic_batch(oargs, 4, seed=9)
    ref = ot.train_step(15, i1, c1, i2, c2, noise, return_grads=True)
    assert abs(out["loss"][0] - float(ref["gen_loss"])) < 1e-12
    assert abs(out["loss"][1] - float(ref["disc_loss"])) < 1e-12
    assert abs(out["loss"][2] - float(ref["adj_loss"])) < 1e-12
    for key in "DGA":
        grads, untouched, rng = out["res"][key]
        assert untouched, key                                        # nothing outside the active range was reduced
        assert sorted(ref["grads"][key]) == sorted(grads)            # same partition group on every rank
        for i, got in grads.items():
            want = ref["grads"][key][i].numpy()                      # oracle: clip(mean-gradient)
            # the arenas are fp32: the averaged fp64 oracle gradients were rounded once on the way in
            assert np.abs(got - want).max() < 1e-6 * max(1.0, np.abs(want).max()), (key, i)
    assert out["res"]["D"][2][0] == 0 and out["res"]["G"][2][0] == 0   # group 0 starts at the optimiser's base

    mu = out["shift"] + out["S1"] / out["n"]
    sigma = (out["S2"] - np.outer(out["S1"], out["S1"]) / out["n"]) / (out["n"] - 1)
    mu_ref, sig_ref = fid_oracle.activation_statistics(out["feats"])
    assert np.abs(mu - mu_ref).max() < 1e-12 and np.abs(sigma - sig_ref).max() < 1e-10
