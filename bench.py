#!/usr/bin/env python
"""Headline benchmark: LittleGAN G+D(+Adjuster) train step, images/sec at 128x128 (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W]          # this repo's CUDA path
    python bench.py --impl reference [...]                        # the CPU restatement, host cores

One "step" = one full `_train_step` (batch_no > 10: input augmentation + generator + discriminator +
adjuster sub-step, use_partition off, sample.config.json hyper-parameters, cond_dim 40) on a synthetic
CelebA-shaped batch.  Per-GPU batch is fixed at 64 (configs[1]; at 8 GPUs this is configs[2]'s
global batch 512) => weak scaling.  images/sec counts `batch_size` images per step (the
reference's progress bar counts 2x that, eager_trainer.py:213).

`value`   : device-resident inputs, K CUDA-graph replays between two CUDA events (max over ranks).
`e2e`     : the public API (`EagerTrainer._train_step(batch_no, iterator)`) with pinned HOST
            batches: every step includes the host->device copies of both real batches and a
            device->host read of the three loss scalars.
`roofline`: the dominant tcgen05 conv kernel timed alone with CUDA events (algorithmic FLOPs /
            launch duration) against the measured bf16 peak in MEASURED_PEAKS.json.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PER_GPU_BATCH = 64
COND_DIM = 40
# conv MACs per image: E = encoder fwd, Dc = decoder + final conv fwd
E_MAC, DC_MAC = 596.4e6, 825.8e6
# The reference step is 7 Dc + 13 E conv passes per image (27.07 GFLOP); this implementation EXECUTES 12 E: the
# encoder pass over `fake` that the reference computes twice (D(fake) and adjuster([real ; fake])) runs once.
STEP_FLOP_PER_IMG = 2 * (7 * DC_MAC + 12 * E_MAC)      # 25.88 GFLOP executed, adjuster on
# DRAM bytes of one captured step at batch 64: ncu dram__bytes_read.sum + dram__bytes_write.sum summed over the 152
# launches of one graph replay (an offline measurement - ncu serialises and cold-starts every launch, so this is an
# upper bound of the replayed step's traffic; scripts/ncu_step_dram.py over profiles/r2_launches_bench_final.csv)
STEP_DRAM_BYTES = {"bytes": 6.348e9, "read": 5.582e9, "write": 0.766e9, "hbm_floor_ms": 6.348e9 / 6484.6e9 * 1e3,
                   "cold_serialised_kernel_ms": 3.947,
                   "source": "profiles/r2_step_dram_final.txt (ncu launch list of `bench.py --steps 1 --warmup 3`, "
                             "final build of the round; not re-measured by this run)"}


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(tf=p["bf16_tflops"], tf_sus=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    hbm=p["hbm_gbs"], src="measured")
    except Exception:
        return dict(tf=1590.0, tf_sus=1400.0, hbm=6650.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); smax = float(r[1])
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def _bench_args(batch):
    from littlegan_b200.config import Arg
    return Arg.from_dict(batch_size=batch, attr=list(range(COND_DIM)), use_partition=False, train_adj=True,
                         dtype="bf16", cuda_graph=True, augment=True)


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle (CPU restatement of the TF-1.15 path) on host cores
# ------------------------------------------------------------------------------------------------
def cpu_oracle_rate(sample_batch, steps, warmup):
    import torch
    from oracle import littlegan_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    oargs = O.make_args(cond_dim=COND_DIM, batch_size=sample_batch, use_partition=False, train_adj=True)
    tr = O.OracleTrainer(oargs, O.init_weights(oargs, 0), dtype=torch.float32)
    data = O.synthetic_batch(oargs, sample_batch, seed=0)
    for i in range(warmup):
        tr.train_step(11 + i, *data)
    t0 = time.perf_counter()
    for i in range(steps):
        tr.train_step(11 + warmup + i, *data)
    dt = time.perf_counter() - t0
    return sample_batch * steps / dt, dt / steps * 1e3, cores


def _config(world):
    """The workload both arms are quoted on (BASELINE.json configs[1]; at 8 GPUs configs[2]'s global batch 512)."""
    return {"workload": "littlegan-128 full train step (G+D+Adjuster), cond 40, batch 64/GPU",
            "per_gpu_batch": PER_GPU_BATCH, "global_batch": PER_GPU_BATCH * world, "parallelism": "dp%d" % world,
            "l2": "per-step working set (GBs of activations) >> 126 MB L2; no flush needed"}


def run_reference(a):
    """The reference's CPU path (its restatement: TF 1.15 is not installable) on rank 0's host cores, on the
    product arm's config.  Every step is one REAL 64-image per-GPU batch - the whole global batch at N = 1, a
    1/N sample of it at N > 1 (one host, one process: the CPU arm has no data parallelism to offer)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(a.gpus)))
    sample = PER_GPU_BATCH
    rate, ms, cores = cpu_oracle_rate(sample, a.steps, a.warmup)
    unit = "images/sec"
    line = {
        "impl": "reference", "metric": "LittleGAN G+D+A train step throughput @128x128", "value": rate,
        "unit": unit, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": _config(world),
        "cpu_baseline": {"value": rate, "unit": unit, "cores": cores, "kind": "port",
                         "sample": "every step = one full train step on a %d-image batch (the per-GPU batch of the "
                                   "config%s) on ONE host process using all %d cores; oracle restatement in "
                                   "PyTorch-CPU fp32 (TF 1.15 not installable)" % (
                                       sample, "" if world == 1 else ": 1/%d of the global batch" % world, cores)},
        "e2e": {"value": rate, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# product arm
# ------------------------------------------------------------------------------------------------
def dominant_kernel_roofline(peaks):
    """Times the largest tcgen05 conv geometry alone (dec2 forward == enc3 dgrad: small 16x16x256 ->
    big 32x32x128, 209.7 MMAC/img) with CUDA events on the launching stream."""
    import torch
    from littlegan_b200 import kernels as K
    N, Hb, Wb, A, B, s = PER_GPU_BATCH * 2, 32, 32, 128, 256, 2
    if not K.tc_supported(K.OP_DGRAD, N, Hb, Wb, A, B, s):
        return None
    x = torch.randn(N, Hb // s, Wb // s, B, device="cuda").to(torch.bfloat16)
    W = (torch.randn(5, 5, A, B, device="cuda") * 0.05)
    bias = torch.zeros(A, device="cuda")
    wp = torch.empty(K.pack_conv_weights_bytes(A, B), dtype=torch.uint8, device="cuda")
    K.pack_conv_weights(W, wp)
    out = torch.empty(N, Hb, Wb, A, device="cuda", dtype=torch.bfloat16)
    stats = torch.zeros(N, 2, dtype=torch.float64, device="cuda")
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    reps, tot = 10, 0.0
    for i in range(3 + reps):
        flush.zero_()                                   # evict L2 between launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        K.conv2d_dgrad(x, W, bias, out, stats, s, K.ACT_NONE, wp, True)
        e1.record()
        e1.synchronize()
        if i >= 3:
            tot += e0.elapsed_time(e1)
    ms = tot / reps
    flops = 2.0 * 25 * A * B * N * (Hb // s) * (Wb // s)
    ach = flops / (ms * 1e-3) / 1e12
    return {"bound": "tensor", "kernel": "tc_conv_kernel<dgrad> dec2 (16x16x256 -> 32x32x128), batch %d" % N,
            "achieved": ach, "peak": peaks["tf"], "unit": "TFLOP/s", "frac": ach / peaks["tf"],
            "peak_source": peaks["src"] + " bf16 burst", "ms_per_launch": ms,
            # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed ncu --set full capture of
            # the round's final build (profiles/r2_ncu_tc_dgrad_dec2.txt): 18.48 MB + 0.05 MB (tensor pipe 57.6% active);
            # algorithmic input bytes 18.4 MB, the 33.6 MB output stays in the 126 MB L2 for the kernel's lifetime
            "traffic": 18.53e6, "traffic_unit": "bytes/launch",
            "traffic_source": "profiles/r2_ncu_tc_dgrad_dec2.txt (ncu --set full of this launch, final build of the "
                              "round; not re-measured by this run)"}


def _time_launch(fn, flush, reps=10, warm=3):
    """Average CUDA-event time (ms) of fn() on the current stream, L2 evicted before every launch."""
    import torch
    tot = 0.0
    for i in range(warm + reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        if i >= warm:
            tot += e0.elapsed_time(e1)
    return tot / reps


def hbm_rooflines(peaks, trainer):
    """The bandwidth-bound kernels of the step, each timed ALONE at its largest step geometry (decoder layer 4 of
    the adjuster's 2B batch: [128,128,128,32] bf16 = 134 MB per map; the generator's Adam range), algorithmic
    bytes (SURVEY 8 d) / launch time against the measured HBM copy bandwidth."""
    import torch
    from littlegan_b200 import kernels as K
    N, H, C = 2 * PER_GPU_BATCH, 128, 32
    M = H * H * C
    bf = torch.bfloat16
    z = torch.randn(N, H, H, C, device="cuda").to(bf)
    g = torch.randn(N, H, H, C, device="cuda").to(bf)
    out = torch.empty_like(z)
    gamma, beta = torch.ones(1, device="cuda"), torch.zeros(1, device="cuda")
    dgamma, dbeta = torch.zeros(1, device="cuda"), torch.zeros(1, device="cuda")
    dbias = torch.zeros(C, device="cuda")
    zf = z.float().reshape(N, -1)
    stats = torch.stack([zf.sum(1), (zf * zf).sum(1)], 1).double().contiguous()
    del zf
    red = torch.zeros(N, 2, dtype=torch.float64, device="cuda")
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    map_bytes = N * M * 2
    res = []

    def add(name, ms, nbytes, what):
        gbs = nbytes / (ms * 1e-3) / 1e9
        res.append({"kernel": name, "bound": "hbm", "achieved": gbs, "peak": peaks["hbm"], "unit": "GB/s",
                    "frac": gbs / peaks["hbm"], "ms_per_launch": ms, "algorithmic_bytes": nbytes, "what": what})

    ms = _time_launch(lambda: K.instnorm_act_fwd(z, stats, gamma, beta, None, out, 1e-3, 1.0, 0.3), flush)
    add("instnorm_fwd (IN apply + LeakyReLU)", ms, 2 * map_bytes, "1 read + 1 write of a [128,128,128,32] bf16 map")
    fuse = K.bias_grad_fusable(z)

    def bwd_two_pass():
        red.zero_()
        K.instnorm_act_bwd(g, z, stats, gamma, beta, red, out, dgamma, dbeta, 1e-3, 1.0, 0.3, dy_ready=False,
                           dbias=dbias if fuse else None)
    ms = _time_launch(bwd_two_pass, flush)
    add("instnorm_bwd_reduce + instnorm_bwd_apply", ms, 5 * map_bytes,
        "reduce: read g, z; apply: read g, z, write dz (+ fused d bias / d gamma / d beta)")
    red.normal_()
    ms = _time_launch(lambda: K.instnorm_act_bwd(g, z, stats, gamma, beta, red, out, dgamma, dbeta, 1e-3, 1.0, 0.3,
                                                 dy_ready=True, dbias=dbias if fuse else None), flush)
    add("instnorm_bwd_apply<dy_ready> (pass 1 fused into the producing conv)", ms, 3 * map_bytes,
        "read dy, z, write dz (+ fused d bias / d gamma / d beta)")
    lo, hi = trainer._range("Generator", 11)
    n = hi - lo
    P, G, Mm, V = (torch.randn(n, device="cuda") for _ in range(4))
    V.abs_()
    st = torch.zeros(4, dtype=torch.float64, device="cuda")
    K.adam_advance(st, 5e-5, 0.5, 0.9)
    ms = _time_launch(lambda: K.adam_apply(P, G, Mm, V, st, 0.5, 0.9, 1e-8, 0.0), flush)
    add("adam_apply (generator range, %d params)" % n, ms, 28 * n, "read g, p, m, v; write p, m, v (fp32)")
    return res


def sustained_run(graph, n_steps, B, world, local, rank):
    """Replays the captured step back to back `n_steps` times (>= 5 s of work; the count is fixed up front and identical
    on every rank - the graph contains the NCCL all-reduces, so all ranks must replay it equally often) with the clock
    sampler on: the throughput a long training run sees (power / clock steady state), next to the short timed region
    of `value`."""
    import torch
    import torch.distributed as dist
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
        sampler.rows.clear()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n_steps):
        graph.replay()
        if i % 500 == 499:
            torch.cuda.current_stream().synchronize()       # bound the launch queue
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n_steps
    clocks = sampler.stop() if rank == 0 else None
    power = None
    if rank == 0:
        pw = []
        for r in sampler.rows:
            try:
                pw.append(float(r[2]))
            except Exception:
                pass
        pw.sort()
        power = pw[len(pw) // 2] if pw else None
    return {"seconds": ms * n_steps / 1e3, "steps": n_steps, "ms_per_step": ms,
            "images_per_sec_per_gpu": B / (ms * 1e-3), "clocks": clocks, "power_w_median": power}


def predict_bench(trainer, world, reps=5):
    """configs[3]: EagerTrainer.predict (eager_trainer.py:265-298), batch 256 per GPU, HOST numpy inputs, wall
    clock per call (every call ends in a device->host read of the MSE scalars and the x100 integer lists)."""
    import numpy as np
    import torch
    B = 256
    a = trainer.args
    rng = np.random.default_rng(0)
    noise = rng.standard_normal((B, a.noise_dim)).astype(np.float32)
    cond = (0.96 * rng.choice([-1.0, 1.0], size=(B, a.cond_dim)) + 0.02).astype(np.float32)
    image = rng.uniform(-1, 1, size=(B, 128, 128, 3)).astype(np.float32)
    for _ in range(2):
        trainer.predict(noise, cond, image)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        trainer.predict(noise, cond, image)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / reps
    return ms, B


def fid50k_bench(world, rank):
    """configs[4]: FID statistics of 50 000 synthetic 128x128 images sharded over the ranks: host image bytes ->
    Inception pool_3 (random weights: the 2015 model file is an external download) -> streaming (n, S1, S2) -> one
    SUM all-reduce -> mu / sigma -> Frechet distance against a second statistics pair on rank 0.  Each rank cycles a
    pool of <= 5000 host images (bounded host memory and generation time); every image crosses PCIe every time."""
    import numpy as np
    import torch
    from littlegan_b200 import fid
    from littlegan_b200.inception import InceptionPool3
    total = 50000
    per = total // world
    pool_n = min(per, 5000)
    pool = np.random.default_rng(100 + rank).integers(0, 256, (pool_n, 128, 128, 3), dtype=np.uint8)
    net = InceptionPool3(seed=0, dtype="bf16")
    acc = fid.FeatureStatistics(2048)
    for feats in fid.get_activations(pool[:300], net, 100):          # warm-up (kernels, pinned staging, NCCL)
        acc.update(feats)
    acc.finalize()
    torch.cuda.synchronize()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    t0 = time.perf_counter()
    acc = fid.FeatureStatistics(2048)
    done = 0
    while done < per:
        n = min(pool_n, per - done)
        for feats in fid.get_activations(pool[:n], net, 100):
            acc.update(feats)
        done += n
    mu, sigma = acc.finalize()
    torch.cuda.synchronize()
    t_stats = time.perf_counter() - t0
    d = None
    t_dist = 0.0
    if rank == 0:
        mu2 = mu * 1.01 + 0.01
        sigma2 = sigma * 1.1 + torch.eye(2048, dtype=torch.float64, device=sigma.device) * 1e-3
        fid.calculate_frechet_distance(mu, sigma, mu2, sigma2)       # warm-up
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        d = fid.calculate_frechet_distance(mu, sigma, mu2, sigma2)
        t_dist = time.perf_counter() - t1
    return {"images": per * world, "per_rank": per, "features_statistics_s": t_stats, "frechet_distance_s": t_dist,
            "images_per_sec": per * world / t_stats, "distance": d,
            "note": "wall clock incl. host->device image bytes; Inception on random weights"}


def run_product(a):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from littlegan_b200 import kernels as K
    from littlegan_b200 import model as M
    from littlegan_b200.dataset import DevicePrefetcher, SyntheticCelebA
    from littlegan_b200.eager_trainer import EagerTrainer

    args = _bench_args(PER_GPU_BATCH)
    M.set_init_seed(0)
    dec, enc = M.Decoder(args), M.Encoder(args)
    gen, disc = M.Generator(args, dec), M.Discriminator(args, enc)
    adj = M.Adjuster(args, disc, gen)
    data = SyntheticCelebA(args, batches=10 ** 9, seed=1 + rank, pool=8)
    trainer = EagerTrainer(args, gen, disc, adj, data)
    it = DevicePrefetcher(data.get_new_iterator(), depth=4)
    B = PER_GPU_BATCH

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                 # nvidia-smi needs ~0.2 s to produce its first sample

    # ---- warm-up through the public API (first step eager, second captures the graph)
    b = 11
    for _ in range(max(a.warmup, 3)):
        b += 1
        trainer._train_step(b, it)
    torch.cuda.synchronize()
    graph = trainer._graphs[(True, None, True, True)][0]
    launches = K.launch_count_of_last_capture()

    # ---- device-resident timing: K replays of the captured step
    if rank == 0:
        sampler.rows.clear()            # keep only samples taken during the timed regions
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        graph.replay()
    e1.record()
    barrier()
    ms_dev = e0.elapsed_time(e1) / a.steps

    # ---- end to end through the public API, host batches, loss read-back every step
    # (the loss scalars of step i are read on the host right after step i+1 has been submitted: every step's
    # result is read inside the timed region, without draining the GPU between steps)
    barrier()
    t0 = time.perf_counter()
    prev = None
    for _ in range(a.steps):
        b += 1
        res = trainer._train_step(b, it)
        if prev is not None:
            losses = (float(prev[3]), float(prev[4]), float(prev[5]))
        prev = res
    losses = (float(prev[3]), float(prev[4]), float(prev[5]))
    barrier()
    ms_e2e = (time.perf_counter() - t0) * 1e3 / a.steps
    clocks = sampler.stop() if rank == 0 else None
    assert all(x == x for x in losses), "non-finite loss"

    t = torch.tensor([ms_dev, ms_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e = float(t[0]), float(t[1])

    # ---- the contract line, complete before any extra runs
    line = None
    if rank == 0:
        peaks = _peaks()
        gb = B * world
        H = args.init_dim * 16
        h2d = 2 * (B * H * H * args.image_channel * 4 + B * args.cond_dim * 4)
        value = gb / (ms_dev * 1e-3)
        line = {
            "metric": "LittleGAN G+D+A train step throughput @128x128", "value": value, "unit": "images/sec",
            "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms_dev,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": _config(world),
            "clocks": clocks,
            "e2e": {"value": gb / (ms_e2e * 1e-3), "unit": "images/sec", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 12},
            "gpu_launches": launches * a.steps,
            "step_tensor_frac": STEP_FLOP_PER_IMG * B / (ms_dev * 1e-3) / 1e12 / peaks["tf_sus"],
            "last_losses": {"gen": losses[0], "disc": losses[1], "adj": losses[2]},
        }
        if world == 1:
            line["roofline"] = dominant_kernel_roofline(peaks)
            if not a.no_cpu_baseline:
                rate, ms, cores = cpu_oracle_rate(16, 4, 1)
                line["cpu_baseline"] = {
                    "value": rate, "unit": "images/sec", "cores": cores, "kind": "port",
                    "sample": "4 full train steps on a 16-image batch (configs[0]), oracle restatement in "
                              "PyTorch-CPU fp32 (TF 1.15 not installable)"}

    # ---- extras (none of them inside the timed regions above): sustained replay, configs[3] predict, configs[4] FID.
    # Every count below is derived from values that are identical on all ranks (the all-reduced ms_dev): the captured
    # step and the FID statistics contain collectives.  A watchdog guards them: should an extra stall (a lost rank,
    # a wedged collective), every rank leaves after `--extras-timeout` seconds and rank 0 still prints the contract
    # line, with the reason under extra_keys.
    extra = {}
    done = threading.Event()

    def _watchdog():
        if done.wait(a.extras_timeout):
            return
        if rank == 0 and line is not None:
            line["extra_keys"] = {"aborted": "extras exceeded %.0f s and were abandoned" % a.extras_timeout}
            print(json.dumps(line), flush=True)
        sys.stdout.flush()
        os._exit(0)

    if not a.no_extras:
        threading.Thread(target=_watchdog, daemon=True).start()
        n_sus = max(100, int(a.sustain_seconds * 1e3 / ms_dev / 100.0 + 0.5) * 100)
        sus = sustained_run(graph, n_sus, B, world, local, rank)
        ms_pred, b_pred = predict_bench(trainer, world)
        fidr = fid50k_bench(world, rank)
        tt = torch.tensor([sus["ms_per_step"], ms_pred, fidr["features_statistics_s"]], dtype=torch.float64,
                          device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        sus["ms_per_step"] = float(tt[0])
        sus["images_per_sec"] = B * world / (float(tt[0]) * 1e-3)
        fidr["features_statistics_s"] = float(tt[2])
        fidr["images_per_sec"] = fidr["images"] / float(tt[2])
        extra = {"sustained": sus,
                 "predict": {"workload": "configs[3]: predict, batch 256/GPU, host numpy inputs", "ms_per_call": float(tt[1]),
                             "images_per_sec": b_pred * world / (float(tt[1]) * 1e-3), "per_gpu_batch": b_pred},
                 "fid50k": fidr}
        if world == 1 and rank == 0:
            extra["roofline_hbm"] = hbm_rooflines(_peaks(), trainer)
            # sum of dram__bytes_read + dram__bytes_write over one replay of the captured step (ncu, batch 64);
            # an offline number: profiles/ holds the launch list it was summed from
            extra["step_dram_bytes"] = STEP_DRAM_BYTES
    # data-parallel invariant: the replicas must hold bit-identical parameters after any number of steps
    if world > 1:
        ref = trainer.P.clone()
        dist.broadcast(ref, src=0)
        same = torch.tensor([1.0 if torch.equal(ref, trainer.P) else 0.0], device="cuda")
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        assert float(same) == 1.0, "parameter replicas diverged across ranks"
        extra["replicas_bit_identical"] = True
    done.set()
    if rank == 0:
        line["extra_keys"] = extra
        print(json.dumps(line), flush=True)
    if world > 1:
        # Captured graphs hold NCCL work; tearing the communicator down underneath them can block.
        # Everything is measured and printed: synchronise, then leave without the destructor chain.
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="graft", choices=["graft", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the sustained / predict / fid50k / HBM-roofline extras")
    ap.add_argument("--sustain-seconds", type=float, default=5.0)
    ap.add_argument("--extras-timeout", type=float, default=150.0)
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_product(a)


if __name__ == "__main__":
    main()
