"""End-to-end parity of the model forwards and the full train step against the CPU oracle with
identical weights and inputs.  -m gpu."""
import json
import os

import pytest
import torch

from oracle import littlegan_oracle as O
from tests.util import build_product, product_args, product_weights_to_oracle, rel_err, small_args, tol

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


class _ListIterator:
    def __init__(self, items):
        self.items, self.i = items, 0

    def get_next(self):
        from littlegan_b200.eager_trainer import OutOfRangeError
        if self.i >= len(self.items):
            raise OutOfRangeError()
        self.i += 1
        return self.items[self.i - 1]


def _setup(oargs, dtype, cuda_graph=False, seed=0):
    from littlegan_b200.eager_trainer import EagerTrainer
    pargs = product_args(oargs, dtype=dtype, cuda_graph=cuda_graph)
    gen, disc, adj = build_product(pargs, seed)
    W = product_weights_to_oracle(gen, disc, adj)
    trainer = EagerTrainer(pargs, gen, disc, adj, None)
    return pargs, gen, disc, adj, trainer, W


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_model_forwards(dtype):
    oargs = small_args()
    pargs, gen, disc, adj, trainer, W = _setup(oargs, dtype)
    i1, c1, i2, c2, noise = O.synthetic_batch(oargs, 4, seed=3)
    t = tol(torch.float32 if dtype == "fp32" else torch.bfloat16)
    with torch.no_grad():
        Wd = {k: [w.double() for w in v] for k, v in W.items()}
        ref_img = O.generator(oargs, noise.double(), c2.double(), Wd["G"])
        ref_pr, ref_c = O.discriminator(oargs, i1.double(), Wd["D"])
        ref_adj = O.adjuster(oargs, i1.double(), ((c2 + 1) * 0.5).double(), Wd["D"], Wd["G"], Wd["A"])
        ref_enc = O.encoder(oargs, i1.double(), Wd["D"][:16])
    img = gen([noise, c2])
    pr, c = disc(i1)
    adj_img = adj([i1, (c2 + 1) * 0.5])
    enc = disc.encoder(i1)
    assert rel_err(img, ref_img) < t * 3
    assert rel_err(pr, ref_pr) < t * 3 and rel_err(c, ref_c) < t * 3
    assert rel_err(adj_img, ref_adj) < t * 3
    for a, b in zip(enc, ref_enc):
        assert rel_err(a, b) < t * 3


def _one_step_parity(oargs, batch_no, B, t_loss, t_grad, t_w, fp32_floor=False):
    """t_grad: bound every gradient tensor must meet.  fp32_floor (full-size maps): LeakyReLU's derivative is
    discontinuous at 0 - in a 10^6-element map a few pre-activations lie within fp32 rounding of 0, and each such
    sign flip moves the affected gradient entries by one of their ~10^3 summands - so no two fp32 implementations
    agree to 1e-4 on every tensor; the bound of a tensor is then 3x the CPU fp32 oracle's OWN distance from the
    fp64 oracle on that tensor (never below t_grad)."""
    pargs, gen, disc, adj, trainer, W = _setup(oargs, "fp32")
    ot = O.OracleTrainer(oargs, W, dtype=torch.float64)
    i1, c1, i2, c2, noise = O.synthetic_batch(oargs, B, seed=5)
    ref = ot.train_step(batch_no, i1, c1, i2, c2, noise, return_grads=True)
    ref32 = None
    if fp32_floor:
        ref32 = O.OracleTrainer(oargs, W, dtype=torch.float32).train_step(batch_no, i1, c1, i2, c2, noise,
                                                                          return_grads=True)
    res = trainer._train_step(batch_no, _ListIterator([(i1, c1), (i2, c2)]), noise=noise)
    assert res[0] is True
    assert rel_err(res[1], ref["fake_image"]) < 1e-4
    assert abs(float(res[3]) - float(ref["gen_loss"])) < t_loss * abs(float(ref["gen_loss"]))
    assert abs(float(res[4]) - float(ref["disc_loss"])) < t_loss * abs(float(ref["disc_loss"]))
    if ref["adj_loss"] is not None:
        assert rel_err(res[2], ref["adj_image"]) < 1e-4
        assert abs(float(res[5]) - float(ref["adj_loss"])) < t_loss * abs(float(ref["adj_loss"]))
    else:
        assert res[2] is None and res[5] is None
    # gradients (the arenas keep the unclipped gradient; the oracle's D grads are clipped)
    names = {"D": disc.weights, "G": gen.weights, "A": adj.weights[16:20]}
    worst, errs = 0.0, []
    for key, ws in names.items():
        if ref["grads"][key] is None:
            continue
        for idx, gref in ref["grads"][key].items():
            got = ws[idx].lg_grad
            if key == "D" and oargs.use_clip:
                got = got.clamp(-oargs.clip_range, oargs.clip_range)
            if gref.numel() == 1:
                # scalar gamma/beta gradients: d(gamma) is analytically ~0 (the next layer's norm
                # removes the scale), i.e. a sum of O(1) terms that cancels - bound it absolutely
                e = abs(float(got) - float(gref)) / max(abs(float(gref)), 1e-1)
            else:
                e = rel_err(got, gref)
            worst = max(worst, e)
            errs.append(e)
            bound = t_grad
            if ref32 is not None:
                g32 = ref32["grads"][key][idx]
                own = (abs(float(g32) - float(gref)) / max(abs(float(gref)), 1e-1)) if gref.numel() == 1 \
                    else rel_err(g32, gref)
                # scalar gamma / beta gradients are cancelling sums over 1e5-1e6 fp32 terms: 1e-3 of their scale
                bound = max(t_grad if gref.numel() > 1 else 1e-3, 3.0 * own)
            assert e < bound, (key, idx, e, bound)
    # updated weights.  One TF-Adam step moves every weight by ~1.58*lr regardless of |g|, so a
    # gradient that is pure rounding noise (the ~0 d(gamma) above) can legitimately flip the step:
    # the bound is a few lr relative to max|w|, not fp32 epsilon.
    step = 1.6 * oargs.lr * (1 - 0.9) ** 0.5 / (1 - 0.5)        # size of the first Adam step
    for key, ws in names.items():
        for idx, w in enumerate(ws):
            diff = (w.detach().double().cpu() - ot.W[key][idx].detach()).abs()
            assert float(diff.max()) < 2.5 * step, (key, idx, float(diff.max()))   # at most a sign flip
            if w.numel() > 1:
                frac = float((diff > 0.05 * step).double().mean())
                assert frac < max(50 * t_w, 2.0 / w.numel()), (key, idx, frac)   # and only for noise-level grads
    return worst


@pytest.mark.parametrize("batch_no", [3, 11, 15, 20])
def test_train_step_small_fp32(batch_no):
    """batch 3: G+D only; 11: with adjuster; 15/20: partition groups 0 and 1 (b % 5 == 0)."""
    oargs = small_args(use_partition=True)
    _one_step_parity(oargs, batch_no, 4, 1e-4, 1e-4, 1e-4)


def test_train_step_full_size_fp32():
    """The real 128x128 architecture (cond 40), batch 2, full step with adjuster."""
    oargs = O.make_args(cond_dim=40, batch_size=2, use_partition=False)
    _one_step_parity(oargs, 11, 2, 1e-4, 1e-4, 1e-4, fp32_floor=True)


def test_use_gp_raises():
    oargs = small_args(use_gp=True)
    pargs, gen, disc, adj, trainer, W = _setup(oargs, "fp32")
    i1, c1, i2, c2, noise = O.synthetic_batch(oargs, 4, seed=5)
    with pytest.raises(NotImplementedError):
        trainer._train_step(1, _ListIterator([(i1, c1), (i2, c2)]))


def test_iterator_protocol():
    oargs = small_args()
    pargs, gen, disc, adj, trainer, W = _setup(oargs, "fp32")
    i1, c1, i2, c2, noise = O.synthetic_batch(oargs, 4, seed=5)
    assert trainer._train_step(1, _ListIterator([(i1, c1)])) == (None,)          # OutOfRange
    assert trainer._train_step(1, _ListIterator([(i1, c1), (i2[:3], c2[:3])])) == (False,)  # short batch


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_trajectory_graph_vs_oracle(dtype):
    """Loss trajectory over 30 steps on the reduced architecture: CUDA-graph replay path against
    the oracle (fp32), within 1% (north_star) - crosses b>10 (adjuster on) and partition steps."""
    oargs = small_args(use_partition=True)
    pargs, gen, disc, adj, trainer, W = _setup(oargs, dtype, cuda_graph=True)
    ot = O.OracleTrainer(oargs, W, dtype=torch.float64)
    steps = 30
    for b in range(1, steps + 1):
        i1, c1, i2, c2, noise = O.synthetic_batch(oargs, 4, seed=100 + b)
        ref = ot.train_step(b, i1, c1, i2, c2, noise)
        res = trainer._train_step(b, _ListIterator([(i1, c1), (i2, c2)]), noise=noise)
        for got, want in ((res[3], ref["gen_loss"]), (res[4], ref["disc_loss"]), (res[5], ref["adj_loss"])):
            if want is None:
                assert got is None
                continue
            assert abs(float(got) - float(want)) < 0.01 * abs(float(want)), (b, float(got), float(want))
    assert len(trainer._graphs) >= 2     # the replay path was actually exercised


def test_trajectory_full_size_golden():
    """100-step G/D/A loss trajectory of the real architecture (cond 40, batch 4) against the
    oracle's committed fp32 trajectory (tests/golden/trajectory_full.json, made by
    tests/golden/make_golden.py).

    The system is chaotic under TF-Adam: the first Adam step is +-1.58*lr for EVERY weight, so
    weights whose gradient is rounding noise (the analytically-zero d(gamma)s) random-walk
    differently in any two fp32 implementations, and the loss gap grows from 1e-7 (step 1) to
    ~1e-4 (step 2) to ~1% (step ~10).  The oracle's own fp32-vs-fp64 gap
    (trajectory_full_fp64.json) is the noise floor (1% at step 11, 1.9% at step 13).  Asserted: within 1%
    (north_star) until the divergence sets in (first 10 steps, fp32), within the bf16 tolerance of 2% over
    the first 8 steps in bf16 mode (measured <= 0.7%; step 10 is a sensitive batch where bf16 rounding
    already moves the loss by 2-4%), and a median deviation over all 100 steps that stays within 2x (fp32) /
    5x (bf16) of that noise floor."""
    import numpy as np
    path = os.path.join(GOLD, "trajectory_full.json")
    gold = json.load(open(path))
    g64 = json.load(open(os.path.join(GOLD, "trajectory_full_fp64.json")))
    floor = np.median([abs(a - b) / abs(b) for a, b in zip(gold["gen"] + gold["disc"], g64["gen"] + g64["disc"])])
    oargs = O.make_args(**gold["args"])
    B = gold["args"]["batch_size"]
    for dtype, first10, nfirst, factor in (("fp32", 0.01, 10, 2.0), ("bf16", 0.02, 8, 5.0)):
        pargs, gen, disc, adj, trainer, W = _setup(oargs, dtype, cuda_graph=True, seed=gold["seed"])
        devs = []
        for b in range(1, len(gold["gen"]) + 1):
            i1, c1, i2, c2, noise = O.synthetic_batch(oargs, B, seed=gold["data_seed"] + b)
            res = trainer._train_step(b, _ListIterator([(i1, c1), (i2, c2)]), noise=noise)
            for got, w in zip(res[3:5], (gold["gen"][b - 1], gold["disc"][b - 1])):
                d = abs(float(got) - w) / abs(w)
                assert np.isfinite(d)
                devs.append(d)
                if b <= nfirst:
                    assert d < first10, (dtype, b, float(got), w)
        med = float(np.median(devs))
        print("trajectory %s: median deviation %.4f (oracle fp32-vs-fp64 floor %.4f)" % (dtype, med, floor))
        assert med < factor * max(floor, 0.005), (dtype, med, floor)


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_predict_matches_oracle(dtype):
    """predict (eager_trainer.py:265-298): 1 G + 2 D + 2 A forwards and the four MSE scalars."""
    oargs = small_args()
    pargs, gen, disc, adj, trainer, W = _setup(oargs, dtype)
    i1, c1, i2, c2, noise = O.synthetic_batch(oargs, 4, seed=7)
    ot = O.OracleTrainer(oargs, W, dtype=torch.float64)
    ref_img, ref_save, ref_ar, ref_af = ot.predict(noise, c1, i1)
    img, save, ar, af = trainer.predict(noise, c1, i1)
    t = tol(torch.float32 if dtype == "fp32" else torch.bfloat16) * 3
    assert rel_err(img, ref_img) < t and rel_err(ar, ref_ar) < t and rel_err(af, ref_af) < t
    for key in ("real_pr_mse", "real_c_mse", "fake_pr_mse", "fake_c_mse"):
        assert abs(save[key] - ref_save[key]) < max(t * abs(ref_save[key]), 1e-6), key
    # the x100-rounded integer lists of the reference's JSON dump
    want = torch.round(ref_save["real_c"] * 100).to(torch.int64)
    got = torch.tensor(save["real_c"])
    assert int((got - want).abs().max()) <= (0 if dtype == "fp32" else 2)
    assert save["real_cond"] == torch.round(c1 * 100).to(torch.int64).tolist()


def test_public_loss_functions():
    """discriminator_loss / generator_loss / adjuster_loss keep the reference's signatures."""
    from littlegan_b200.eager_trainer import EagerTrainer
    oargs = small_args()
    pargs, gen, disc, adj, trainer, W = _setup(oargs, "fp32")
    g = torch.Generator().manual_seed(0)
    c_true = O.soft((torch.rand(4, oargs.cond_dim, generator=g) < 0.5).float() * 2 - 1)
    c_pred, pr_r, pr_f = (torch.rand(4, oargs.cond_dim, generator=g), torch.rand(4, 1, generator=g),
                          torch.rand(4, 1, generator=g))
    img_a, img_b = torch.rand(4, 32, 32, 3, generator=g) * 2 - 1, torch.rand(4, 32, 32, 3, generator=g) * 2 - 1
    d = EagerTrainer.discriminator_loss(c_true.cuda(), c_pred.cuda(), pr_r.cuda(), pr_f.cuda())
    assert abs(float(d) - float(O.discriminator_loss(c_true, c_pred, pr_r, pr_f))) < 1e-5
    gl = trainer.generator_loss(c_true.cuda(), c_pred.cuda(), pr_f.cuda(), img_a.cuda(), img_b.cuda())
    assert abs(float(gl) - float(O.generator_loss(oargs, c_true, c_pred, pr_f, img_a, img_b))) < 1e-5
    al = trainer.adjuster_loss(c_true.cuda(), c_pred.cuda(), pr_f.cuda(), img_a.cuda(), img_b.cuda())
    assert abs(float(al) - float(gl)) < 1e-7


def _run_steps(oargs, n, **extra):
    """Final parameter arena and per-step losses of n train steps (batch_no 11.., adjuster on from 11) from the
    same initial weights and batches."""
    from littlegan_b200.eager_trainer import EagerTrainer
    pargs = product_args(oargs, dtype="fp32", cuda_graph=extra.pop("cuda_graph", False))
    for k, v in extra.items():
        setattr(pargs, k, v)
    gen, disc, adj = build_product(pargs, 0)
    trainer = EagerTrainer(pargs, gen, disc, adj, None)
    i1, c1, i2, c2, noise = O.synthetic_batch(oargs, 4, seed=5)
    losses = []
    for b in range(11, 11 + n):
        res = trainer._train_step(b, _ListIterator([(i1, c1), (i2, c2)]), noise=noise, new_image=i1)
        losses.append(res)
    vals = [[float(r[3]), float(r[4]), float(r[5])] for r in losses]
    return trainer.P.clone(), vals, losses


def test_parallel_chains_match_the_serial_schedule():
    """The three chains of the step on three streams (and as branches of a captured graph) compute what the
    single-stream schedule computes; fp32 atomics make the sums order dependent, hence a tolerance."""
    oargs = small_args()
    P0, l0, _ = _run_steps(oargs, 4, overlap_chains=False, overlap_wgrad=False)
    P1, l1, _ = _run_steps(oargs, 4)
    P2, l2, _ = _run_steps(oargs, 4, cuda_graph=True)
    for P, l in ((P1, l1), (P2, l2)):
        assert rel_err(P, P0) < 1e-5
        for a, b in zip(l, l0):
            for x, y in zip(a, b):
                assert abs(x - y) <= 1e-4 * abs(y) + 1e-6


def test_loss_values_survive_slot_recycling():
    """LossValue reads the pinned copy made behind its step; once the ring slot has been reused by a later
    step it falls back to the device scalar - either way float() is that step's loss."""
    from littlegan_b200.eager_trainer import EagerTrainer
    oargs = small_args()
    n = EagerTrainer._RB_SLOTS + 3
    _, vals, res = _run_steps(oargs, n)
    for r, v in zip(res, vals):            # vals were read after ALL steps: the first ones through the fallback
        assert [float(r[3]), float(r[4]), float(r[5])] == v
        assert abs(float(r[3].tensor) - v[0]) == 0.0 and format(r[4], ".3f") == "%.3f" % v[1]


def test_upload_pinned_chunks_roundtrip():
    import numpy as np
    from littlegan_b200.utils import upload
    rng = np.random.default_rng(0)
    x = rng.standard_normal((37, 128, 128, 3)).astype(np.float32)       # 7 MB: 2 chunks of 3 MB + a tail
    y = upload(x, torch.device("cuda"), chunk_bytes=3 << 20)
    assert y.shape == x.shape and y.dtype == torch.float32
    assert torch.equal(y.cpu(), torch.from_numpy(x))
    small = upload(x[:1], torch.device("cuda"))
    assert torch.equal(small.cpu(), torch.from_numpy(x[:1]))


def test_train_step_with_device_augmentation():
    """`augment: true` (the reference's behaviour): new_image is drawn on the device inside the captured step -
    a different image every replay, the same sequence for the same seed."""
    from littlegan_b200.eager_trainer import EagerTrainer
    oargs = small_args()
    runs = []
    for rep in range(2):
        pargs = product_args(oargs, dtype="fp32", cuda_graph=True, augment=True, seed=7)
        gen, disc, adj = build_product(pargs, 0)
        trainer = EagerTrainer(pargs, gen, disc, adj, None)
        i1, c1, i2, c2, noise = O.synthetic_batch(oargs, 4, seed=5)
        imgs, losses = [], []
        for b in range(11, 16):
            res = trainer._train_step(b, _ListIterator([(i1, c1), (i2, c2)]), noise=noise)
            B = pargs.batch_size
            imgs.append(trainer._static["img3"][B:2 * B].float().cpu().clone())
            losses.append([float(res[3]), float(res[4]), float(res[5])])
        assert all(torch.isfinite(torch.tensor(l)).all() for l in losses)
        for a, b in zip(imgs, imgs[1:]):
            assert not torch.equal(a, b)
        d = (imgs[0] - i1).abs()
        flipped = (imgs[0] - i1.flip(2)).abs()
        per = torch.minimum(d.reshape(4, -1).mean(1), flipped.reshape(4, -1).mean(1))
        assert float(per.max()) < 0.3 and float(per.min()) > 1e-3     # an augmented copy, not the image itself
        runs.append((imgs, losses))
    for a, b in zip(runs[0][0], runs[1][0]):
        assert torch.equal(a, b)


def test_train_loop_checkpoints_and_resumes(tmp_path):
    """train(): epoch loop, periodic image / predict dumps, one checkpoint + status.json per epoch; a new trainer
    with `restore` continues from it (eager_trainer.py:37-43, 180-229)."""
    from littlegan_b200.dataset import SyntheticCelebA
    from littlegan_b200.eager_trainer import EagerTrainer
    oargs = small_args()

    def make(epoch):
        pargs = product_args(oargs, dtype="fp32", cuda_graph=True, epoch=epoch, freq_gen=2, freq_test=3, restore=True,
                             reuse=False, init_dirs=True, all_result_dir=str(tmp_path), exp_name="exp",
                             test_data_dir=str(tmp_path), log_every=2)
        pargs.result_dir = os.path.join(str(tmp_path), "exp")
        gen, disc, adj = build_product(pargs, 0)
        data = SyntheticCelebA(pargs, batches=8, seed=3, pool=2)      # a step draws two batches
        return EagerTrainer(pargs, gen, disc, adj, data)

    t1 = make(2)
    assert t1.global_epoch == 1
    t1.train()
    r = t1.args.result_dir
    assert os.path.isfile(os.path.join(r, "checkpoint", "2.pt"))
    st = json.load(open(os.path.join(r, "checkpoint", "status.json")))
    assert st == {"epoch": 3, "latest": "2.pt"}
    assert os.path.isfile(os.path.join(r, "train", "gen", "1-2.jpg")) and os.path.isfile(os.path.join(r, "train", "gen", "2-4.jpg"))
    assert os.path.isfile(os.path.join(r, "test", "disc", "1-3.json")) and os.path.isfile(os.path.join(r, "test", "adj", "2-3.jpg"))
    t2 = make(3)
    assert t2.global_epoch == 3
    assert torch.equal(t2.P, t1.P) and torch.equal(t2.M, t1.M) and torch.equal(t2.V, t1.V)
    for k in t1.adam_state:
        assert torch.equal(t2.adam_state[k], t1.adam_state[k])
    t2.train()
    assert os.path.isfile(os.path.join(r, "checkpoint", "3.pt")) and not torch.equal(t2.P, t1.P)


def test_train_step_from_decoded_bytes(tmp_path):
    """dataset.CelebA -> DevicePrefetcher -> _train_step: uint8 batches are rescaled on the device and give the
    same step as the reference's host-side data_rescale of the same files."""
    from littlegan_b200.dataset import CelebA, DevicePrefetcher
    from littlegan_b200.utils import data_rescale
    from tests.test_host_logic import _write_celeba
    oargs = small_args()
    _write_celeba(tmp_path, 8, 32, attrs=5)
    extra = dict(image_path=str(tmp_path / "img"), attr_path=str(tmp_path / "list.txt"), image_ext="png",
                 threads=2, prefetch_batch=1)
    runs = []
    for mode in ("bytes", "float"):
        pargs, gen, disc, adj, trainer, W = _setup(oargs, "fp32")
        for k, v in extra.items():
            setattr(pargs, k, v)
        ds = CelebA(pargs, seed=1, as_float=(mode == "float"))
        it = ds.get_new_iterator()
        if mode == "bytes":
            it = DevicePrefetcher(it, depth=2)
        noise = torch.randn(4, oargs.noise_dim, generator=torch.Generator().manual_seed(2))
        res = trainer._train_step(11, it, noise=noise)
        assert res[0] is True
        runs.append([float(res[3]), float(res[4]), float(res[5]), res[1].float().cpu(),
                     trainer._static["in_img1"].float().cpu(), trainer._static["in_img2"].float().cpu()])
        assert trainer._train_step(12, it, noise=noise) == (None,)       # 8 files = 2 batches = one step
    # the staged inputs are bit-identical; the step itself accumulates with atomics (last-bit run-to-run noise)
    assert torch.equal(runs[0][4], runs[1][4]) and torch.equal(runs[0][5], runs[1][5])
    for a, b in zip(runs[0][:3], runs[1][:3]):
        assert abs(a - b) < 1e-5 * abs(b)
    assert float((runs[0][3] - runs[1][3]).abs().max()) < 1e-5


def test_device_prefetcher_keeps_unread_buffers():
    """`_train_step` draws two batches and only then reads them.  With the consumer stream stalled
    (the host runs far ahead of the GPU) no host->device copy may land in a buffer whose reads are not queued yet:
    every batch must arrive intact although the ring has only depth + 2 slots."""
    from littlegan_b200.dataset import DevicePrefetcher
    from littlegan_b200.eager_trainer import OutOfRangeError

    class _Numbered:
        def __init__(self, n):
            self.n, self.i = n, 0
            self.pool = [(torch.full((256, 1024), float(k)).pin_memory(), torch.full((4, 5), float(k)).pin_memory())
                         for k in range(n)]

        def get_next(self):
            if self.i >= self.n:
                raise OutOfRangeError()
            self.i += 1
            return self.pool[self.i - 1]

    n = 24
    pf = DevicePrefetcher(_Numbered(n), depth=2)
    torch.cuda._sleep(int(4e8))                       # ~0.2 s: every read below is queued behind this
    kept = []
    for step in range(n // 2):
        a = pf.get_next()
        b = pf.get_next()                             # second draw BEFORE the first batch is read
        kept.append((a[0].clone(), a[1].clone(), b[0].clone(), b[1].clone()))
    with pytest.raises(OutOfRangeError):
        pf.get_next()
    torch.cuda.synchronize()
    for step, (ai, ac, bi, bc) in enumerate(kept):
        for t, k in ((ai, 2 * step), (ac, 2 * step), (bi, 2 * step + 1), (bc, 2 * step + 1)):
            assert float(t.min()) == float(t.max()) == float(k), (step, k, float(t.min()), float(t.max()))
