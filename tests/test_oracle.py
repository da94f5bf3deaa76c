"""CPU tests of the oracle itself (no GPU): definition-level loops for the TF conv semantics,
closed forms for the norm backward and TF-Adam, the committed golden vectors, and the FID oracle
against NumPy/SciPy called directly."""
import json
import math
import os

import pytest

import numpy as np
import torch

from oracle import fid_oracle
from oracle import littlegan_oracle as O
from tests.util import build_product, product_args, product_weights_to_oracle, small_args

GOLD = os.path.join(os.path.dirname(__file__), "golden")
D = torch.float64


def _loop_conv(x, w, b, s, pad):
    N, H, W_, A = x.shape
    B = w.shape[3]
    y = torch.zeros(N, H // s, W_ // s, B, dtype=D)
    for i in range(H // s):
        for j in range(W_ // s):
            for ky in range(5):
                for kx in range(5):
                    yy, xx = s * i + ky - pad, s * j + kx - pad
                    if 0 <= yy < H and 0 <= xx < W_:
                        y[:, i, j] += x[:, yy, xx] @ w[ky, kx]
    return y + b


def _loop_convT(x, w, b, s, pad):
    N, H, W_, Bc = x.shape
    A = w.shape[2]
    y = torch.zeros(N, H * s, W_ * s, A, dtype=D)
    for i in range(H):
        for j in range(W_):
            for ky in range(5):
                for kx in range(5):
                    Y, X = s * i + ky - pad, s * j + kx - pad
                    if 0 <= Y < H * s and 0 <= X < W_ * s:
                        y[:, Y, X] += x[:, i, j] @ w[ky, kx].T
    return y + b


def test_conv_definitions():
    g = torch.Generator().manual_seed(0)
    for s, pad in ((2, 1), (1, 2)):
        x = torch.randn(2, 8, 8, 3, generator=g, dtype=D)
        w = torch.randn(5, 5, 3, 4, generator=g, dtype=D)
        b = torch.randn(4, generator=g, dtype=D)
        assert (O.conv2d_same(x, w, b, s) - _loop_conv(x, w, b, s, pad)).abs().max() < 1e-12
        xs = torch.randn(2, 4, 4, 4, generator=g, dtype=D)
        bt = torch.randn(3, generator=g, dtype=D)
        assert (O.conv2d_transpose_same(xs, w, bt, s) - _loop_convT(xs, w, bt, s, pad)).abs().max() < 1e-12


def test_transpose_is_adjoint_of_conv():
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 16, 16, 5, generator=g, dtype=D)
    y = torch.randn(2, 8, 8, 7, generator=g, dtype=D)
    w = torch.randn(5, 5, 5, 7, generator=g, dtype=D)
    z0, z5 = torch.zeros(7, dtype=D), torch.zeros(5, dtype=D)
    lhs = (O.conv2d_same(x, w, z0, 2) * y).sum()
    rhs = (x * O.conv2d_transpose_same(y, w, z5, 2)).sum()
    assert abs(float(lhs - rhs)) < 1e-9 * abs(float(lhs))


def test_instance_norm_closed_form_backward():
    """dx = (gamma/s)(g - mean g - xhat (s/sigma) mean(g xhat)), eps on the std."""
    g_ = torch.Generator().manual_seed(2)
    x = torch.randn(3, 4, 4, 6, generator=g_, dtype=D, requires_grad=True)
    gamma = torch.tensor([1.7], dtype=D, requires_grad=True)
    beta = torch.tensor([0.3], dtype=D, requires_grad=True)
    y = O.instance_norm(x, gamma, beta)
    gout = torch.randn(y.shape, generator=g_, dtype=D)
    dx, dg, db = torch.autograd.grad((y * gout).sum(), [x, gamma, beta])
    xd = x.detach()
    mu = xd.mean(dim=(1, 2, 3), keepdim=True)
    sigma = ((xd - mu) ** 2).mean(dim=(1, 2, 3), keepdim=True).sqrt()
    s = sigma + 1e-3
    xh = (xd - mu) / s
    mg = gout.mean(dim=(1, 2, 3), keepdim=True)
    mgx = (gout * xh).mean(dim=(1, 2, 3), keepdim=True)
    dx_cf = gamma.detach() / s * (gout - mg - xh * (s / sigma) * mgx)
    assert (dx - dx_cf).abs().max() < 1e-12
    assert abs(float(dg - (gout * xh).sum())) < 1e-10 and abs(float(db - gout.sum())) < 1e-10
    # mean ~ 0, std slightly below 1 because eps is added to the std, not the variance
    y0 = O.instance_norm(xd, torch.ones(1, dtype=D), torch.zeros(1, dtype=D))
    assert abs(float(y0[0].mean())) < 1e-12
    assert abs(float(y0[0].std(unbiased=False)) - float(sigma[0] / s[0])) < 1e-12


def test_bce_matches_keras_form():
    t = torch.tensor([[0.98, -0.94], [0.02, 0.98]], dtype=D)
    p = torch.tensor([[0.3, 0.9], [1.0, 0.0]], dtype=D)
    pc = p.clamp(1e-7, 1 - 1e-7)
    want = (-(t * torch.log(pc + 1e-7) + (1 - t) * torch.log(1 - pc + 1e-7))).mean(-1).mean()
    assert abs(float(O.bce(t, p) - want)) < 1e-9
    pr = p.clone().requires_grad_(True)
    (gr,) = torch.autograd.grad(O.bce(t, pr), pr)
    assert float(gr[0, 0]) != 0.0 and float(gr[1, 0]) == 0.0   # p = 1.0 lies outside [1e-7, 1-1e-7]
    pr2 = torch.tensor([[1.0 + 1e-3, -1e-3]], dtype=D, requires_grad=True)
    (g2,) = torch.autograd.grad(O.bce(t[:1], pr2), pr2)
    assert float(g2.abs().max()) == 0.0    # outside the clip range the gradient is zero


def test_tf_adam_form():
    opt = O.TFAdam(0.1, 0.5, 0.9)
    p = torch.tensor([1.0], dtype=D)
    g = torch.tensor([0.2], dtype=D)
    opt.apply([(g, p)])
    m, v = 0.5 * 0.2, 0.1 * 0.04
    lr_t = 0.1 * math.sqrt(1 - 0.9) / (1 - 0.5)
    assert abs(float(p) - (1.0 - lr_t * m / (math.sqrt(v) + 1e-8))) < 1e-12
    opt.apply([])                         # a step that trains other variables still advances t
    assert opt.t == 2


def test_partition_schedule():
    a = small_args(use_partition=True)
    tr = O.OracleTrainer(a, O.init_weights(a))
    assert tr._train_idx("G", 3) == list(range(22))
    assert tr._train_idx("G", 5) == [4, 5, 6, 7]          # (5//5)%3 = 1
    assert tr._train_idx("D", 10) == [16, 17, 18, 19]      # (10//5)%3 = 2
    assert tr._train_idx("D", 15) == list(range(12))
    assert tr._train_idx("A", 15) == [0, 1, 2, 3]


def test_parameter_counts_match_survey():
    a = O.make_args(cond_dim=7)
    W = O.init_weights(a)
    assert sum(w.numel() for w in W["D"]) == 3683856
    assert sum(w.numel() for w in W["G"]) == 6017869
    assert sum(w.numel() for w in W["A"]) == 196610
    a = O.make_args(cond_dim=40)
    W = O.init_weights(a)
    assert [sum(w.numel() for w in W[k]) for k in "DGA"] == [4494897, 6828877, 1007618]


def test_golden_small_step_is_reproduced():
    """The committed fixture pins the oracle: any drift of oracle/ shows up here."""
    gold = np.load(os.path.join(GOLD, "small_step.npz"))
    oargs = small_args(use_partition=True)
    gen, disc, adj = build_product(product_args(oargs, dtype="fp32"), seed=0)
    ot = O.OracleTrainer(oargs, product_weights_to_oracle(gen, disc, adj), dtype=torch.float64)
    r = ot.train_step(11, *O.synthetic_batch(oargs, 4, seed=5), return_grads=True)
    assert np.abs(r["fake_image"].numpy() - gold["fake_image"]).max() < 1e-6
    assert np.abs(r["adj_image"].numpy() - gold["adj_image"]).max() < 1e-6
    got = np.array([float(r["gen_loss"]), float(r["disc_loss"]), float(r["adj_loss"])])
    assert np.abs(got - gold["losses"]).max() < 1e-9
    for key in "DGA":
        gs = np.array([float(g.sum()) for g in r["grads"][key].values()])
        assert np.allclose(gs, gold["grad_sum_" + key], rtol=1e-7, atol=1e-9)


def test_golden_trajectory_prefix_is_reproduced():
    gold = json.load(open(os.path.join(GOLD, "trajectory_full.json")))
    assert len(gold["gen"]) == 100 and gold["adj"][9] is None and gold["adj"][10] is not None
    oargs = O.make_args(**gold["args"])
    gen, disc, adj = build_product(product_args(oargs, dtype="fp32"), seed=gold["seed"])
    ot = O.OracleTrainer(oargs, product_weights_to_oracle(gen, disc, adj), dtype=torch.float32)
    r = ot.train_step(1, *O.synthetic_batch(oargs, 4, seed=gold["data_seed"] + 1))
    assert abs(float(r["gen_loss"]) - gold["gen"][0]) < 1e-4 * gold["gen"][0]
    assert abs(float(r["disc_loss"]) - gold["disc"][0]) < 1e-4 * gold["disc"][0]


def test_fid_oracle_is_numpy_scipy():
    rng = np.random.RandomState(0)
    act = rng.randn(300, 16) @ rng.randn(16, 16) + 2.0
    mu, sigma = fid_oracle.activation_statistics(act)
    assert np.allclose(mu, act.mean(0)) and np.allclose(sigma, np.cov(act, rowvar=False))
    assert fid_oracle.used_rows(1037, 100) == 1000 and fid_oracle.used_rows(30, 50) == 30
    act2 = rng.randn(200, 16) * 0.5 + 1.0
    mu2, sig2 = fid_oracle.activation_statistics(act2)
    d = fid_oracle.frechet_distance(mu, sigma, mu2, sig2)
    assert d > 0 and abs(fid_oracle.frechet_distance(mu, sigma, mu, sigma)) < 1e-6
    # symmetric-form cross-check of Tr sqrtm(s1 s2)
    w, v = np.linalg.eigh(sigma)
    root = (v * np.sqrt(np.clip(w, 0, None))) @ v.T
    tr = np.sqrt(np.clip(np.linalg.eigvalsh(root @ sig2 @ root), 0, None)).sum()
    alt = ((mu - mu2) ** 2).sum() + np.trace(sigma) + np.trace(sig2) - 2 * tr
    assert abs(alt - d) < 1e-8 * abs(d)


def test_augment_restatement_against_colorsys_and_definitions():
    """oracle.augment (eager_trainer.py:127-131): the hue rotation against Python's colorsys on [0,1] pixels and
    its invariance under x -> 2x - 1 (the reference feeds [-1,1] images); flip / brightness / contrast by definition."""
    import colorsys
    g = torch.Generator().manual_seed(0)
    x = torch.rand(40, 3, generator=g, dtype=torch.float64)
    for d in (0.03, -0.03, 0.31, -0.77):
        y = O.adjust_hue(x, d)
        for i in range(x.shape[0]):
            h, s, v = colorsys.rgb_to_hsv(*x[i].tolist())
            want = colorsys.hsv_to_rgb((h + d) % 1.0, s, v)
            assert max(abs(a - b) for a, b in zip(want, y[i].tolist())) < 1e-12
        assert float((O.adjust_hue(2 * x - 1, d) - (2 * y - 1)).abs().max()) < 1e-12
    grey = torch.full((2, 3), 0.4, dtype=torch.float64)
    assert torch.equal(O.adjust_hue(grey, 0.2), grey)
    img = torch.rand(3, 4, 6, 3, generator=g, dtype=torch.float64) * 2 - 1
    flips = torch.tensor([1.0, 0.0, 1.0], dtype=torch.float64)
    zero = torch.zeros_like(img)
    out = O.augment(img, flips, 0.01, 0.9, 0.0, zero)
    for n in range(3):
        src = img[n].flip(1) if flips[n] > 0.5 else img[n]
        m = (src + 0.01).mean(dim=(0, 1), keepdim=True)
        assert float((out[n] - (((src + 0.01) - m) * 0.9 + m)).abs().max()) < 1e-14


def test_inception_oracle_wiring_matches_torchvision():
    """The Inception oracle is pinned against torchvision's inception_v3 (same published architecture) with
    identical weights; the FID variant then changes only the three documented pooling details."""
    import pytest
    torchvision = pytest.importorskip("torchvision")
    from oracle import inception_oracle as IO
    m = torchvision.models.inception_v3(weights=None, aux_logits=False, transform_input=False, init_weights=False).eval()
    W = IO.random_weights(seed=1)
    g = torch.Generator().manual_seed(0)
    sd = m.state_dict()
    for name, p in W.items():
        p["gamma"] = torch.rand(p["beta"].shape, generator=g) * 0.5 + 0.75
        sd[name + ".conv.weight"].copy_(p["W"].permute(3, 2, 0, 1))
        sd[name + ".bn.weight"].copy_(p["gamma"])
        sd[name + ".bn.bias"].copy_(p["beta"])
        sd[name + ".bn.running_mean"].copy_(p["mean"])
        sd[name + ".bn.running_var"].copy_(p["var"])
    assert len(W) == 94 and sum(k.endswith("conv.weight") for k in sd) == 94
    m.load_state_dict(sd)
    x = torch.randn(1, 3, 299, 299, generator=g)
    got = {}
    m.avgpool.register_forward_hook(lambda mod, i, o: got.__setitem__("pool", o.flatten(1)))
    with torch.no_grad():
        m(x)
        f = IO.InceptionOracle(W, variant="torchvision", dtype=torch.float32).features(x)
        f_fid = IO.InceptionOracle(W, variant="fid", dtype=torch.float32).features(x)
    assert f.shape == (1, 2048)
    assert float((f - got["pool"]).abs().max() / got["pool"].abs().max()) < 1e-5
    assert float((f - f_fid).abs().max()) > 1e-2          # max pool in Mixed_7c, pad-excluding averages


def test_inception_resize_is_tf1_bilinear():
    """TF-1.x ResizeBilinear (align_corners=False): src = dst * in/out; identity at equal size, exact 2x up-sample
    values, last row/column clamped."""
    from oracle import inception_oracle as IO
    x = torch.arange(12, dtype=torch.float64).reshape(1, 3, 4, 1)
    assert torch.equal(IO.resize_bilinear_tf1(x, 3, 4), x)
    y = IO.resize_bilinear_tf1(x, 6, 8)
    assert torch.allclose(y[0, 0, :, 0], torch.tensor([0, 0.5, 1, 1.5, 2, 2.5, 3, 3], dtype=torch.float64))
    assert torch.allclose(y[0, :, 0, 0], torch.tensor([0, 2, 4, 6, 8, 8], dtype=torch.float64))


# ------------------------------------------------------------------------------------------------
# reference-held goldens (written by scripts/make_tf_goldens.py where TensorFlow 1.15 exists)
# ------------------------------------------------------------------------------------------------
_TF_STEP = os.path.join(os.path.dirname(__file__), "golden", "tf_step.npz")
_TF_TRAJ = os.path.join(os.path.dirname(__file__), "golden", "tf_trajectory.json")


@pytest.mark.skipif(not os.path.isfile(_TF_STEP), reason="tests/golden/tf_step.npz absent: TensorFlow 1.15 is not "
                    "installable here; a maintainer runs scripts/make_tf_goldens.py once (oracle stays unpinned)")
def test_oracle_matches_tf_goldens():
    """Pins the train-step oracle to the UNMODIFIED reference on TF 1.15: per-layer activations, losses, every
    gradient and the updated weights of one step at batch_no 11 (adjuster on) and 15 (partition step)."""
    from tests.util import small_args
    z = np.load(_TF_STEP, allow_pickle=False)
    t = lambda a: torch.from_numpy(np.asarray(a))
    oargs = small_args(use_partition=True)
    W = {k: [t(z["w_%s_%d" % (k, j)]) for j in range(n)] for k, n in (("D", 20), ("G", 22), ("A", 4))}
    i1, c1, i2, c2, noise = (t(z[k]) for k in ("i1", "c1", "i2", "c2", "noise"))
    rel = lambda got, want: float((got.double() - want.double()).abs().max() / want.double().abs().max().clamp_min(1e-30))
    for batch_no in (11, 15):
        p = "b%d_" % batch_no
        Wd = {k: [w.double() for w in v] for k, v in W.items()}
        with torch.no_grad():
            for i, m in enumerate(O.encoder(oargs, i1.double(), Wd["D"][:16])):
                assert rel(m, t(z[p + "enc%d_real1" % (i + 1)])) < 1e-4
            assert rel(O.generator(oargs, noise.double(), c2.double(), Wd["G"]), t(z[p + "fake0"])) < 1e-4
            pr, c = O.discriminator(oargs, i1.double(), Wd["D"])
            assert rel(pr, t(z[p + "pr_real1"])) < 1e-4 and rel(c, t(z[p + "c_real1"])) < 1e-4
            adj0 = O.adjuster(oargs, i1.double(), ((c2 + 1) * 0.5).double(), Wd["D"], Wd["G"], Wd["A"])
            assert rel(adj0, t(z[p + "adj0"])) < 1e-4
        ot = O.OracleTrainer(oargs, W, dtype=torch.float64)
        ref = ot.train_step(batch_no, i1, c1, i2, c2, noise, return_grads=True)
        assert rel(ref["fake_image"], t(z[p + "fake_image"])) < 1e-4
        assert rel(ref["adj_image"], t(z[p + "adj_image"])) < 1e-4
        for got, want in zip((ref["gen_loss"], ref["disc_loss"], ref["adj_loss"]), z[p + "losses"]):
            assert abs(float(got) - float(want)) < 1e-4 * abs(float(want))
        for key in "DGA":
            idx = sorted(ref["grads"][key])              # the partition group of this step, in weight order
            for j, i in enumerate(idx):
                want = t(z[p + "grad_%s_%d" % (key, j)])
                got = ref["grads"][key][i]
                if want.numel() == 1:
                    assert abs(float(got) - float(want)) < 1e-4 * max(abs(float(want)), 1e-1), (key, i)
                else:
                    assert rel(got, want) < 1e-4, (batch_no, key, i)
        step = 1.6 * oargs.lr * (1 - 0.9) ** 0.5 / (1 - 0.5)         # size of the first TF-Adam step
        for key, n in (("D", 20), ("G", 22), ("A", 4)):
            for j in range(n):
                diff = (ot.W[key][j].detach().double() - t(z[p + "new_%s_%d" % (key, j)]).double()).abs()
                assert float(diff.max()) < 2.5 * step, (key, j)      # at most a sign flip of a noise-level gradient
                if diff.numel() > 1:
                    assert float((diff > 0.05 * step).double().mean()) < 5e-3, (key, j)


@pytest.mark.skipif(not os.path.isfile(_TF_TRAJ), reason="tests/golden/tf_trajectory.json absent (see above)")
def test_oracle_trajectory_matches_tf():
    """100 free-running steps: within 1% of TensorFlow's losses until the TF-Adam chaos sets in (first 10 steps),
    and a median deviation over all steps no larger than two fp32 implementations show against each other."""
    import json
    from tests.util import small_args
    tr = json.load(open(_TF_TRAJ))
    oargs = small_args(use_partition=True)
    ot = O.OracleTrainer(oargs, O.init_weights(oargs, tr["weight_seed"]), dtype=torch.float32)
    devs = []
    for b in range(1, len(tr["gen"]) + 1):
        i1, c1, i2, c2, noise = O.synthetic_batch(oargs, tr["B"], seed=tr["data_seed"] + b)
        ref = ot.train_step(b, i1, c1, i2, c2, noise)
        for got, want in ((ref["gen_loss"], tr["gen"][b - 1]), (ref["disc_loss"], tr["disc"][b - 1])):
            d = abs(float(got) - want) / abs(want)
            devs.append(d)
            if b <= 10:
                assert d < 0.01, (b, float(got), want)
    assert float(np.median(devs)) < 0.02


# ------------------------------------------------------------------------------------------------
# helpers the bf16-mode parity tests rely on (tests/test_parity_bf16_gpu.py)
# ------------------------------------------------------------------------------------------------
def test_bf16_storage_emulation_rounds_value_and_gradient():
    """`store="bf16"`: the stored map AND the gradient flowing back through it are rounded to bf16; the default
    oracle is exact.  The emulation stays a small perturbation of the forward pass and leaves its structure alone."""
    from tests.util import small_args
    x = torch.tensor([1.0 + 2 ** -10, -3.3, 0.1], dtype=torch.float64, requires_grad=True)
    y = O._RoundBF16.apply(x)
    assert torch.equal(y.detach(), x.detach().to(torch.bfloat16).double())
    g = torch.tensor([1.0 + 2 ** -10, 0.3, -7.77], dtype=torch.float64)
    y.backward(g)
    assert torch.equal(x.grad, g.to(torch.bfloat16).double())
    x2 = torch.tensor([0.123456789], dtype=torch.float64, requires_grad=True)
    z = O._RoundGradBF16.apply(x2)
    assert torch.equal(z.detach(), x2.detach())                  # value untouched
    z.backward(torch.tensor([0.123456789], dtype=torch.float64))
    assert torch.equal(x2.grad, torch.tensor([0.123456789]).to(torch.bfloat16).double())

    oargs, eargs = small_args(), small_args(store="bf16")
    W = O.init_weights(oargs, 0)
    i1, c1, i2, c2, noise = O.synthetic_batch(oargs, 4, seed=5)
    t0, t1 = {}, {}
    r0 = O.OracleTrainer(oargs, W, dtype=torch.float64).train_step(11, i1, c1, i2, c2, noise, taps=t0)
    r1 = O.OracleTrainer(eargs, W, dtype=torch.float64).train_step(11, i1, c1, i2, c2, noise, taps=t1)
    assert set(t0) == set(t1) and len(t0) >= 34
    for k in ("gen_loss", "disc_loss", "adj_loss"):
        assert 0 < abs(float(r0[k]) - float(r1[k])) < 1e-2 * abs(float(r0[k]))
    for k in ("g_dec4", "dr_enc4", "a_dec4", "da_enc4"):
        d = float((t0[k] - t1[k]).abs().max() / t0[k].abs().max())
        assert 0 < d < 5e-2, (k, d)
        # every stored map of the emulating oracle is bf16-representable
        assert torch.equal(t1[k], t1[k].to(torch.bfloat16).to(t1[k].dtype))


def test_forward_losses_equal_the_train_step_losses():
    """OracleTrainer.forward_losses (forward only) reproduces the three losses of train_step and reports each
    loss's scale = the sum of the absolute values of its terms."""
    from tests.util import small_args
    oargs = small_args(use_partition=True)
    W = O.init_weights(oargs, 0)
    ot = O.OracleTrainer(oargs, W, dtype=torch.float64)
    for b in (3, 11):
        i1, c1, i2, c2, noise = O.synthetic_batch(oargs, 4, seed=20 + b)
        fl = ot.forward_losses(b, i1, c1, i2, c2, noise)
        ref = ot.train_step(b, i1, c1, i2, c2, noise)
        for k in ("gen", "disc", "adj"):
            if ref[k + "_loss"] is None:
                assert fl[k + "_loss"] is None and fl[k + "_scale"] is None
                continue
            assert abs(fl[k + "_loss"] - float(ref[k + "_loss"])) < 1e-12
            assert fl[k + "_scale"] >= abs(fl[k + "_loss"]) - 1e-12
