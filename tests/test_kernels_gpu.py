"""Per-kernel parity against the CPU oracle (fp64), through the C-ABI.  -m gpu."""
import numpy as np
import pytest
import torch

from oracle import fid_oracle
from oracle import littlegan_oracle as O
from tests.util import rel_err, tol

pytestmark = pytest.mark.gpu

DTYPES = [torch.float32, torch.bfloat16]
# (N, Hb, Wb, A, B, stride)
GEOMS = [
    (2, 16, 16, 3, 8, 2),      # 3-channel edge layer
    (3, 8, 8, 16, 24, 2),
    (5, 4, 4, 8, 48, 2),       # tiny maps: one tile spans several samples
    (1, 12, 12, 5, 7, 1),      # stride 1, odd channel counts
    (2, 16, 16, 3, 32, 1),     # final-conv geometry
    (2, 32, 32, 64, 128, 2),   # enc2 / dec3 shape at reduced resolution
]


def _rand(shape, seed, dtype, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    x = (torch.randn(shape, generator=g, dtype=torch.float64) * scale)
    return x.to(dtype)           # rounded to the storage type; the oracle sees the rounded values


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("geom", GEOMS)
def test_conv_fprop(geom, dtype):
    from littlegan_b200 import kernels as K
    N, Hb, Wb, A, B, s = geom
    x = _rand((N, Hb, Wb, A), 1, dtype)
    W = _rand((5, 5, A, B), 2, torch.float32, 0.1)
    b = _rand((B,), 3, torch.float32)
    ref = O.conv2d_same(x.double(), W.double(), b.double(), s)
    out = torch.empty(N, Hb // s, Wb // s, B, dtype=dtype, device="cuda")
    stats = torch.zeros(N, 2, dtype=torch.float64, device="cuda")
    K.conv2d_fprop(x.cuda(), W.cuda(), b.cuda(), out, stats, s)
    assert rel_err(out, ref) < tol(dtype)
    ref_stats = torch.stack([ref.reshape(N, -1).sum(1), (ref.reshape(N, -1) ** 2).sum(1)], 1)
    assert rel_err(stats, ref_stats) < 1e-4


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("act", [0, 1])
@pytest.mark.parametrize("geom", GEOMS)
def test_conv_dgrad(geom, act, dtype):
    from littlegan_b200 import kernels as K
    N, Hb, Wb, A, B, s = geom
    x = _rand((N, Hb // s, Wb // s, B), 4, dtype)
    W = _rand((5, 5, A, B), 5, torch.float32, 0.1)
    b = _rand((A,), 6, torch.float32)
    pre = O.conv2d_transpose_same(x.double(), W.double(), b.double(), s)
    ref = torch.tanh(pre) if act else pre
    out = torch.empty(N, Hb, Wb, A, dtype=dtype, device="cuda")
    stats = torch.zeros(N, 2, dtype=torch.float64, device="cuda")
    K.conv2d_dgrad(x.cuda(), W.cuda(), b.cuda(), out, stats, s, act)
    assert rel_err(out, ref) < tol(dtype)
    ref_stats = torch.stack([pre.reshape(N, -1).sum(1), (pre.reshape(N, -1) ** 2).sum(1)], 1)
    assert rel_err(stats, ref_stats) < 1e-4


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("geom", GEOMS)
def test_conv_wgrad(geom, dtype):
    from littlegan_b200 import kernels as K
    N, Hb, Wb, A, B, s = geom
    big = _rand((N, Hb, Wb, A), 7, dtype)
    small = _rand((N, Hb // s, Wb // s, B), 8, dtype)
    W = torch.zeros(5, 5, A, B, dtype=torch.float64, requires_grad=True)
    y = O.conv2d_same(big.double(), W, torch.zeros(B, dtype=torch.float64), s)
    (ref,) = torch.autograd.grad((y * small.double()).sum(), W)
    dW = torch.zeros(5, 5, A, B, dtype=torch.float32, device="cuda")
    K.conv2d_wgrad(big.cuda(), small.cuda(), dW, s)
    assert rel_err(dW, ref) < 1e-4          # inputs are shared exactly; accumulation is fp32
    K.conv2d_wgrad(big.cuda(), small.cuda(), dW, s)   # accumulates
    assert rel_err(dW, 2 * ref) < 1e-4


def test_conv_adjoint_identity():
    """<fprop(x), y> == <x, dgrad(y)> - size-independent property, at a full-size layer."""
    from littlegan_b200 import kernels as K
    N, Hb, Wb, A, B, s = 2, 64, 64, 64, 128, 2
    x = _rand((N, Hb, Wb, A), 11, torch.float32).cuda()
    y = _rand((N, Hb // s, Wb // s, B), 12, torch.float32).cuda()
    W = _rand((5, 5, A, B), 13, torch.float32, 0.05).cuda()
    fx = torch.empty_like(y)
    K.conv2d_fprop(x, W, None, fx, None, s)
    dy = torch.empty_like(x)
    K.conv2d_dgrad(y, W, None, dy, None, s)
    lhs = (fx.double() * y.double()).sum()
    rhs = (x.double() * dy.double()).sum()
    scale = (fx.double() * y.double()).abs().sum()       # the inner product itself nearly cancels
    assert abs(float(lhs - rhs)) / float(scale) < 1e-6


@pytest.mark.parametrize("zdt,odt", [(torch.float32, torch.float32), (torch.bfloat16, torch.bfloat16),
                                     (torch.float32, torch.bfloat16)])
@pytest.mark.parametrize("shape", [(3, 8, 8, 16), (2, 4, 4, 6), (4, 1030)])
@pytest.mark.parametrize("alphas", [(1.0, 0.3), (0.3, 1.0)])
def test_instnorm_fwd_bwd(shape, alphas, zdt, odt):
    from littlegan_b200 import kernels as K
    a_pre, a_post = alphas
    N = shape[0]
    z = _rand(shape, 20, zdt, 2.0) + 0.5
    z = z.to(zdt)
    skip = _rand(shape, 21, odt)
    gout = _rand(shape, 22, odt)
    gamma = torch.tensor([1.3]); beta = torch.tensor([-0.2])
    zr = z.double().requires_grad_(True)
    gr = gamma.double().requires_grad_(True); br = beta.double().requires_grad_(True)
    u = torch.nn.functional.leaky_relu(zr, a_pre)
    ref = torch.nn.functional.leaky_relu(O.instance_norm(u, gr, br), a_post) + skip.double()
    dz_ref, dg_ref, db_ref = torch.autograd.grad((ref * gout.double()).sum(), [zr, gr, br])

    zc = z.cuda()
    stats = torch.zeros(N, 2, dtype=torch.float64, device="cuda")
    K.rowstats(zc, stats, a_pre)
    out = torch.empty(shape, dtype=odt, device="cuda")
    K.instnorm_act_fwd(zc, stats, gamma.cuda(), beta.cuda(), skip.cuda(), out, 1e-3, a_pre, a_post)
    assert rel_err(out, ref) < tol(odt)

    red = torch.zeros(N, 2, dtype=torch.float64, device="cuda")
    dz = torch.empty_like(zc)
    dgam = torch.zeros(1, device="cuda"); dbet = torch.zeros(1, device="cuda")
    K.instnorm_act_bwd(gout.cuda(), zc, stats, gamma.cuda(), beta.cuda(), red, dz, dgam, dbet, 1e-3, a_pre, a_post)
    assert rel_err(dz, dz_ref) < tol(zdt) * 2
    assert rel_err(dgam, dg_ref) < 1e-3 and rel_err(dbet, db_ref) < 1e-3


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("rows,C", [(512, 64), (100, 384), (8192, 384), (1001, 192), (333, 3), (64, 32), (7, 1), (6, 24576)])
def test_bias_grad(rows, C, dtype):
    from littlegan_b200 import kernels as K
    g = _rand((rows, C), 30, dtype)
    db = torch.zeros(C, device="cuda")
    K.bias_grad(g.cuda(), db)
    assert rel_err(db, g.double().sum(0)) < 1e-4


@pytest.mark.parametrize("tA,tB,acc", [(False, False, False), (True, False, True), (False, True, True),
                                       (False, False, True)])
@pytest.mark.parametrize("adt", DTYPES)
def test_gemm(tA, tB, acc, adt):
    from littlegan_b200 import kernels as K
    M, N, Kd = 37, 70, 1000
    A = _rand((Kd, M) if tA else (M, Kd), 40, adt)
    Bm = _rand((N, Kd) if tB else (Kd, N), 41, torch.float32)
    bias = None if acc else _rand((N,), 42, torch.float32)
    ref = (A.double().T if tA else A.double()) @ (Bm.double().T if tB else Bm.double())
    if bias is not None:
        ref = ref + bias.double()
    C = torch.zeros(M, N, device="cuda")
    K.gemm(A.cuda(), Bm.cuda(), C, M, N, Kd, bias=None if bias is None else bias.cuda(), transA=tA, transB=tB,
           accumulate=acc)
    assert rel_err(C, ref) < 1e-4


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("N,F,U1", [(5, 128, 7), (128, 1536, 40), (200, 512, 47), (64, 24576, 40)])
def test_dense_heads_fwd_bwd(N, F, U1, dtype):
    """Discriminator heads (model.py:62-63,70-72) as one fused pass; repeated calls share a workspace."""
    from littlegan_b200 import kernels as K
    f = _rand((N, F), 50, dtype, 0.5)
    W0 = _rand((F, 1), 51, torch.float32, 0.05)
    W1 = _rand((F, U1), 52, torch.float32, 0.05)
    b0 = _rand((1,), 53, torch.float32)
    b1 = _rand((U1,), 54, torch.float32)
    dl0 = _rand((N, 1), 55, torch.float32)
    dl1 = _rand((N, U1), 56, torch.float32)
    fd = f.double()
    ref0 = torch.sigmoid(fd @ W0.double() + b0.double())
    ref1 = torch.sigmoid(fd @ W1.double() + b1.double())
    ws = K.dense_heads_workspace(N, "cuda")
    fc, W0c, W1c = f.cuda(), W0.cuda(), W1.cuda()
    for _ in range(3):                                   # the kernel must leave the workspace zeroed
        o0 = torch.full((N, 1), 7.0, device="cuda")
        o1 = torch.full((N, U1), 7.0, device="cuda")
        K.dense_heads_fwd(fc, W0c, b0.cuda(), W1c, b1.cuda(), o0, o1, K.ACT_SIGMOID, ws)
        assert rel_err(o0, ref0) < 1e-4 and rel_err(o1, ref1) < 1e-4
    assert int(ws.to(torch.int32).abs().sum()) == 0

    ref_df = dl0.double() @ W0.double().T + dl1.double() @ W1.double().T
    df = torch.empty(N, F, dtype=dtype, device="cuda")
    dW0 = torch.ones(F, 1, device="cuda"); dW1 = torch.ones(F, U1, device="cuda")
    db0 = torch.ones(1, device="cuda"); db1 = torch.ones(U1, device="cuda")
    K.dense_heads_bwd(fc, dl0.cuda(), dl1.cuda(), W0c, W1c, df, dW0, dW1, db0, db1)
    assert rel_err(df, ref_df) < tol(dtype)
    assert rel_err(dW0, 1 + fd.T @ dl0.double()) < 1e-4
    assert rel_err(dW1, 1 + fd.T @ dl1.double()) < 1e-4
    assert rel_err(db0, 1 + dl0.double().sum(0)) < 1e-4
    assert rel_err(db1, 1 + dl1.double().sum(0)) < 1e-4
    # input-gradient only (generator / adjuster steps), one head's gradient absent
    df2 = torch.empty(N, F, dtype=dtype, device="cuda")
    K.dense_heads_bwd(None, None, dl1.cuda(), W0c, W1c, df2, None, None, None, None)
    assert rel_err(df2, dl1.double() @ W1.double().T) < tol(dtype)


def test_bce_and_l1():
    from littlegan_b200 import kernels as K
    rows, cols = 6, 5
    logits = _rand((rows, cols), 50, torch.float32, 3.0)
    logits[0, 0] = 40.0          # saturates the sigmoid: clipped region, zero gradient
    logits[1, 1] = -40.0
    t = O.soft((torch.rand(rows, cols) < 0.5).float() * 2 - 1)
    lr = logits.clone().requires_grad_(True)
    ref = O.bce(t, torch.sigmoid(lr)) * 2.0
    (dref,) = torch.autograd.grad(ref, lr)
    p = torch.sigmoid(logits).cuda()
    acc = torch.zeros(1, device="cuda"); dl = torch.empty(rows, cols, device="cuda")
    K.bce_sigmoid(p, t.cuda(), 2.0, acc, dl)
    assert abs(float(acc) - float(ref)) / abs(float(ref)) < 1e-5
    assert float((dl.cpu() - dref).abs().max()) < 1e-6
    acc.zero_()
    K.bce_sigmoid(p, 0.98, 1.0, acc, None)
    assert abs(float(acc) - float(O.bce(torch.full((rows, cols), 0.98), torch.sigmoid(logits)))) < 1e-5

    for dtype in DTYPES:
        y = torch.tanh(_rand((2, 8, 8, 3), 51, torch.float32)).to(dtype)
        tgt = _rand((2, 8, 8, 3), 52, dtype)
        gin = _rand((2, 8, 8, 3), 53, dtype)
        pre = torch.atanh(y.double().clamp(-0.999, 0.999)).requires_grad_(True)
        yy = torch.tanh(pre)
        loss = 0.02 * (tgt.double() - yy).abs().mean()
        (dref,) = torch.autograd.grad(loss + (yy * gin.double()).sum(), pre)
        acc = torch.zeros(1, device="cuda"); dpre = torch.empty(2, 8, 8, 3, dtype=dtype, device="cuda")
        K.l1_tanh_bwd(y.cuda(), tgt.cuda(), gin.cuda(), dpre, 0.02, acc)
        assert abs(float(acc) - float(0.02 * (tgt.double() - y.double()).abs().mean())) < 1e-5
        ok = y.double().abs() < 0.998
        assert rel_err(dpre.cpu()[ok], dref[ok]) < tol(dtype)


@pytest.mark.parametrize("n,off", [(1000, 0), (1003, 0), (1000, 1), (3, 0), (700001, 0)])
def test_adam_matches_tf_form(n, off):
    """16-byte vector path (aligned ranges, n % 4 tail), the scalar path (off = 1: misaligned views) and a range
    that spans several grid passes."""
    from littlegan_b200 import kernels as K
    p0 = _rand((n,), 60, torch.float32);
    opt = O.TFAdam(5e-5, 0.5, 0.9)
    p_ref = p0.clone().double()
    buf = lambda src=None: (torch.zeros(n + off, device="cuda")[off:] if src is None else
                            torch.cat([torch.zeros(off), src]).cuda()[off:])
    p = buf(p0); m = buf(); v = buf()
    state = torch.zeros(4, dtype=torch.float64, device="cuda")
    for step in range(3):
        g = _rand((n,), 61 + step, torch.float32)
        opt.apply([(g.double().clamp(-0.5, 0.5), p_ref)])
        K.adam_advance(state, 5e-5, 0.5, 0.9)
        K.adam_apply(p, buf(g), m, v, state, 0.5, 0.9, 1e-8, 0.5)
    assert float(state[0]) == 3.0
    assert float((p.cpu().double() - p_ref).abs().max()) < 5e-7   # fp32 rounding of p ~ 1


def test_fid_statistics_and_distance():
    """Pinned row: np.mean / np.cov / scipy sqrtm are the reference's own arithmetic."""
    from littlegan_b200 import fid
    rng = np.random.RandomState(0)
    n, d = 1037, 200
    base = rng.randn(n, d).astype(np.float32) @ rng.randn(d, d).astype(np.float32) * 0.1 + 3.0
    mu_ref, sig_ref = fid_oracle.activation_statistics(base[: fid_oracle.used_rows(n, 100)])
    mu, sigma = fid.calculate_activation_statistics(torch.from_numpy(base), None, batch_size=100)
    assert np.abs(mu - mu_ref).max() < 1e-10
    assert np.abs(sigma - sig_ref).max() / np.abs(sig_ref).max() < 1e-10
    assert np.abs(sigma - sigma.T).max() == 0.0
    other = rng.randn(900, d).astype(np.float32) * 0.7 + 2.5
    mu2, sig2 = fid_oracle.activation_statistics(other)
    ref = fid_oracle.frechet_distance(mu_ref, sig_ref, mu2, sig2)
    got = fid.calculate_frechet_distance(mu, sigma, mu2, sig2)
    assert abs(got - ref) / abs(ref) < 1e-6
    assert abs(fid.calculate_frechet_distance(mu, sigma, mu, sigma)) < 1e-6 * np.trace(sig_ref)


# ------------------------------------------------------------------------------------------------
# tcgen05 / TMA path (bf16 operands): same oracle, weights pre-rounded to bf16 so that the only
# differences are fp32 accumulation order and the bf16 rounding of the stored output.
# ------------------------------------------------------------------------------------------------
TC_F = [
    (2, 64, 64, 64, 128, 2),     # enc2 geometry, SWIZZLE_128B
    (3, 16, 16, 256, 384, 2),    # enc4: tile spans 2 samples (3rd tile half empty), N split 2x192
    (2, 128, 128, 32, 64, 2),    # dec4 dgrad: KC=32 / SWIZZLE_64B
    (2, 32, 32, 64, 32, 1),      # stride-1 fprop
]
TC_T = [
    (2, 64, 64, 64, 128, 2, 0),
    (3, 16, 16, 256, 384, 2, 0),
    (2, 128, 128, 32, 64, 2, 0),
    (2, 128, 128, 3, 64, 2, 0),  # enc1 dgrad: 3 output channels padded to N=16
    (2, 128, 128, 3, 32, 1, 1),  # final conv: stride 1, tanh
]


def _pack(W):
    from littlegan_b200 import kernels as K
    A, B = W.shape[2], W.shape[3]
    buf = torch.empty(K.pack_conv_weights_bytes(A, B), dtype=torch.uint8, device="cuda")
    K.pack_conv_weights(W, buf)
    return buf


@pytest.mark.parametrize("geom", TC_F)
def test_tc_fprop(geom):
    from littlegan_b200 import kernels as K
    N, Hb, Wb, A, B, s = geom
    assert K.tc_supported(K.OP_FPROP, N, Hb, Wb, A, B, s)
    x = _rand((N, Hb, Wb, A), 1, torch.bfloat16)
    W = _rand((5, 5, A, B), 2, torch.bfloat16, 0.05).float()
    b = _rand((B,), 3, torch.float32)
    ref = O.conv2d_same(x.double(), W.double(), b.double(), s)
    out = torch.zeros(N, Hb // s, Wb // s, B, dtype=torch.bfloat16, device="cuda")
    stats = torch.zeros(N, 2, dtype=torch.float64, device="cuda")
    Wc = W.cuda()
    K.conv2d_fprop(x.cuda(), Wc, b.cuda(), out, stats, s, wpack=_pack(Wc), use_tc=True)
    torch.cuda.synchronize()
    assert rel_err(out, ref) < 1e-2
    ref_stats = torch.stack([ref.reshape(N, -1).sum(1), (ref.reshape(N, -1) ** 2).sum(1)], 1)
    assert float(((stats.cpu() - ref_stats).abs() / ref_stats.abs().max()).max()) < 1e-4


@pytest.mark.parametrize("geom", TC_T)
def test_tc_dgrad(geom):
    from littlegan_b200 import kernels as K
    N, Hb, Wb, A, B, s, act = geom
    assert K.tc_supported(K.OP_DGRAD, N, Hb, Wb, A, B, s)
    x = _rand((N, Hb // s, Wb // s, B), 4, torch.bfloat16)
    W = _rand((5, 5, A, B), 5, torch.bfloat16, 0.05).float()
    b = _rand((A,), 6, torch.float32)
    pre = O.conv2d_transpose_same(x.double(), W.double(), b.double(), s)
    ref = torch.tanh(pre) if act else pre
    out = torch.zeros(N, Hb, Wb, A, dtype=torch.bfloat16, device="cuda")
    stats = torch.zeros(N, 2, dtype=torch.float64, device="cuda")
    Wc = W.cuda()
    K.conv2d_dgrad(x.cuda(), Wc, b.cuda(), out, stats, s, act, wpack=_pack(Wc), use_tc=True)
    torch.cuda.synchronize()
    assert rel_err(out, ref) < 1e-2
    ref_stats = torch.stack([pre.reshape(N, -1).sum(1), (pre.reshape(N, -1) ** 2).sum(1)], 1)
    assert float(((stats.cpu() - ref_stats).abs() / ref_stats.abs().max()).max()) < 1e-4


@pytest.fixture
def forced_cta_pairs():
    """Run the generic tcgen05 conv kernels as CTA pairs (cta_group::2) even on these few-tile shapes."""
    from littlegan_b200 import kernels as K
    prev = K.set_cta_pairs(2)
    yield
    K.set_cta_pairs(prev)


@pytest.mark.parametrize("geom", TC_F + [(8, 32, 32, 128, 256, 2), (64, 16, 16, 256, 384, 2)])
def test_tc_fprop_cta_pairs(geom, forced_cta_pairs):
    test_tc_fprop(geom)


@pytest.mark.parametrize("geom", TC_T + [(8, 32, 32, 128, 256, 2, 0), (64, 16, 16, 256, 384, 2, 0)])
def test_tc_dgrad_cta_pairs(geom, forced_cta_pairs):
    test_tc_dgrad(geom)


@pytest.mark.parametrize("geom", [(3, 128, 128, 3, 64, 2), (2, 128, 128, 3, 32, 1)])
def test_tc_three_channel_layers_via_padding(geom):
    """enc1 fprop / wgrad and the final conv's input-gradient (fprop, stride 1) with the image padded
    3 -> 16 channels (KC = 16, SWIZZLE_32B)."""
    from littlegan_b200 import kernels as K
    N, Hb, Wb, A, B, s = geom
    x = _rand((N, Hb, Wb, A), 1, torch.bfloat16)
    W = _rand((5, 5, A, B), 2, torch.bfloat16, 0.1).float()
    b = _rand((B,), 3, torch.float32)
    small = _rand((N, Hb // s, Wb // s, B), 8, torch.bfloat16)
    Wr = W.double().requires_grad_(True)
    ref = O.conv2d_same(x.double(), Wr, b.double(), s)
    (dW_ref,) = torch.autograd.grad((ref * small.double()).sum(), Wr)
    xp = K.pad_channels(x.cuda(), torch.empty(N, Hb, Wb, 16, dtype=torch.bfloat16, device="cuda"))
    assert torch.equal(xp[..., :A].cpu(), x) and float(xp[..., A:].abs().max()) == 0.0
    assert K.tc_supported(K.OP_FPROP, N, Hb, Wb, 16, B, s) and K.tc_supported(K.OP_WGRAD, N, Hb, Wb, 16, B, s)
    out = torch.zeros(N, Hb // s, Wb // s, B, dtype=torch.bfloat16, device="cuda")
    stats = torch.zeros(N, 2, dtype=torch.float64, device="cuda")
    Wc = W.cuda()
    K.conv2d_fprop(xp, Wc, b.cuda(), out, stats, s, wpack=_pack(Wc), use_tc=True)
    assert rel_err(out, ref) < 1e-2
    ref_stats = torch.stack([ref.reshape(N, -1).sum(1), (ref.reshape(N, -1) ** 2).sum(1)], 1).detach()
    assert float(((stats.cpu() - ref_stats).abs() / ref_stats.abs().max()).max()) < 1e-4
    dW = torch.zeros(5, 5, A, B, dtype=torch.float32, device="cuda")
    K.conv2d_wgrad_padded(xp, small.cuda(), dW, s)
    assert rel_err(dW, dW_ref) < 1e-4


@pytest.mark.parametrize("geom", [(3, 128, 128, 3, 64, 2), (2, 128, 128, 3, 32, 1), (5, 64, 64, 3, 64, 2)])
def test_tc_cin3_fprop_wgrad(geom):
    """RGB-input layers on the in-shared-memory im2col kernels (image NOT padded)."""
    from littlegan_b200 import kernels as K
    N, Hb, Wb, A, B, s = geom
    assert K.tc_supported(K.OP_FPROP, N, Hb, Wb, A, B, s) and K.tc_supported(K.OP_WGRAD, N, Hb, Wb, A, B, s)
    x = _rand((N, Hb, Wb, A), 1, torch.bfloat16)
    W = _rand((5, 5, A, B), 2, torch.bfloat16, 0.1).float()
    b = _rand((B,), 3, torch.float32)
    small = _rand((N, Hb // s, Wb // s, B), 8, torch.bfloat16)
    Wr = W.double().requires_grad_(True)
    ref = O.conv2d_same(x.double(), Wr, b.double(), s)
    (dW_ref,) = torch.autograd.grad((ref * small.double()).sum(), Wr)
    out = torch.zeros(N, Hb // s, Wb // s, B, dtype=torch.bfloat16, device="cuda")
    stats = torch.zeros(N, 2, dtype=torch.float64, device="cuda")
    Wc = W.cuda()
    K.conv2d_fprop(x.cuda(), Wc, b.cuda(), out, stats, s, wpack=_pack(Wc), use_tc=True)
    torch.cuda.synchronize()
    assert rel_err(out, ref) < 1e-2
    ref_stats = torch.stack([ref.reshape(N, -1).sum(1), (ref.reshape(N, -1) ** 2).sum(1)], 1).detach()
    assert float(((stats.cpu() - ref_stats).abs() / ref_stats.abs().max()).max()) < 1e-4
    dW = torch.zeros(5, 5, A, B, dtype=torch.float32, device="cuda")
    K.conv2d_wgrad(x.cuda(), small.cuda(), dW, s, use_tc=True)
    torch.cuda.synchronize()
    assert rel_err(dW, dW_ref) < 1e-4


def _norm_bwd_reference(g, z, gamma, beta, alpha, eps=1e-3):
    """fp64: dy = g * LeakyReLU'(y) and the per-sample sums (sum dy, sum dy*xhat) of instance.py's norm."""
    N = z.shape[0]
    zf = z.double().reshape(N, -1)
    mu = zf.mean(1, keepdim=True)
    sigma = ((zf - mu) ** 2).mean(1, keepdim=True).sqrt()
    xhat = (zf - mu) / (sigma + eps)
    y = gamma * xhat + beta
    dy = g.double().reshape(N, -1) * torch.where(y > 0, 1.0, alpha)
    red = torch.stack([dy.sum(1), (dy * xhat).sum(1)], 1)
    return dy.reshape(g.shape), red


NB_CASES = [
    ("fprop", (2, 64, 64, 64, 128, 2)),      # tc_conv fprop (decoder backward), N tile 128
    ("fprop", (3, 16, 16, 256, 384, 2)),     # 2 x 192-channel tiles, a tile spans two samples
    ("fprop", (2, 128, 128, 32, 64, 2)),     # dec4 backward
    ("fprop", (2, 128, 128, 3, 32, 1)),      # final conv backward on the RGB im2col kernel
    ("dgrad", (3, 16, 16, 256, 384, 2)),     # tc_conv dgrad (encoder backward), 256 output channels
    ("dgrad", (2, 32, 32, 128, 256, 2)),
    ("dgrad", (2, 64, 64, 64, 128, 2)),      # four-phase kernel
    ("dgrad", (2, 128, 128, 32, 64, 2)),
]


@pytest.mark.parametrize("op,geom", NB_CASES)
def test_tc_fused_norm_backward_epilogue(op, geom):
    """A backward conv launch with the lg_norm_bwd_t epilogue == the plain launch followed by pass 1 of the
    norm backward (and the whole IN backward agrees with autograd of the oracle's instance_norm)."""
    from littlegan_b200 import kernels as K
    N, Hb, Wb, A, B, s = geom
    kop = K.OP_FPROP if op == "fprop" else K.OP_DGRAD
    assert K.tc_supported(kop, N, Hb, Wb, A, B, s) and K.norm_bwd_supported(kop, N, Hb, Wb, A, B, s)
    W = _rand((5, 5, A, B), 2, torch.bfloat16, 0.05).float()
    if op == "fprop":
        x = _rand((N, Hb, Wb, A), 1, torch.bfloat16)
        g_ref = O.conv2d_same(x.double(), W.double(), torch.zeros(B, dtype=torch.float64), s)
        oshape = (N, Hb // s, Wb // s, B)
    else:
        x = _rand((N, Hb // s, Wb // s, B), 1, torch.bfloat16)
        g_ref = O.conv2d_transpose_same(x.double(), W.double(), torch.zeros(A, dtype=torch.float64), s)
        oshape = (N, Hb, Wb, A)
    z = (_rand(oshape, 3, torch.float32, 1.5) + 0.3).to(torch.bfloat16)
    gamma, beta, alpha = torch.tensor([0.9]), torch.tensor([0.15]), 0.3
    dy_ref, red_ref = _norm_bwd_reference(g_ref, z, 0.9, 0.15, alpha)

    zc, Wc = z.cuda(), W.cuda()
    stats = torch.zeros(N, 2, dtype=torch.float64, device="cuda")
    K.rowstats(zc, stats, 1.0)
    red = torch.zeros(N, 2, dtype=torch.float64, device="cuda")
    nb = K.norm_bwd_desc(zc, stats, gamma.cuda(), beta.cuda(), red, 1e-3, alpha)
    out = torch.zeros(oshape, dtype=torch.bfloat16, device="cuda")
    if op == "fprop":
        K.conv2d_fprop(x.cuda(), Wc, None, out, None, s, wpack=_pack(Wc), use_tc=True, norm_bwd=nb)
    else:
        K.conv2d_dgrad(x.cuda(), Wc, None, out, None, s, K.ACT_NONE, wpack=_pack(Wc), use_tc=True, norm_bwd=nb)
    torch.cuda.synchronize()
    assert rel_err(out, dy_ref) < 1e-2
    assert float(((red.cpu() - red_ref).abs() / red_ref.abs().max()).max()) < 1e-4

    # pass 2 on the handed-over dy (+ fused bias gradient) against autograd through the oracle's norm
    zr = z.double().requires_grad_(True)
    gr = gamma.double().requires_grad_(True); br = beta.double().requires_grad_(True)
    a = torch.nn.functional.leaky_relu(O.instance_norm(zr, gr, br), alpha)
    dz_ref, dg_ref, db_ref = torch.autograd.grad((a * g_ref).sum(), [zr, gr, br])
    dz = torch.empty_like(zc)
    dgam = torch.zeros(1, device="cuda"); dbet = torch.zeros(1, device="cuda")
    C = oshape[-1]
    dbias = torch.zeros(C, device="cuda") if K.bias_grad_fusable(zc) else None
    K.instnorm_act_bwd(out, zc, stats, gamma.cuda(), beta.cuda(), red, dz, dgam, dbet, 1e-3, 1.0, alpha,
                       dy_ready=True, dbias=dbias)
    assert rel_err(dz, dz_ref) < 2e-2
    assert rel_err(dgam, dg_ref) < 2e-3 and rel_err(dbet, db_ref) < 2e-3
    if dbias is not None:      # sums of 10^3..10^4 terms built from the bf16-rounded dy, with cancellation
        assert rel_err(dbias, dz_ref.reshape(-1, C).sum(0)) < 1e-2


@pytest.mark.parametrize("op,geom", NB_CASES)
def test_tc_fused_norm_backward_epilogue_cta_pairs(op, geom, forced_cta_pairs):
    test_tc_fused_norm_backward_epilogue(op, geom)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(3, 16, 16, 64), (2, 8, 8, 32), (2, 64, 64, 256)])
def test_instnorm_bwd_fused_bias_grad(shape, dtype):
    from littlegan_b200 import kernels as K
    N, C = shape[0], shape[-1]
    z = (_rand(shape, 20, torch.float32, 2.0) + 0.5).to(dtype)
    gout = _rand(shape, 22, dtype)
    gamma = torch.tensor([1.3]); beta = torch.tensor([-0.2])
    zr = z.double().requires_grad_(True)
    ref = torch.nn.functional.leaky_relu(O.instance_norm(zr, gamma.double(), beta.double()), 0.3)
    (dz_ref,) = torch.autograd.grad((ref * gout.double()).sum(), [zr])
    zc = z.cuda()
    assert K.bias_grad_fusable(zc)
    stats = torch.zeros(N, 2, dtype=torch.float64, device="cuda")
    K.rowstats(zc, stats, 1.0)
    red = torch.zeros(N, 2, dtype=torch.float64, device="cuda")
    dz = torch.empty_like(zc)
    dbias = torch.ones(C, device="cuda")
    K.instnorm_act_bwd(gout.cuda(), zc, stats, gamma.cuda(), beta.cuda(), red, dz, None, None, 1e-3, 1.0, 0.3,
                       dbias=dbias)
    assert rel_err(dz, dz_ref) < tol(dtype) * 2
    assert rel_err(dbias - 1, dz_ref.reshape(-1, C).sum(0)) < (1e-4 if dtype == torch.float32 else 2e-3)


TC_W = [
    (4, 64, 64, 64, 128, 2),     # enc2 / dec3: 2 taps per 128-row tile
    (6, 32, 32, 128, 256, 2),    # enc3 / dec2
    (5, 16, 16, 256, 384, 2),    # enc4 / dec1: 2 a-tiles, b split 2 x 192, 8x8 position boxes (odd batch)
    (3, 128, 128, 32, 64, 2),    # dec4: row-streaming wgrad kernel (all taps resident in TMEM)
    (5, 32, 128, 32, 64, 2),     # ... short maps: several strips per CTA, ring wrap
    (2, 32, 32, 64, 64, 1),      # stride 1
]


@pytest.mark.parametrize("geom", TC_W)
def test_tc_wgrad(geom):
    from littlegan_b200 import kernels as K
    N, Hb, Wb, A, B, s = geom
    assert K.tc_supported(K.OP_WGRAD, N, Hb, Wb, A, B, s)
    big = _rand((N, Hb, Wb, A), 7, torch.bfloat16)
    small = _rand((N, Hb // s, Wb // s, B), 8, torch.bfloat16)
    W = torch.zeros(5, 5, A, B, dtype=torch.float64, requires_grad=True)
    y = O.conv2d_same(big.double(), W, torch.zeros(B, dtype=torch.float64), s)
    (ref,) = torch.autograd.grad((y * small.double()).sum(), W)
    dW = torch.zeros(5, 5, A, B, dtype=torch.float32, device="cuda")
    K.conv2d_wgrad(big.cuda(), small.cuda(), dW, s, use_tc=True)
    torch.cuda.synchronize()
    assert rel_err(dW, ref) < 1e-4          # bf16 operands are exact inputs; accumulation is fp32
    K.conv2d_wgrad(big.cuda(), small.cuda(), dW, s, use_tc=True)
    assert rel_err(dW, 2 * ref) < 1e-4


# (N, Hb, A_big, A, B, stride): the three layers served by the row-streaming kernel (csrc/tc_rowconv.cu)
ROW_CASES = [
    (3, 128, 8, 3, 32, 1),       # final conv input-gradient (M = 128 rows)
    (3, 128, 8, 3, 64, 2),       # encoder conv1 forward (64-pixel rows)
    (2, 128, 32, 32, 64, 2),     # decoder conv4 input-gradient
    (5, 32, 8, 3, 64, 2),        # short maps: several strips per CTA, ring wrap
]


@pytest.mark.parametrize("with_nb", [False, True])
@pytest.mark.parametrize("case", ROW_CASES)
def test_rowconv_fprop(case, with_nb):
    """Row-streaming shift-GEMM fprop == the oracle's SAME convolution (bias + statistics, or the fused
    norm-backward epilogue)."""
    from littlegan_b200 import kernels as K
    N, Hb, A_big, A, B, s = case
    Wb = 128
    assert K.fprop_rows_supported(N, Hb, Wb, A_big, A, B, s)
    x = _rand((N, Hb, Wb, A), 11, torch.bfloat16)
    xp = torch.zeros(N, Hb, Wb, A_big, dtype=torch.bfloat16)
    xp[..., :A] = x
    W = _rand((5, 5, A, B), 12, torch.bfloat16, 0.05).float()
    oshape = (N, Hb // s, Wb // s, B)
    out = torch.zeros(oshape, dtype=torch.bfloat16, device="cuda")
    if not with_nb:
        b = _rand((B,), 13, torch.float32)
        ref = O.conv2d_same(x.double(), W.double(), b.double(), s)
        stats = torch.zeros(N, 2, dtype=torch.float64, device="cuda")
        K.conv2d_fprop_rows(xp.cuda(), K.pack_rowconv_weights(W.cuda(), A_big, s), b.cuda(), out, stats, s, A)
        torch.cuda.synchronize()
        assert rel_err(out, ref) < 2e-2
        ref_stats = torch.stack([ref.reshape(N, -1).sum(1), (ref.reshape(N, -1) ** 2).sum(1)], 1)
        assert float(((stats.cpu() - ref_stats).abs() / ref_stats.abs().max()).max()) < 1e-4
        return
    g_ref = O.conv2d_same(x.double(), W.double(), torch.zeros(B, dtype=torch.float64), s)
    z = (_rand(oshape, 14, torch.float32, 1.5) + 0.3).to(torch.bfloat16)
    alpha = 0.3
    dy_ref, red_ref = _norm_bwd_reference(g_ref, z, 0.9, 0.15, alpha)
    zc = z.cuda()
    stats = torch.zeros(N, 2, dtype=torch.float64, device="cuda")
    K.rowstats(zc, stats, 1.0)
    red = torch.zeros(N, 2, dtype=torch.float64, device="cuda")
    nb = K.norm_bwd_desc(zc, stats, torch.tensor([0.9]).cuda(), torch.tensor([0.15]).cuda(), red, 1e-3, alpha)
    K.conv2d_fprop_rows(xp.cuda(), K.pack_rowconv_weights(W.cuda(), A_big, s), None, out, None, s, A, norm_bwd=nb)
    torch.cuda.synchronize()
    assert rel_err(out, dy_ref) < 1e-2
    assert float(((red.cpu() - red_ref).abs() / red_ref.abs().max()).max()) < 1e-4


@pytest.mark.parametrize("N,act", [(3, 1), (5, 0)])
def test_rowdeconv_rgb(N, act):
    """Final Conv2DTranspose(3, s1) on the row-streaming kernel, with the 8-channel padded second output."""
    from littlegan_b200 import kernels as K
    Hb, Wb, A, B, s = 128, 128, 3, 32, 1
    assert K.dgrad_rgb_supported(N, Hb, Wb, A, B, s)
    x = _rand((N, Hb, Wb, B), 4, torch.bfloat16)
    W = _rand((5, 5, A, B), 5, torch.bfloat16, 0.05).float()
    b = _rand((A,), 6, torch.float32)
    pre = O.conv2d_transpose_same(x.double(), W.double(), b.double(), s)
    ref = torch.tanh(pre) if act else pre
    out = torch.zeros(N, Hb, Wb, A, dtype=torch.bfloat16, device="cuda")
    out8 = torch.full((N, Hb, Wb, 8), 7.0, dtype=torch.bfloat16, device="cuda")
    stats = torch.zeros(N, 2, dtype=torch.float64, device="cuda")
    K.conv2d_dgrad_rgb(x.cuda(), W.cuda(), b.cuda(), out, out8, stats, s, act)
    torch.cuda.synchronize()
    assert rel_err(out, ref) < 1e-2
    assert torch.equal(out8[..., :3], out) and float(out8[..., 3:].abs().max()) == 0.0
    ref_stats = torch.stack([pre.reshape(N, -1).sum(1), (pre.reshape(N, -1) ** 2).sum(1)], 1)
    assert float(((stats.cpu() - ref_stats).abs() / ref_stats.abs().max()).max()) < 1e-4


def test_pad_channels_rgb8():
    from littlegan_b200 import kernels as K
    x = _rand((3, 16, 24, 3), 9, torch.bfloat16).cuda()
    xp = K.pad_channels(x, torch.full((3, 16, 24, 8), 5.0, dtype=torch.bfloat16, device="cuda"))
    assert torch.equal(xp[..., :3], x) and float(xp[..., 3:].abs().max()) == 0.0


@pytest.mark.parametrize("N,Hb", [(3, 128), (5, 32)])
def test_rowdgrad(N, Hb):
    """Decoder conv4 forward (64 -> 32 channels, stride 2) on the row-streaming dgrad kernel."""
    from littlegan_b200 import kernels as K
    Wb, A, B, s = 128, 32, 64, 2
    assert K.dgrad_rows_supported(N, Hb, Wb, A, B, s)
    x = _rand((N, Hb // s, Wb // s, B), 4, torch.bfloat16)
    W = _rand((5, 5, A, B), 5, torch.bfloat16, 0.05).float()
    b = _rand((A,), 6, torch.float32)
    ref = O.conv2d_transpose_same(x.double(), W.double(), b.double(), s)
    out = torch.zeros(N, Hb, Wb, A, dtype=torch.bfloat16, device="cuda")
    stats = torch.zeros(N, 2, dtype=torch.float64, device="cuda")
    K.conv2d_dgrad_rows(x.cuda(), K.pack_rowdgrad_weights(W.cuda()), b.cuda(), out, stats, s)
    torch.cuda.synchronize()
    assert rel_err(out, ref) < 1e-2
    ref_stats = torch.stack([ref.reshape(N, -1).sum(1), (ref.reshape(N, -1) ** 2).sum(1)], 1)
    assert float(((stats.cpu() - ref_stats).abs() / ref_stats.abs().max()).max()) < 1e-4


@pytest.mark.parametrize("N,Hb", [(3, 128), (5, 32)])
def test_rowdeconv_rgb_stride2(N, Hb):
    """Input gradient of the encoder's first Conv2D (64 -> 3 channels, stride 2) on the row-streaming kernel."""
    from littlegan_b200 import kernels as K
    Wb, A, B, s = 128, 3, 64, 2
    assert K.dgrad_rgb_supported(N, Hb, Wb, A, B, s)
    x = _rand((N, Hb // s, Wb // s, B), 4, torch.bfloat16)
    W = _rand((5, 5, A, B), 5, torch.bfloat16, 0.05).float()
    ref = O.conv2d_transpose_same(x.double(), W.double(), torch.zeros(A, dtype=torch.float64), s)
    out = torch.zeros(N, Hb, Wb, A, dtype=torch.bfloat16, device="cuda")
    stats = torch.zeros(N, 2, dtype=torch.float64, device="cuda")
    K.conv2d_dgrad_rgb(x.cuda(), W.cuda(), None, out, None, stats, s, K.ACT_NONE)
    torch.cuda.synchronize()
    assert rel_err(out, ref) < 1e-2
    ref_stats = torch.stack([ref.reshape(N, -1).sum(1), (ref.reshape(N, -1) ** 2).sum(1)], 1)
    assert float(((stats.cpu() - ref_stats).abs() / ref_stats.abs().max()).max()) < 1e-4


# ------------------------------------------------------------------------------------------------
# input augmentation of the train step (eager_trainer.py:127-131)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("N,H", [(5, 128), (3, 32)])
def test_augment_matches_oracle_with_injected_draws(N, H, dtype):
    from littlegan_b200 import kernels as K
    g = torch.Generator().manual_seed(11)
    x = torch.rand(N, H, H, 3, generator=g) * 2 - 1
    flips = (torch.rand(N, generator=g) < 0.5).float()
    flips[0], flips[1] = 1.0, 0.0
    db, fc, dh = 0.013, 0.81, -0.027
    noise = torch.randn(N, H, H, 3, generator=g) * 0.02
    ref = O.augment(x.double(), flips.double(), db, fc, dh, noise.double())
    params = torch.zeros(4 + 4 * N)
    params[0], params[1], params[2] = db, fc, dh
    params[7::4] = flips
    out = torch.empty(N, H, H, 3, dtype=dtype, device="cuda")
    K.augment(x.cuda(), out, params.cuda(), state=None, noise=noise.cuda(), draw=False)
    torch.cuda.synchronize()
    # a pixel whose hue sits on a sector boundary may land on either side of it; the map is continuous there
    assert rel_err(out, ref) < (2e-6 if dtype == torch.float32 else 5e-3)
    assert float((out.double().cpu() - ref).abs().max()) < (1e-5 if dtype == torch.float32 else 2e-2)


def test_augment_draws_and_noise_statistics():
    from littlegan_b200 import kernels as K
    N, H = 64, 128
    g = torch.Generator().manual_seed(3)
    x = (torch.rand(N, H, H, 3, generator=g) * 2 - 1).cuda()
    state = K.augment_state(1234, "cuda")
    params = torch.zeros(4 + 4 * N, device="cuda")
    out = torch.empty(N, H, H, 3, device="cuda")
    seen, flips_all, groups = [], [], []
    for step in range(6):
        K.augment(x, out, params, state=state)
        p = params.cpu()
        db, fc, dh = float(p[0]), float(p[1]), float(p[2])
        assert -0.02 <= db < 0.02 and 0.75 <= fc < 1.003 and -0.03 <= dh < 0.03
        assert int(state[1]) == step + 1
        flips = p[7::4]
        assert set(flips.tolist()) <= {0.0, 1.0}
        flips_all.append(flips)
        seen.append((db, fc, dh))
        # the noise is what remains after the deterministic chain with the drawn parameters
        det = O.augment(x.double().cpu(), flips.double(), db, fc, dh, torch.zeros(N, H, H, 3, dtype=torch.float64))
        nz = out.double().cpu() - det
        assert abs(float(nz.mean())) < 2e-4 and abs(float(nz.std()) - 0.02) < 4e-4
        assert abs(float((nz ** 4).mean() / nz.var() ** 2) - 3.0) < 0.05          # Gaussian kurtosis
        per_img = nz.reshape(N, -1)
        assert float(torch.corrcoef(per_img[:8, :4096])[0, 1:].abs().max()) < 0.08   # independent across images
        groups.append(nz.reshape(-1, 12)[:65536])               # the 12 values one thread draws for its 4 pixels
    # independent across steps, at every shift inside a thread's 12 draws (a Philox offset that advances by
    # fewer than 12 outputs per step would replay values of step k at shifted positions in step k+1 / k+2)
    for d in (1, 2, 3):
        for k in range(6 - d):
            a, b = groups[k], groups[k + d]
            a = (a - a.mean(0)) / a.std(0)
            b = (b - b.mean(0)) / b.std(0)
            cross = (a.T @ b) / a.shape[0]
            assert float(cross.abs().max()) < 0.03, (k, d, float(cross.abs().max()))
    assert len(set(seen)) == 6
    frac = float(torch.cat(flips_all).mean())
    assert 0.38 < frac < 0.62
    # the streams are functions of (seed, step): the same state reproduces the same batch
    state2 = K.augment_state(1234, "cuda")
    out2 = torch.empty_like(out)
    K.augment(x, out2, torch.zeros_like(params), state=state2)
    state3 = K.augment_state(1234, "cuda")
    out3 = torch.empty_like(out)
    K.augment(x, out3, torch.zeros_like(params), state=state3)
    assert torch.equal(out2, out3) and not torch.equal(out2, out)


@pytest.mark.parametrize("n", [1, 15, 16, 4099, 2 * 128 * 128 * 3])
def test_u8_rescale_is_bit_exact(n):
    """data_rescale (utils.py:51-52): x / 127.5 - 1 in fp32, two roundings - bit-exact against NumPy; the bf16
    form is that value rounded once more.  Sizes cover the scalar tail and the 16-byte vector body."""
    from littlegan_b200 import kernels as K
    g = torch.Generator().manual_seed(n)
    x = torch.randint(0, 256, (n,), generator=g, dtype=torch.uint8)
    if n >= 256:
        x[:256] = torch.arange(256, dtype=torch.uint8)           # every byte value
    want = (x.numpy().astype(np.float32) / np.float32(127.5)) - np.float32(1)
    out = torch.empty(n, dtype=torch.float32, device="cuda")
    K.u8_rescale(x.cuda(), out)
    assert np.array_equal(out.cpu().numpy(), want)
    outb = torch.empty(n, dtype=torch.bfloat16, device="cuda")
    K.u8_rescale(x.cuda(), outb)
    assert torch.equal(outb.cpu(), torch.from_numpy(want).to(torch.bfloat16))
    with pytest.raises(Exception):
        K.u8_rescale(x.cuda().float(), out)


# ------------------------------------------------------------------------------------------------
# Frechet distance without a library eigen-solver (fid.py:112-163): fp64 GEMM + Newton-Schulz
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 37, 128, 200, 512])
def test_dgemm_matches_fp64_matmul(n):
    """lg_dgemm: C = alpha A B + diag I on row-major fp64 matrices, any n (tile tails)."""
    from littlegan_b200 import kernels as K
    g = torch.Generator().manual_seed(n)
    A = torch.randn(n, n, generator=g, dtype=torch.float64)
    B = torch.randn(n, n, generator=g, dtype=torch.float64)
    C = torch.empty(n, n, dtype=torch.float64, device="cuda")
    K.dgemm(A.cuda(), B.cuda(), C, -0.5, 1.5)
    ref = -0.5 * (A @ B) + 1.5 * torch.eye(n, dtype=torch.float64)
    assert float((C.cpu() - ref).abs().max()) < 1e-12 * max(1.0, float(ref.abs().max()))
    st = torch.empty(3, dtype=torch.float64, device="cuda")
    K.dmat_stats(C, st)
    assert abs(float(st[0]) - float(ref.trace())) < 1e-10 * max(1.0, abs(float(ref.trace())))
    assert abs(float(st[1]) - float((ref * ref).sum())) < 1e-10 * float((ref * ref).sum())
    assert abs(float(st[2]) - float((ref - ref.T).abs().max())) < 1e-12 * max(1.0, float(ref.abs().max()))
    with pytest.raises(Exception):
        K.dgemm(C, C, C)                                  # the output must not alias an operand


@pytest.mark.parametrize("d,n1,n2", [(256, 3000, 2000), (24, 400, 300), (200, 1000, 150), (384, 60, 90)])
def test_frechet_distance_matches_scipy(d, n1, n2):
    """Full-rank, odd-sized and SINGULAR (fewer samples than dimensions) covariances against the reference's own
    arithmetic (scipy.linalg.sqrtm of sigma1 sigma2).  scipy's Schur square root of a singular product is itself
    only good to ~1e-7 relative, hence the looser bound there."""
    from littlegan_b200 import fid
    rng = np.random.RandomState(d)
    x1 = (rng.randn(n1, d) @ (rng.randn(d, d) / np.sqrt(d))) * 0.5 + 0.3
    x2 = (rng.randn(n2, d) @ (rng.randn(d, d) / np.sqrt(d))) * 0.7 + 0.1
    m1, s1 = fid_oracle.activation_statistics(x1)
    m2, s2 = fid_oracle.activation_statistics(x2)
    ref = fid_oracle.frechet_distance(m1, s1, m2, s2)
    got = fid.calculate_frechet_distance(m1, s1, m2, s2)
    singular = min(n1, n2) <= d
    assert abs(got - ref) < (1e-6 if singular else 1e-9) * abs(ref), (got, ref)
    assert abs(fid.calculate_frechet_distance(m1, s1, m1, s1)) < 1e-7 * np.trace(s1)
    # the square root itself
    R, tr, iters = fid._sqrtm_psd(torch.from_numpy(s1).cuda())
    assert iters < fid._NS_MAX_ITERS
    assert float((R @ R - torch.from_numpy(s1).cuda()).abs().max()) < 1e-9 * np.abs(s1).max()


def test_scale_shift_and_normal_fill():
    """The in-step copies / (cond + 1) / 2 (eager_trainer.py:155) and the generator noise (eager_trainer.py:125) on
    this library's own kernels."""
    from littlegan_b200 import kernels as K
    x = _rand((7, 41), 90, torch.float32)
    for ddt in DTYPES:
        out = torch.empty(7, 41, dtype=ddt, device="cuda")
        K.scale_shift(x.cuda(), out, 0.5, 0.5)
        ref = (x * 0.5 + 0.5).to(ddt)
        assert torch.equal(out.cpu(), ref)
    out = torch.empty(7, 41, device="cuda")
    K.scale_shift(x.cuda(), out)
    assert torch.equal(out.cpu(), x)
    # N(0,1): moments, reproducible from (seed, step), different at every step, any length (vector tail)
    for n in (64 * 93, 1000003, 5):
        st = K.normal_state(1234, "cuda")
        a = K.normal_fill(torch.empty(n, device="cuda"), st).clone()
        assert int(st[1]) == 1 and int(st[2]) == 0
        b = K.normal_fill(torch.empty(n, device="cuda"), st).clone()
        assert int(st[1]) == 2
        st2 = K.normal_state(1234, "cuda")
        a2 = K.normal_fill(torch.empty(n, device="cuda"), st2)
        assert torch.equal(a, a2) and not torch.equal(a, b)
        if n > 1000:
            for t in (a, b):
                assert abs(float(t.mean())) < 4.0 / n ** 0.5 and abs(float(t.std()) - 1.0) < 4.0 / n ** 0.5
                assert abs(float((t ** 4).mean()) - 3.0) < 0.2
            assert abs(float((a * b).mean())) < 4.0 / n ** 0.5
            assert abs(float((a[:-1] * a[1:]).mean())) < 4.0 / n ** 0.5
