"""Inception-2015 pool_3 forward (fid.py:36-106) against the CPU oracle, through the C-ABI.  -m gpu."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import inception_oracle as IO
from tests.util import rel_err, tol

pytestmark = pytest.mark.gpu


def _rand(shape, seed, dtype=torch.float32, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g, dtype=torch.float64) * scale).to(dtype)


# (N, H, W, Cin, Cout, (kh,kw), stride, (ph,pw))
UNIT_GEOMS = [
    (2, 31, 31, 3, 32, (3, 3), 2, (0, 0)),      # Conv2d_1a: RGB in, stride 2, VALID, odd size (SIMT in every mode)
    (2, 17, 17, 24, 48, (1, 7), 1, (0, 3)),     # 1x7
    (2, 17, 17, 24, 48, (7, 1), 1, (3, 0)),     # 7x1
    (3, 9, 9, 48, 64, (5, 5), 1, (2, 2)),       # 5x5 SAME
    (1, 8, 8, 80, 80, (1, 1), 1, (0, 0)),       # 1x1, K = 80 (K tail inside a 64-wide stage)
    (2, 9, 11, 16, 32, (3, 3), 2, (0, 0)),      # non-square map, stride 2
    (2, 8, 8, 448, 384, (3, 3), 1, (1, 1)),     # Mixed_7 dbl_2: K = 4032 (63 stages), Cout split into 2 column tiles
    (1, 35, 35, 64, 96, (3, 3), 1, (1, 1)),     # Mixed_5: 1225 positions = 10 tiles, M tail of 73 rows
    (3, 8, 8, 128, 320, (1, 1), 1, (0, 0)),     # Cout 320 = 2 x 160 columns
    (2, 17, 17, 40, 24, (3, 3), 1, (1, 1)),     # Cout % 16 != 0: no tensor-core form, SIMT also in bf16
]


@pytest.mark.parametrize("mode", ["fp32", "bf16", "bf16-tc"])
@pytest.mark.parametrize("geom", UNIT_GEOMS)
def test_conv_bn_relu_unit(geom, mode):
    """One unit reading a channel slice and writing a channel slice of a wider concat buffer; bf16 maps on the SIMT
    path (from W) and on the tcgen05 path (from the packed operand)."""
    from littlegan_b200 import kernels as K
    dtype = torch.float32 if mode == "fp32" else torch.bfloat16
    N, H, W_, Cin, Cout, (kh, kw), s, (ph, pw) = geom
    x_off, y_off, Cx, Cy = 8, 16, Cin + 16, Cout + 24
    x = _rand((N, H, W_, Cx), 1, dtype)
    Wt = _rand((kh, kw, Cin, Cout), 2, torch.float32, 0.2)
    scale = _rand((Cout,), 3).abs() + 0.5
    shift = _rand((Cout,), 4)
    wpack = None
    if mode == "bf16-tc":
        wpack = K.pack_conv_bn_weights(Wt.cuda())
        assert (wpack is not None) == (Cin % 8 == 0 and Cout % 16 == 0)
        Wt = Wt.to(torch.bfloat16).float()                # the tensor-core operand is W rounded to bf16
    xs = x[..., x_off:x_off + Cin].double().permute(0, 3, 1, 2)
    ref = F.conv2d(xs, Wt.double().permute(3, 2, 0, 1), None, stride=s, padding=(ph, pw))
    ref = F.relu(ref * scale.double()[None, :, None, None] + shift.double()[None, :, None, None]).permute(0, 2, 3, 1)
    y = torch.full((N, ref.shape[1], ref.shape[2], Cy), 7.0, dtype=dtype, device="cuda")
    K.conv2d_bn_relu(x.cuda(), Wt.cuda(), scale.cuda(), shift.cuda(), y, y_off, stride=s, pad=(ph, pw), x_off=x_off,
                     cin=Cin, wpack=wpack)
    torch.cuda.synchronize()
    assert rel_err(y[..., y_off:y_off + Cout], ref) < tol(dtype)
    keep = torch.cat([y[..., :y_off], y[..., y_off + Cout:]], -1)
    assert bool((keep == 7.0).all())                       # the rest of the concat buffer is untouched


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("cfg", [(3, 2, 0, 0), (3, 1, 1, 0), (3, 1, 1, 1), (3, 1, 1, 2)])
def test_pool2d_modes(cfg, dtype):
    from littlegan_b200 import kernels as K
    k, s, p, mode = cfg
    x = _rand((2, 17, 15, 24), 5, dtype)
    xc = x.double().permute(0, 3, 1, 2)
    if mode == 0:
        ref = F.max_pool2d(xc, k, stride=s, padding=p)
    else:
        ref = F.avg_pool2d(xc, k, stride=s, padding=p, count_include_pad=(mode == 2))
    ref = ref.permute(0, 2, 3, 1)
    y = torch.zeros(2, ref.shape[1], ref.shape[2], 40, dtype=dtype, device="cuda")
    K.pool2d(x.cuda(), y, 8, k, s, p, mode)
    assert rel_err(y[..., 8:32], ref) < (1e-6 if dtype == torch.float32 else 1e-2)
    assert float(y[..., :8].abs().max()) == 0 and float(y[..., 32:].abs().max()) == 0


@pytest.mark.parametrize("cfg", [(64, 35, 35, 288, 1, 1, 1), (64, 35, 35, 288, 1, 1, 2), (32, 71, 71, 192, 2, 0, 0),
                                 (100, 8, 8, 2048, 1, 1, 0), (16, 17, 17, 768, 2, 0, 0)])
def test_pool3_sliding_window_full_size(cfg):
    """The Inception pools at their real sizes (bf16): the kernel walks multi-row segments with a 3-row window."""
    from littlegan_b200 import kernels as K
    N, H, W_, C, s, p, mode = cfg
    x = _rand((N, H, W_, C), 9, torch.bfloat16).cuda()
    xc = x.float().permute(0, 3, 1, 2)
    if mode == 0:
        ref = F.max_pool2d(xc, 3, stride=s, padding=p)
    else:
        ref = F.avg_pool2d(xc, 3, stride=s, padding=p, count_include_pad=(mode == 2))
    ref = ref.permute(0, 2, 3, 1)
    y = torch.zeros(N, ref.shape[1], ref.shape[2], C + 16, dtype=torch.bfloat16, device="cuda")
    K.pool2d(x, y, 8, 3, s, p, mode)
    assert rel_err(y[..., 8:8 + C], ref) < 1e-2
    if mode == 0:
        assert torch.equal(y[..., 8:8 + C].float(), ref)        # a max of bf16 values is exact
    assert float(y[..., :8].abs().max()) == 0 and float(y[..., 8 + C:].abs().max()) == 0


@pytest.mark.parametrize("src", ["u8", "f32"])
def test_resize_bilinear_tf1_and_normalise(src):
    from littlegan_b200 import kernels as K
    g = torch.Generator().manual_seed(6)
    img = torch.randint(0, 256, (2, 128, 128, 3), generator=g, dtype=torch.uint8)
    ref = (IO.resize_bilinear_tf1(img.double(), 299, 299) - 128.0) / 128.0
    x = img.cuda() if src == "u8" else img.float().cuda()
    y = torch.empty(2, 299, 299, 3, dtype=torch.float32, device="cuda")
    K.resize_bilinear_norm(x, y, 128.0, 1.0 / 128.0)
    # the kernel computes source coordinates in fp32 as TF's op does (dst * float(in/out)); the oracle in fp64
    assert float((y.double().cpu() - ref).abs().max()) < 1e-4
    y8 = torch.full((2, 299, 299, 8), 5.0, dtype=torch.bfloat16, device="cuda")       # channel-padded bf16 form
    K.resize_bilinear_norm(x, y8, 128.0, 1.0 / 128.0)
    assert float(y8[..., 3:].abs().max()) == 0 and float((y8[..., :3].double().cpu() - ref).abs().max()) < 1e-2
    # an up-scale whose source coordinates hit the last row / column clamp
    small = torch.randint(0, 256, (1, 5, 7, 3), generator=g, dtype=torch.uint8)
    ref2 = IO.resize_bilinear_tf1(small.double(), 11, 13)
    y2 = torch.empty(1, 11, 13, 3, dtype=torch.float32, device="cuda")
    K.resize_bilinear_norm(small.cuda(), y2, 0.0, 1.0)
    assert float((y2.double().cpu() - ref2).abs().max()) < 1e-4


def test_global_avgpool():
    from littlegan_b200 import kernels as K
    x = _rand((3, 8, 8, 2048), 7)
    y = torch.empty(3, 2048, dtype=torch.float32, device="cuda")
    K.global_avgpool(x.cuda(), y)
    assert rel_err(y, x.double().mean(dim=(1, 2))) < 1e-6


@pytest.mark.parametrize("variant", ["fid", "torchvision"])
def test_pool3_features_match_oracle_fp32(variant):
    """Every unit's activation and pool_3 within 1e-4 (max-norm relative) of the fp64 oracle on identical weights."""
    from littlegan_b200.inception import InceptionPool3
    W = IO.random_weights(seed=1)
    g = torch.Generator().manual_seed(3)
    img = torch.randint(0, 256, (2, 128, 128, 3), generator=g, dtype=torch.uint8)
    ora = IO.InceptionOracle(W, variant=variant, dtype=torch.float64)
    want = ora(img.double())
    net = InceptionPool3(weights=W, dtype="fp32", fid_variant=(variant == "fid"))
    net.taps = {}
    got = net(img.numpy())
    assert got.shape == (2, 2048) and got.dtype == torch.float32 and got.is_cuda
    worst = 0.0
    assert set(ora.taps) == set(net.taps)
    for name, ref in ora.taps.items():
        e = rel_err(net.taps[name], ref)
        worst = max(worst, e)
        assert e < 1e-4, (name, e)
    assert rel_err(got, want) < 1e-4
    print("worst unit error %.2e, pool_3 %.2e" % (worst, rel_err(got, want)))


def test_pool3_features_bf16_storage():
    from littlegan_b200.inception import InceptionPool3
    W = IO.random_weights(seed=2)
    g = torch.Generator().manual_seed(4)
    img = torch.randint(0, 256, (2, 128, 128, 3), generator=g, dtype=torch.uint8)
    want = IO.InceptionOracle(W, dtype=torch.float64)(img.double())
    net = InceptionPool3(weights=W, dtype="bf16")
    assert sum(p[3] is not None for p in net.params.values()) == 94          # every unit on tcgen05 (RGB stem padded to 8)
    got = net(img)
    # north_star's bf16 tolerance, max-norm relative (measured 4e-3: pool_3 averages 64 positions)
    assert rel_err(got, want) < 2e-2
    d = (got.double().cpu() - want).norm() / want.norm()
    assert float(d) < 1e-2


def test_fid_statistics_through_the_inception_session():
    """fid.calculate_activation_statistics(images, sess=InceptionPool3) == np.mean / np.cov of the oracle features
    (fid.py:169-188), including the dropped N % batch tail (fid.py:89-94)."""
    from littlegan_b200 import fid
    from littlegan_b200.inception import InceptionPool3
    from oracle import fid_oracle
    W = IO.random_weights(seed=5)
    g = torch.Generator().manual_seed(8)
    imgs = torch.randint(0, 256, (7, 64, 64, 3), generator=g, dtype=torch.uint8).numpy()
    net = InceptionPool3(weights=W)
    mu, sigma = fid.calculate_activation_statistics(imgs, net, batch_size=3)
    feats = IO.InceptionOracle(W, dtype=torch.float64)(torch.from_numpy(imgs[:6]).double()).numpy()
    mu_ref, sigma_ref = fid_oracle.activation_statistics(feats)
    assert np.abs(mu - mu_ref).max() < 1e-4 * np.abs(mu_ref).max()
    assert np.abs(sigma - sigma_ref).max() < 1e-3 * np.abs(sigma_ref).max()


def test_fid_given_paths_and_file_statistics(tmp_path):
    """fid.py:207-318: statistics from image files (batched, decoded ahead on a host thread) equal the in-memory
    path and the oracle; calculate_fid_given_paths over directory / .npz inputs."""
    from PIL import Image
    from littlegan_b200 import fid
    from littlegan_b200.inception import InceptionPool3
    from oracle import fid_oracle
    W = IO.random_weights(seed=6)
    rng = np.random.default_rng(5)
    dirs = []
    for d, shift in (("a", 0), ("b", 40)):
        (tmp_path / d).mkdir()
        arr = np.clip(rng.integers(0, 216, (9, 32, 32, 3)) + shift, 0, 255).astype(np.uint8)
        for i, im in enumerate(arr):
            Image.fromarray(im, "RGB").save(str(tmp_path / d / ("%03d.png" % i)))
        dirs.append((str(tmp_path / d), arr))
    net = InceptionPool3(weights=W)
    files = sorted((tmp_path / "a").glob("*.png"))
    mu, sigma = fid.calculate_activation_statistics_from_files(files, net, batch_size=4)       # 8 of 9 used
    feats = IO.InceptionOracle(W, dtype=torch.float64)(torch.from_numpy(dirs[0][1][:8]).double()).numpy()
    mu_ref, sigma_ref = fid_oracle.activation_statistics(feats)
    assert np.abs(mu - mu_ref).max() < 1e-4 * np.abs(mu_ref).max()
    assert np.abs(sigma - sigma_ref).max() < 1e-3 * np.abs(sigma_ref).max()
    mu_mem, sigma_mem = fid.calculate_activation_statistics(dirs[0][1][:8], net, batch_size=4)
    assert np.abs(mu - mu_mem).max() < 1e-6 * np.abs(mu_mem).max()
    d_ab = fid.calculate_fid_given_paths([dirs[0][0], dirs[1][0]], None, sess=net)
    d_ba = fid.calculate_fid_given_paths([dirs[1][0], dirs[0][0]], None, sess=net, low_profile=True)
    assert d_ab > 0 and abs(d_ab - d_ba) < 1e-3 * d_ab            # low_profile batches of 9 < 50 -> all 9 used in both
    mu_b, sigma_b = fid._handle_path(dirs[1][0], net)
    np.savez(tmp_path / "b.npz", mu=mu_b, sigma=sigma_b)
    d_npz = fid.calculate_fid_given_paths([dirs[0][0], str(tmp_path / "b.npz")], None, sess=net)
    assert abs(d_npz - d_ab) < 1e-6 * d_ab
    assert abs(fid.calculate_fid_given_paths([dirs[0][0], dirs[0][0]], None, sess=net)) < 1e-6 * d_ab
