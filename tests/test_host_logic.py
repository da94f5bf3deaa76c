"""Host-side logic of the product that needs no GPU: config layering, builders' surface and
`.weights` ordering, parameter arenas, partition ranges, error behaviour."""
import json
import os

import pytest
import torch

from littlegan_b200.config import Arg
from tests.util import build_product, product_args, small_args


def test_arg_layering(tmp_path):
    (tmp_path / "sample.config.json").write_text(open(
        os.path.join(os.path.dirname(__file__), "..", "littlegan_b200", "sample.config.json")).read())
    (tmp_path / "lab.config.json").write_text(json.dumps({"batch_size": 8, "attr": None, "dtype": "fp32"}))
    a = Arg(["train", "exp1", "-e", "lab", "-g", "0,1", "--debug"], config_dir=str(tmp_path))
    assert a.batch_size == 8 and a.cond_dim == 40 and a.gpu == [0, 1] and a.debug
    assert a.prefetch == a.prefetch_batch * 8 and a.result_dir.endswith("exp1") and a.dtype == "fp32"
    b = Arg.from_dict()
    assert b.cond_dim == 7 and b.noise_dim == 93 and b.conv_filter == [384, 256, 128, 64, 32]
    assert b.dtype == "bf16" and b.cuda_graph is True
    assert b.augment is True                     # the reference always augments new_image (eager_trainer.py:127-131)
    assert product_args(small_args()).augment is False and product_args(small_args(), augment=True).augment is True


def test_loss_value_host_protocol():
    """LossValue without a device: the ring-slot generation decides between the pinned copy and the fallback."""
    from littlegan_b200.eager_trainer import LossValue

    class _Ev:
        def __init__(self):
            self.waits = 0

        def synchronize(self):
            self.waits += 1

    host, ev, gen = torch.tensor([1.5, -2.25, 0.125]), _Ev(), [7]
    dev = torch.tensor([1.5, -2.25, 0.125])
    v = LossValue(dev[1], (host, ev, gen), 7, 1)
    assert float(v) == -2.25 and ev.waits == 1 and v.item() == -2.25 and "%.2f" % v.numpy() == "-2.25"
    assert format(v, ".3f") == "-2.250" and float(v.cpu()) == -2.25
    host[1] = 99.0
    gen[0] = 8                                     # the slot now belongs to a later step
    waits = ev.waits
    assert float(v) == -2.25 and ev.waits == waits   # read from the step's own device scalar, no event wait


def test_builders_surface_and_weight_order():
    pargs = product_args(small_args())
    gen, disc, adj = build_product(pargs)
    assert len(disc.weights) == 20 and len(gen.weights) == 22 and len(adj.weights) == 38
    enc = disc.encoder
    assert disc.weights[0] is enc.conv1.kernel and disc.weights[3] is enc.norm1.beta
    assert disc.weights[16] is disc.dense_pr.kernel and disc.weights[19] is disc.dense_cond.bias
    assert gen.weights[0] is gen.dense.kernel and gen.weights[2] is gen.norm.gamma
    assert gen.weights[4] is gen.decoder.conv1.kernel and gen.weights[20] is gen.conv.kernel
    assert adj.encoder is disc.encoder and adj.decoder is gen.decoder and adj.conv is gen.conv
    assert adj.weights[16] is adj.dense.kernel and adj.weights[19] is adj.norm.beta
    cf, k = pargs.conv_filter, pargs.kernel_size
    assert tuple(enc.conv1.kernel.shape) == (k, k, 3, cf[4])            # HWIO
    assert tuple(gen.decoder.conv1.kernel.shape) == (k, k, cf[1], cf[0])  # [k,k,out,in]
    assert tuple(gen.conv.kernel.shape) == (k, k, 3, cf[4])
    assert tuple(gen.dense.kernel.shape) == (pargs.noise_dim + pargs.cond_dim, pargs.init_dim ** 2 * cf[0])


def test_trainer_arenas_and_partition_ranges():
    from littlegan_b200.eager_trainer import EagerTrainer
    pargs = product_args(small_args(use_partition=True))
    gen, disc, adj = build_product(pargs)
    before = [w.clone() for w in gen.weights]
    tr = EagerTrainer(pargs, gen, disc, adj, None)
    for w, b in zip(gen.weights, before):                     # re-homing keeps values and identity
        assert torch.equal(w, b) and w.untyped_storage().data_ptr() == tr.P.untyped_storage().data_ptr()
        assert w.lg_grad.shape == w.shape
    assert tr.part_weights["Generator"][1][0] is gen.weights[4]
    assert tr.all_weights["Adjuster"][0] is adj.weights[16]
    offs = tr._offsets["Generator"]
    assert tr._range("Generator", 3) == (offs[0], offs[-1])
    assert tr._range("Generator", 5) == (offs[4], offs[8])    # (5//5)%3 = 1 -> tensors 4..7
    assert tr._range("Generator", 15) == (offs[0], offs[4])
    d = tr._offsets["Discriminator"]
    assert tr._range("Discriminator", 10) == (d[16], d[20])
    a_ = tr._offsets["Adjuster"]
    assert tr._range("Adjuster", 15) == (a_[0], a_[4])
    assert tr._get_train_weight(gen, 5)[0] is gen.weights[4] and len(tr._get_train_weight(disc, 3)) == 20
    assert tr._variant(5) == (False, 1) and tr._variant(11) == (True, None) and tr._variant(15) == (True, 0)
    gen.weights[0].zero_()                                     # weights are views of the arena
    assert float(tr.P[offs[0]:offs[0] + 8].abs().sum()) == 0.0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU error path")
def test_no_cpu_fallback():
    from littlegan_b200._lib import LittleGANError
    from littlegan_b200 import kernels as K
    pargs = product_args(small_args())
    gen, disc, adj = build_product(pargs)
    with pytest.raises(LittleGANError):
        gen([torch.zeros(2, pargs.noise_dim), torch.zeros(2, pargs.cond_dim)])
    with pytest.raises(LittleGANError):
        K.rowstats(torch.zeros(2, 8), torch.zeros(2, 2, dtype=torch.float64))


def test_instance_norm_only_reference_mode():
    from littlegan_b200.instance import InstanceNormalization
    with pytest.raises(NotImplementedError):
        InstanceNormalization(axis=3)
    n = InstanceNormalization()
    assert n.epsilon == 1e-3 and tuple(n.gamma.shape) == (1,) and float(n.gamma) == 1.0 and float(n.beta) == 0.0


def _write_celeba(tmp_path, n, dim, attrs=6, ext="png", seed=0):
    """n lossless images + an attribute list in the reference's `<name> a0 a1 ...` form."""
    import numpy as np
    from PIL import Image
    rng = np.random.default_rng(seed)
    imgs = rng.integers(0, 256, (n, dim, dim, 3), dtype=np.uint8)
    labs = rng.integers(0, 2, (n, attrs)) * 2 - 1
    d = tmp_path / "img"
    d.mkdir()
    lines = []
    for i in range(n):
        name = "%06d.%s" % (i + 1, ext)
        Image.fromarray(imgs[i], "RGB").save(str(d / name))
        lines.append(name + " " + " ".join(str(int(v)) for v in labs[i]))
    (tmp_path / "list.txt").write_text("\n".join(lines) + "\n")
    return imgs, labs


def test_celeba_pipeline_batches_shuffle_and_labels(tmp_path):
    """dataset.py:8-48: file/attribute pairing, attr filter, soft labels, batch-then-shuffle, short last batch,
    one pass per iterator; bytes out by default, the reference's fp32 tensors with as_float."""
    import numpy as np
    from littlegan_b200.dataset import ALL_LABEL, CelebA
    from littlegan_b200.eager_trainer import OutOfRangeError
    from littlegan_b200.utils import soft
    imgs, labs = _write_celeba(tmp_path, 22, 32)
    a = Arg.from_dict(batch_size=4, init_dim=2, image_dim=32, attr=[1, 4, 5], image_path=str(tmp_path / "img"),
                      attr_path=str(tmp_path / "list.txt"), image_ext="png", threads=3, prefetch_batch=1)
    ds = CelebA(a, seed=3, pin=False)
    assert ds.batches == 5 and ds.label == [ALL_LABEL[1], ALL_LABEL[4], ALL_LABEL[5]]
    seen = []
    for epoch in range(2):
        it = ds.get_new_iterator()
        order = []
        while True:
            try:
                image, cond = it.get_next()
            except OutOfRangeError:
                break
            assert image.dtype == torch.uint8 and image.shape[1:] == (32, 32, 3) and cond.dtype == torch.float32
            lo = int(np.flatnonzero((imgs.reshape(22, -1) == image[0].numpy().reshape(-1)).all(1))[0])
            assert lo % 4 == 0                                   # batches are formed before the shuffle
            n = image.shape[0]
            assert n == (2 if lo == 20 else 4)                   # tf.data keeps the short last batch
            assert np.array_equal(image.numpy(), imgs[lo:lo + n])
            want = soft(torch.tensor(labs[lo:lo + n][:, [1, 4, 5]], dtype=torch.float32))
            assert torch.equal(cond, want)
            order.append(lo)
        assert sorted(order) == [0, 4, 8, 12, 16, 20]
        seen.append(order)
    assert seen[0] != seen[1]                                    # a new shuffle per epoch
    f = CelebA(a, seed=3, pin=False, as_float=True).get_new_iterator().get_next()[0]
    assert f.dtype == torch.float32 and float(f.min()) >= -1 and float(f.max()) <= 1
    o1 = CelebA(a, seed=9, pin=False).get_new_iterator()
    o2 = CelebA(a, seed=9, pin=False).get_new_iterator()
    for _ in range(6):
        assert torch.equal(o1.get_next()[0], o2.get_next()[0])


def test_celeba_rejects_wrong_size_and_short_attr_list(tmp_path):
    from littlegan_b200.dataset import CelebA
    _write_celeba(tmp_path, 4, 16, attrs=40)
    a = Arg.from_dict(batch_size=2, init_dim=2, image_dim=32, attr=None, image_path=str(tmp_path / "img"),
                      attr_path=str(tmp_path / "list.txt"), image_ext="png", threads=1)
    with pytest.raises(ValueError):
        CelebA(a, pin=False).get_new_iterator().get_next()       # 16x16 files, image_dim 32 (tf set_shape)
    (tmp_path / "list.txt").write_text("000001.png" + " 1" * 40 + "\n")
    with pytest.raises(ValueError):
        CelebA(a, pin=False)


def test_fid_file_helpers_host_side(tmp_path):
    """fid.py:197-204, 273-318 host glue: batch loading, path validation, the weight-file check (no download here)."""
    import numpy as np
    from littlegan_b200 import fid
    imgs, _ = _write_celeba(tmp_path, 5, 16)
    files = sorted((tmp_path / "img").glob("*.png"))
    batch = fid.load_image_batch(files)
    assert batch.dtype == np.uint8 and np.array_equal(batch, imgs)
    with pytest.raises(RuntimeError, match="Invalid path"):
        fid.calculate_fid_given_paths([str(tmp_path / "img"), str(tmp_path / "nope")], None, sess=lambda x: x)
    with pytest.raises(RuntimeError, match="no Inception model file"):
        fid.check_or_download_inception(str(tmp_path))
    assert fid.check_or_download_inception(None) is None
    np.savez(tmp_path / "stats.npz", mu=np.arange(4.0), sigma=np.eye(4))
    mu, sigma = fid._handle_path(str(tmp_path / "stats.npz"), None)
    assert np.array_equal(mu, np.arange(4.0)) and np.array_equal(sigma, np.eye(4))


def test_inception_program_matches_the_oracle_table():
    """The product's graph program and the oracle's unit table are written independently; they must list the same 94
    conv units (name, channels, kernel, stride, padding) in the same order, and the blocks must concatenate to the
    widths of the published graph."""
    from littlegan_b200 import inception as I
    from oracle import inception_oracle as IO
    got = [(u["name"], u["cin"], u["cout"], tuple(u["k"]), u["s"], tuple(u["p"])) for u in I.unit_specs()]
    assert got == [tuple(u) for u in IO.UNITS]
    assert len(got) == 94 and sum(1 for _ in I._units(I.program(False))) == 94
    width, widths = 3, []
    for st in I.program():
        if isinstance(st, dict):
            assert st["cin"] == width
            width = st["cout"]
        elif isinstance(st, list):
            total = 0
            for br in st:
                c = width
                for piece in br:
                    if isinstance(piece, dict):
                        assert piece["cin"] == c, piece["name"]
                    elif isinstance(piece, list):
                        assert all(u["cin"] == c for u in piece)
                    c = I.InceptionPool3._width(piece, c)
                total += c
            width = total
            widths.append(width)
    assert widths == [256, 288, 288, 768, 768, 768, 768, 768, 1280, 2048, 2048]


def _pb_varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def _pb_len(field, payload):
    return _pb_varint((field << 3) | 2) + _pb_varint(len(payload)) + payload


def _pb_const_node(name, arr, how="content", dtype=1):
    """A serialized NodeDef{name, op='Const', attr{'value': AttrValue{tensor}}} as TensorFlow writes it."""
    import numpy as np
    shape = b"".join(_pb_len(2, _pb_varint((1 << 3) | 0) + _pb_varint(d)) for d in arr.shape)
    tensor = _pb_varint((1 << 3) | 0) + _pb_varint(dtype) + _pb_len(2, shape)
    flat = np.ascontiguousarray(arr, dtype="<f4").reshape(-1)
    if how == "content":
        tensor += _pb_len(4, flat.tobytes())
    elif how == "packed":
        tensor += _pb_len(5, flat.tobytes())
    elif how == "fill":                                        # one float_val standing for a constant-filled tensor
        tensor += _pb_varint((5 << 3) | 5) + flat[:1].tobytes()
    attr = _pb_len(5, _pb_len(1, b"value") + _pb_len(2, _pb_len(8, tensor)))
    dt_attr = _pb_len(5, _pb_len(1, b"dtype") + _pb_len(2, _pb_varint((6 << 3) | 0) + _pb_varint(dtype)))
    return _pb_len(1, _pb_len(1, name.encode()) + _pb_len(2, b"Const") + dt_attr + attr)


def test_inception_weights_from_the_reference_graphdef_format():
    """fid.py:36-42 loads classify_image_graph_def.pb through TensorFlow; here the same file is read by a minimal
    protobuf wire parser.  A synthetic GraphDef with the 2015 graph's scopes and real kernel shapes round-trips."""
    import numpy as np
    from littlegan_b200 import inception as I
    nodes, want = [], {}
    for i, u in enumerate(I.unit_specs()):
        sc = I.graphdef_scope(u["name"])
        kh, kw = u["k"]
        W = np.full((kh, kw, u["cin"], u["cout"]), 0.001 * (i + 1), np.float32)
        W[0, 0, 0, :] = np.arange(u["cout"], dtype=np.float32)
        beta = np.linspace(-1, 1, u["cout"]).astype(np.float32)
        mean = np.full(u["cout"], 0.25 * (i % 5), np.float32)
        var = np.linspace(0.5, 1.5, u["cout"]).astype(np.float32)
        want[u["name"]] = (W, beta, mean, var)
        nodes.append(_pb_const_node(sc + "/conv2d_params", W))
        nodes.append(_pb_const_node(sc + "/batchnorm/beta", beta, how="packed"))
        nodes.append(_pb_const_node(sc + "/batchnorm/moving_mean", mean, how="fill"))
        nodes.append(_pb_const_node(sc + "/batchnorm/moving_variance", var))
        nodes.append(_pb_const_node(sc + "/batchnorm/gamma", np.ones(u["cout"], np.float32)))     # present, unused
        nodes.append(_pb_len(1, _pb_len(1, (sc + "/Conv2D").encode()) + _pb_len(2, b"Conv2D")
                             + _pb_len(3, (sc + "/conv2d_params").encode())))                     # a non-Const node
    nodes.append(_pb_const_node("pool_3/_reshape/shape", np.array([1, 2048], np.float32), dtype=3))   # int32: skipped
    graph = b"".join(nodes) + _pb_len(4, _pb_varint((1 << 3) | 0) + _pb_varint(21))                # versions field
    weights = I.weights_from_graphdef(graph)
    assert len(weights) == 94
    for name, (W, beta, mean, var) in want.items():
        w = weights[name]
        assert np.array_equal(w["W"].numpy(), W) and np.array_equal(w["beta"].numpy(), beta)
        assert np.array_equal(w["mean"].numpy(), mean) and np.array_equal(w["var"].numpy(), var)
        assert "gamma" not in w
    consts = I.parse_graphdef_consts(graph)
    assert "pool_3/_reshape/shape" not in consts and "conv/Conv2D" not in consts
    # a kernel of the wrong shape under a unit's scope is an error, as is a missing constant
    bad = _pb_const_node("conv/conv2d_params", np.zeros((3, 3, 3, 16), np.float32)) + b"".join(nodes[1:])
    with pytest.raises(ValueError, match="Conv2d_1a_3x3"):
        I.weights_from_graphdef(bad)
    with pytest.raises(ValueError, match="not found"):
        I.weights_from_graphdef(b"".join(nodes[6:]))


def test_unsupported_kernel_size_fails_fast():
    """the CUDA kernels and the C-ABI are 5x5-only; any other `kernel_size` must raise at
    construction instead of reading 25 taps out of a smaller buffer."""
    from littlegan_b200 import model as M
    pargs = product_args(small_args())
    pargs.kernel_size = 3
    with pytest.raises(ValueError, match="kernel_size 5 only"):
        M.Encoder(pargs)
    with pytest.raises(ValueError, match="kernel_size 5 only"):
        M.Decoder(pargs)
    with pytest.raises(ValueError):
        M.Conv2D(8, 8, 5, 3)


def test_fid_without_inception_weights_raises():
    """`calculate_fid_given_paths(paths, None)` is the reference's default call (it downloads the
    model); here there is no download, and a silently random network would give a meaningless 'FID'."""
    from littlegan_b200 import fid
    with pytest.raises(RuntimeError, match="no Inception model file"):
        fid.create_inception_graph(None)
    with pytest.raises(RuntimeError, match="no Inception model file"):
        fid.create_inception_graph(fid.check_or_download_inception(None))
