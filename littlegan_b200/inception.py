"""Inception-2015 `pool_3:0` feature extractor for FID (reference `fid.py:36-106`: `create_inception_graph`,
`_get_inception_layer`, `get_activations`).  An instance is the `sess` argument of
`fid.get_activations` / `fid.calculate_activation_statistics`: images [b,H,W,3] in 0..255 -> CUDA fp32 [b,2048].

The reference downloads `classify_image_graph_def.pb` (`fid.py:276-287`); there is no network here, so the weights
are random unless `load_npz` is given converted ones (keys `<unit>/W` HWIO, `<unit>/beta`, `/mean`, `/var`, optional
`/gamma`).  The graph is a program of conv+BN+ReLU units, pools and concats over NHWC tensors; every branch of a
Mixed block writes its channel slice of the block output directly (no concat copies), batch-norm is folded into a
per-channel scale / shift in the conv epilogue, and the input stage (TF-1.x bilinear resize to 299x299 and
(x - 128) / 128) is one kernel that can read the image bytes as they are.
"""
import math

import numpy as np
import torch

from . import kernels as K
from ._lib import LittleGANError

BN_EPS = 1e-3
MAXP, AVGP = "max", "avg"


def _conv(name, cin, cout, k, s=1, p=(0, 0)):
    k = (k, k) if isinstance(k, int) else k
    p = (p, p) if isinstance(p, int) else p
    return dict(name=name, cin=cin, cout=cout, k=k, s=s, p=p)


def _block_a(n, cin, pool):
    return [[_conv(n + ".branch1x1", cin, 64, 1)],
            [_conv(n + ".branch5x5_1", cin, 48, 1), _conv(n + ".branch5x5_2", 48, 64, 5, p=2)],
            [_conv(n + ".branch3x3dbl_1", cin, 64, 1), _conv(n + ".branch3x3dbl_2", 64, 96, 3, p=1),
             _conv(n + ".branch3x3dbl_3", 96, 96, 3, p=1)],
            [(AVGP, 3, 1, 1), _conv(n + ".branch_pool", cin, pool, 1)]]


def _block_b(n, cin):
    return [[_conv(n + ".branch3x3", cin, 384, 3, s=2)],
            [_conv(n + ".branch3x3dbl_1", cin, 64, 1), _conv(n + ".branch3x3dbl_2", 64, 96, 3, p=1),
             _conv(n + ".branch3x3dbl_3", 96, 96, 3, s=2)],
            [(MAXP, 3, 2, 0)]]


def _block_c(n, c7):
    return [[_conv(n + ".branch1x1", 768, 192, 1)],
            [_conv(n + ".branch7x7_1", 768, c7, 1), _conv(n + ".branch7x7_2", c7, c7, (1, 7), p=(0, 3)),
             _conv(n + ".branch7x7_3", c7, 192, (7, 1), p=(3, 0))],
            [_conv(n + ".branch7x7dbl_1", 768, c7, 1), _conv(n + ".branch7x7dbl_2", c7, c7, (7, 1), p=(3, 0)),
             _conv(n + ".branch7x7dbl_3", c7, c7, (1, 7), p=(0, 3)), _conv(n + ".branch7x7dbl_4", c7, c7, (7, 1), p=(3, 0)),
             _conv(n + ".branch7x7dbl_5", c7, 192, (1, 7), p=(0, 3))],
            [(AVGP, 3, 1, 1), _conv(n + ".branch_pool", 768, 192, 1)]]


def _block_d(n):
    return [[_conv(n + ".branch3x3_1", 768, 192, 1), _conv(n + ".branch3x3_2", 192, 320, 3, s=2)],
            [_conv(n + ".branch7x7x3_1", 768, 192, 1), _conv(n + ".branch7x7x3_2", 192, 192, (1, 7), p=(0, 3)),
             _conv(n + ".branch7x7x3_3", 192, 192, (7, 1), p=(3, 0)), _conv(n + ".branch7x7x3_4", 192, 192, 3, s=2)],
            [(MAXP, 3, 2, 0)]]


def _block_e(n, cin, pool):
    # a branch may end in a fork: a list of parallel units whose outputs are concatenated in order
    fork = lambda stem: [_conv(n + stem + "a", 384, 384, (1, 3), p=(0, 1)), _conv(n + stem + "b", 384, 384, (3, 1), p=(1, 0))]
    return [[_conv(n + ".branch1x1", cin, 320, 1)],
            [_conv(n + ".branch3x3_1", cin, 384, 1), fork(".branch3x3_2")],
            [_conv(n + ".branch3x3dbl_1", cin, 448, 1), _conv(n + ".branch3x3dbl_2", 448, 384, 3, p=1),
             fork(".branch3x3dbl_3")],
            [(pool, 3, 1, 1), _conv(n + ".branch_pool", cin, 192, 1)]]


def program(fid_variant=True):
    """The graph as a list of stages: a conv dict, a pool tuple (kind, k, stride, pad), or a list of branches."""
    return ([_conv("Conv2d_1a_3x3", 3, 32, 3, s=2), _conv("Conv2d_2a_3x3", 32, 32, 3), _conv("Conv2d_2b_3x3", 32, 64, 3, p=1),
             (MAXP, 3, 2, 0), _conv("Conv2d_3b_1x1", 64, 80, 1), _conv("Conv2d_4a_3x3", 80, 192, 3), (MAXP, 3, 2, 0)]
            + [_block_a("Mixed_5b", 192, 32), _block_a("Mixed_5c", 256, 64), _block_a("Mixed_5d", 288, 64),
               _block_b("Mixed_6a", 288), _block_c("Mixed_6b", 128), _block_c("Mixed_6c", 160), _block_c("Mixed_6d", 160),
               _block_c("Mixed_6e", 192), _block_d("Mixed_7a"), _block_e("Mixed_7b", 1280, AVGP),
               # the 2015 graph's last block pools with MAX (the torchvision model averages)
               _block_e("Mixed_7c", 2048, MAXP if fid_variant else AVGP)])


def _units(prog):
    for st in prog:
        if isinstance(st, dict):
            yield st
        elif isinstance(st, list):
            for u in _units(st):
                yield u


def unit_specs(fid_variant=True):
    return list(_units(program(fid_variant)))


def random_weights(seed=0):
    """He-normal kernels, near-identity BN statistics (there are no pretrained weights in the container)."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for u in unit_specs():
        kh, kw = u["k"]
        std = math.sqrt(2.0 / (kh * kw * u["cin"]))
        out[u["name"]] = dict(W=torch.randn(kh, kw, u["cin"], u["cout"], generator=g) * std,
                              beta=torch.randn(u["cout"], generator=g) * 0.1,
                              mean=torch.randn(u["cout"], generator=g) * 0.1,
                              var=torch.rand(u["cout"], generator=g) * 0.5 + 0.75)
    return out


def load_npz(path):
    data = np.load(path)
    out = {}
    for u in unit_specs():
        n = u["name"]
        out[n] = {k: torch.from_numpy(np.asarray(data[n + "/" + k], np.float32)) for k in ("W", "beta", "mean", "var")}
        if n + "/gamma" in data:
            out[n]["gamma"] = torch.from_numpy(np.asarray(data[n + "/gamma"], np.float32))
    return out


def _out_hw(h, w, k, s, p):
    return (h + 2 * p[0] - k[0]) // s + 1, (w + 2 * p[1] - k[1]) // s + 1


class InceptionPool3:
    def __init__(self, weights=None, seed=0, dtype="fp32", device=None, fid_variant=True, avg_excludes_pad=None):
        if not torch.cuda.is_available():
            raise LittleGANError("littlegan_b200 has no CPU path: a CUDA device is required")
        self.device = torch.device(device or "cuda")
        self.act_dtype = {"fp32": torch.float32, "bf16": torch.bfloat16}[dtype]
        self.prog = program(fid_variant)
        excl = fid_variant if avg_excludes_pad is None else avg_excludes_pad
        self.avg_mode = K.POOL_AVG_VALID if excl else K.POOL_AVG_PADDED
        weights = random_weights(seed) if weights is None else weights
        self.in_pad = 8 if self.act_dtype == torch.bfloat16 else 3      # channels per pixel of the resized input
        self.params = {}
        for u in _units(self.prog):
            p = weights[u["name"]]
            W = p["W"].to(torch.float32)
            if tuple(W.shape) != (u["k"][0], u["k"][1], u["cin"], u["cout"]):
                raise ValueError("%s: kernel shape %s" % (u["name"], tuple(W.shape)))
            scale = 1.0 / torch.sqrt(p["var"].double() + BN_EPS)
            if p.get("gamma") is not None:
                scale = scale * p["gamma"].double()
            shift = p["beta"].double() - p["mean"].double() * scale
            if self.act_dtype == torch.bfloat16 and u["cin"] % 8 != 0:
                # the RGB stem: zero kernel rows for the zero channels the input stage pads each pixel with
                W = torch.cat([W, W.new_zeros(W.shape[0], W.shape[1], self.in_pad - u["cin"], W.shape[3])], 2)
            W, scale, shift = (t.to(self.device, torch.float32).contiguous() for t in (W, scale, shift))
            # bf16 mode: the unit's kernel as the tensor-core operand, packed once (None: SIMT path, e.g. the RGB stem)
            wpack = K.pack_conv_bn_weights(W) if self.act_dtype == torch.bfloat16 else None
            self.params[u["name"]] = (W, scale, shift, wpack)
        self.taps = None                      # set to a dict to keep every unit's output (layer-parity tests)

    # ------------------------------------------------------------------ executor
    def _new(self, n, h, w, c):
        return torch.empty(n, h, w, c, dtype=self.act_dtype, device=self.device)

    def _run_conv(self, u, x, y=None, y_off=0):
        W, scale, shift, wpack = self.params[u["name"]]
        if y is None:
            ho, wo = _out_hw(x.shape[1], x.shape[2], u["k"], u["s"], u["p"])
            y = self._new(x.shape[0], ho, wo, u["cout"])
        K.conv2d_bn_relu(x, W, scale, shift, y, y_off, stride=u["s"], pad=u["p"], wpack=wpack)
        if self.taps is not None:
            self.taps[u["name"]] = y[..., y_off:y_off + u["cout"]]
        return y

    def _run_pool(self, st, x, y=None, y_off=0):
        kind, k, s, p = st
        if y is None:
            ho, wo = _out_hw(x.shape[1], x.shape[2], (k, k), s, (p, p))
            y = self._new(x.shape[0], ho, wo, x.shape[3])
        return K.pool2d(x, y, y_off, k, s, p, K.POOL_MAX if kind == MAXP else self.avg_mode)

    @staticmethod
    def _width(st, cin):
        if isinstance(st, dict):
            return st["cout"]
        if isinstance(st, tuple):
            return cin
        return sum(u["cout"] for u in st)      # fork

    def _run_block(self, branches, x):
        n, h, w, cin = x.shape
        widths, hw = [], None
        for br in branches:
            c, hh, ww = cin, h, w
            for st in br:
                c = self._width(st, c)
                if isinstance(st, dict):
                    hh, ww = _out_hw(hh, ww, st["k"], st["s"], st["p"])
                elif isinstance(st, tuple):
                    hh, ww = _out_hw(hh, ww, (st[1], st[1]), st[2], (st[3], st[3]))
            widths.append(c)
            hw = (hh, ww)
        out = self._new(n, hw[0], hw[1], sum(widths))
        off = 0
        for br, width in zip(branches, widths):
            t = x
            for i, st in enumerate(br):
                last = i == len(br) - 1
                if isinstance(st, list):           # fork: always the tail of its branch
                    o = off
                    for u in st:
                        self._run_conv(u, t, out, o)
                        o += u["cout"]
                elif isinstance(st, dict):
                    t = self._run_conv(st, t, out, off) if last else self._run_conv(st, t)
                else:
                    t = self._run_pool(st, t, out, off) if last else self._run_pool(st, t)
            off += width
        return out

    @torch.no_grad()
    def features_from_normalised(self, x):
        """x [b,299,299,3] NHWC already resized and normalised -> pool_3 [b,2048] fp32."""
        if x.shape[3] < self.in_pad:
            x = torch.nn.functional.pad(x, (0, self.in_pad - x.shape[3])).contiguous()
        for st in self.prog:
            if isinstance(st, dict):
                x = self._run_conv(st, x)
            elif isinstance(st, tuple):
                x = self._run_pool(st, x)
            else:
                x = self._run_block(st, x)
        if self.taps is not None:
            self.taps["Mixed_7c"] = x
        out = torch.empty(x.shape[0], x.shape[3], dtype=torch.float32, device=self.device)
        return K.global_avgpool(x, out)

    @torch.no_grad()
    def preprocess(self, images):
        from .utils import upload
        x = upload(images, self.device)
        if x.dtype != torch.uint8:
            x = x.to(torch.float32)
        x = x.contiguous()
        y = self._new(x.shape[0], 299, 299, max(self.in_pad, x.shape[3]))
        K.resize_bilinear_norm(x, y, 128.0, 1.0 / 128.0)
        if self.taps is not None:
            self.taps["input"] = y[..., :x.shape[3]]
        return y

    def __call__(self, images):
        """images [b,H,W,3] in 0..255 (numpy / torch, uint8 or float) -> CUDA fp32 [b,2048]."""
        return self.features_from_normalised(self.preprocess(images))


# ------------------------------------------------------------------ weights from the reference's own model file
# The reference feeds `classify_image_graph_def.pb` (inception-2015-12-05.tgz, fid.py:276-287) to TensorFlow.  Without
# TensorFlow the file is still just a protobuf: the functions below read its Const nodes with a minimal wire-format
# parser and map the graph's scopes onto the unit names used here, so a user who has the file gets the real network.
def _pb_fields(buf):
    """(field number, wire type, value) of one protobuf message; length-delimited values are memoryviews."""
    buf = memoryview(buf)
    i, n = 0, len(buf)
    while i < n:
        key = shift = 0
        while True:
            b = buf[i]; i += 1
            key |= (b & 0x7F) << shift
            shift += 7
            if not b & 0x80:
                break
        field, wt = key >> 3, key & 7
        if wt == 0:
            v = shift = 0
            while True:
                b = buf[i]; i += 1
                v |= (b & 0x7F) << shift
                shift += 7
                if not b & 0x80:
                    break
        elif wt == 1:
            v = buf[i:i + 8]; i += 8
        elif wt == 5:
            v = buf[i:i + 4]; i += 4
        elif wt == 2:
            ln = shift = 0
            while True:
                b = buf[i]; i += 1
                ln |= (b & 0x7F) << shift
                shift += 7
                if not b & 0x80:
                    break
            v = buf[i:i + ln]; i += ln
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        yield field, wt, v


def _pb_tensor(buf):
    """TensorProto -> float32 ndarray (DT_FLOAT only; tensor_content or float_val), or None for other dtypes."""
    dtype, shape, content, floats = 0, [], None, []
    for f, wt, v in _pb_fields(buf):
        if f == 1 and wt == 0:
            dtype = v
        elif f == 2 and wt == 2:                               # TensorShapeProto { repeated Dim dim = 2 }
            for f2, wt2, v2 in _pb_fields(v):
                if f2 == 2 and wt2 == 2:
                    size = 0
                    for f3, wt3, v3 in _pb_fields(v2):
                        if f3 == 1 and wt3 == 0:
                            size = v3
                    shape.append(size)
        elif f == 4 and wt == 2:
            content = bytes(v)
        elif f == 5 and wt == 2:                               # packed repeated float
            floats.append(np.frombuffer(bytes(v), dtype="<f4"))
        elif f == 5 and wt == 5:
            floats.append(np.frombuffer(bytes(v), dtype="<f4"))
    if dtype != 1:
        return None
    count = int(np.prod(shape)) if shape else 1
    if content is not None:
        arr = np.frombuffer(content, dtype="<f4")
    else:
        arr = np.concatenate(floats) if floats else np.zeros(0, np.float32)
        if arr.size == 1 and count > 1:                        # TF stores a constant-filled tensor as one value
            arr = np.full(count, arr[0], np.float32)
    if arr.size != count:
        raise ValueError("tensor has %d values for shape %s" % (arr.size, shape))
    return arr.astype(np.float32).reshape(shape)


def parse_graphdef_consts(buf, wanted=None):
    """name -> float32 array of the Const nodes of a serialized GraphDef (optionally only the names in `wanted`)."""
    out = {}
    for f, wt, node in _pb_fields(buf):
        if f != 1 or wt != 2:                                  # GraphDef { repeated NodeDef node = 1 }
            continue
        name, op, tensor = None, None, None
        for f2, wt2, v2 in _pb_fields(node):
            if f2 == 1 and wt2 == 2:
                name = bytes(v2).decode()
            elif f2 == 2 and wt2 == 2:
                op = bytes(v2).decode()
            elif f2 == 5 and wt2 == 2:                         # map<string, AttrValue> attr: entry {key = 1, value = 2}
                key, val = None, None
                for f3, wt3, v3 in _pb_fields(v2):
                    if f3 == 1 and wt3 == 2:
                        key = bytes(v3).decode()
                    elif f3 == 2 and wt3 == 2:
                        val = v3
                if key == "value" and val is not None:
                    for f4, wt4, v4 in _pb_fields(val):        # AttrValue { TensorProto tensor = 8 }
                        if f4 == 8 and wt4 == 2:
                            tensor = v4
        if op == "Const" and tensor is not None and (wanted is None or name in wanted):
            arr = _pb_tensor(tensor)
            if arr is not None:
                out[name] = arr
    return out


def graphdef_scope(unit_name):
    """Scope of a unit in the 2015 graph: the stem units are conv, conv_1 .. conv_4; Mixed_5b .. Mixed_7c are mixed,
    mixed_1 .. mixed_10 with the branches as towers (tower, tower_1, tower_2) and the 1x3 / 3x1 forks as tower*/mixed."""
    stem = {"Conv2d_1a_3x3": "conv", "Conv2d_2a_3x3": "conv_1", "Conv2d_2b_3x3": "conv_2", "Conv2d_3b_1x1": "conv_3",
            "Conv2d_4a_3x3": "conv_4"}
    if unit_name in stem:
        return stem[unit_name]
    block, branch = unit_name.split(".")
    order = ["Mixed_5b", "Mixed_5c", "Mixed_5d", "Mixed_6a", "Mixed_6b", "Mixed_6c", "Mixed_6d", "Mixed_6e", "Mixed_7a",
             "Mixed_7b", "Mixed_7c"]
    idx = order.index(block)
    scope = "mixed" if idx == 0 else "mixed_%d" % idx
    nth = lambda i: "conv" if i == 0 else "conv_%d" % i
    kind = "A" if idx < 3 else "B" if idx == 3 else "C" if idx < 8 else "D" if idx == 8 else "E"
    table = {
        "A": {"branch1x1": "conv", "branch5x5_1": "tower/conv", "branch5x5_2": "tower/conv_1",
              "branch3x3dbl_1": "tower_1/conv", "branch3x3dbl_2": "tower_1/conv_1", "branch3x3dbl_3": "tower_1/conv_2",
              "branch_pool": "tower_2/conv"},
        "B": {"branch3x3": "conv", "branch3x3dbl_1": "tower/conv", "branch3x3dbl_2": "tower/conv_1",
              "branch3x3dbl_3": "tower/conv_2"},
        "C": dict([("branch1x1", "conv"), ("branch_pool", "tower_2/conv")]
                  + [("branch7x7_%d" % (i + 1), "tower/" + nth(i)) for i in range(3)]
                  + [("branch7x7dbl_%d" % (i + 1), "tower_1/" + nth(i)) for i in range(5)]),
        "D": dict([("branch3x3_%d" % (i + 1), "tower/" + nth(i)) for i in range(2)]
                  + [("branch7x7x3_%d" % (i + 1), "tower_1/" + nth(i)) for i in range(4)]),
        "E": {"branch1x1": "conv", "branch3x3_1": "tower/conv", "branch3x3_2a": "tower/mixed/conv",
              "branch3x3_2b": "tower/mixed/conv_1", "branch3x3dbl_1": "tower_1/conv", "branch3x3dbl_2": "tower_1/conv_1",
              "branch3x3dbl_3a": "tower_1/mixed/conv", "branch3x3dbl_3b": "tower_1/mixed/conv_1",
              "branch_pool": "tower_2/conv"},
    }
    return scope + "/" + table[kind][branch]


def weights_from_graphdef(buf):
    """The weight dict of `InceptionPool3` from the bytes of `classify_image_graph_def.pb`.  Every kernel's shape is
    checked against the unit it is mapped to.  The graph normalises with scale_after_normalization = false, so its
    gamma constants are not applied."""
    specs = unit_specs()
    wanted = set()
    for u in specs:
        sc = graphdef_scope(u["name"])
        wanted.update([sc + "/conv2d_params", sc + "/batchnorm/beta", sc + "/batchnorm/moving_mean",
                       sc + "/batchnorm/moving_variance"])
    consts = parse_graphdef_consts(buf, wanted)
    out = {}
    for u in specs:
        sc = graphdef_scope(u["name"])
        try:
            W = consts[sc + "/conv2d_params"]
            beta, mean, var = (consts[sc + "/batchnorm/" + k] for k in ("beta", "moving_mean", "moving_variance"))
        except KeyError as e:
            raise ValueError("%s: constant %s not found in the graph" % (u["name"], e)) from None
        want = (u["k"][0], u["k"][1], u["cin"], u["cout"])
        if tuple(W.shape) != want or any(t.reshape(-1).shape != (u["cout"],) for t in (beta, mean, var)):
            raise ValueError("%s (%s): kernel %s, expected %s" % (u["name"], sc, tuple(W.shape), want))
        out[u["name"]] = dict(W=torch.from_numpy(W.copy()), beta=torch.from_numpy(beta.reshape(-1).copy()),
                              mean=torch.from_numpy(mean.reshape(-1).copy()), var=torch.from_numpy(var.reshape(-1).copy()))
    return out


def load_graphdef(path):
    """`classify_image_graph_def.pb`, or the inception-2015-12-05.tgz that contains it (fid.py:276-287)."""
    path = str(path)
    if path.endswith((".tgz", ".tar.gz")):
        import tarfile
        with tarfile.open(path, mode="r") as tf_:
            buf = tf_.extractfile("classify_image_graph_def.pb").read()
    else:
        with open(path, "rb") as f:
            buf = f.read()
    return weights_from_graphdef(buf)
