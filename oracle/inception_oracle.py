"""CPU restatement (PyTorch-CPU, fp32 or fp64) of the Inception-2015 `pool_3:0` forward that the reference's
`fid.py:36-106` runs through a TF session (`create_inception_graph` + `_get_inception_layer` + `get_activations`).

TEST INFRASTRUCTURE - only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may import this;
`littlegan_b200/` never does.

PARITY UNPINNED against the real graph: `classify_image_graph_def.pb` is an external download (`fid.py:276-287`,
inception-2015-12-05.tgz) that is neither in the reference tree nor in this container, and TensorFlow is not
installable here.  What IS pinned: the wiring below equals torchvision's `inception_v3` (same published
architecture; `tests/test_oracle.py` loads identical weights into both and compares in `variant="torchvision"`).
`variant="fid"` then applies the three places where the 2015 TF graph is known to differ from torchvision's
model (as documented by the pytorch-fid port of that graph):
  * the 3x3 average pools of the Mixed_5/6/7b blocks exclude the zero padding from the divisor,
  * the pool branch of the LAST block (Mixed_7c) is a 3x3 MAX pool,
  * batch-norm epsilon is 1e-3 and the graph has no BN scale (gamma = 1).
Input handling follows `fid.py:73-106` + the graph's own pre-processing: images [b,H,W,3] in 0..255 ->
TF-1.x `ResizeBilinear` (align_corners=False, no half-pixel centres: src = dst * in/out) to 299x299 ->
(x - 128) / 128 -> network -> pool_3 = mean over the final 8x8 map -> [b, 2048].
"""
import math

import torch
import torch.nn.functional as F

BN_EPS = 1e-3

# (name, cin, cout, (kh, kw), stride, (ph, pw)) in forward order; names are torchvision's module names
STEM = [("Conv2d_1a_3x3", 3, 32, (3, 3), 2, (0, 0)), ("Conv2d_2a_3x3", 32, 32, (3, 3), 1, (0, 0)),
        ("Conv2d_2b_3x3", 32, 64, (3, 3), 1, (1, 1)), ("Conv2d_3b_1x1", 64, 80, (1, 1), 1, (0, 0)),
        ("Conv2d_4a_3x3", 80, 192, (3, 3), 1, (0, 0))]


def _a(name, cin, pool):
    return [(name + ".branch1x1", cin, 64, (1, 1), 1, (0, 0)),
            (name + ".branch5x5_1", cin, 48, (1, 1), 1, (0, 0)), (name + ".branch5x5_2", 48, 64, (5, 5), 1, (2, 2)),
            (name + ".branch3x3dbl_1", cin, 64, (1, 1), 1, (0, 0)),
            (name + ".branch3x3dbl_2", 64, 96, (3, 3), 1, (1, 1)),
            (name + ".branch3x3dbl_3", 96, 96, (3, 3), 1, (1, 1)),
            (name + ".branch_pool", cin, pool, (1, 1), 1, (0, 0))]


def _b(name, cin):
    return [(name + ".branch3x3", cin, 384, (3, 3), 2, (0, 0)),
            (name + ".branch3x3dbl_1", cin, 64, (1, 1), 1, (0, 0)),
            (name + ".branch3x3dbl_2", 64, 96, (3, 3), 1, (1, 1)),
            (name + ".branch3x3dbl_3", 96, 96, (3, 3), 2, (0, 0))]


def _c(name, c7):
    return [(name + ".branch1x1", 768, 192, (1, 1), 1, (0, 0)),
            (name + ".branch7x7_1", 768, c7, (1, 1), 1, (0, 0)), (name + ".branch7x7_2", c7, c7, (1, 7), 1, (0, 3)),
            (name + ".branch7x7_3", c7, 192, (7, 1), 1, (3, 0)),
            (name + ".branch7x7dbl_1", 768, c7, (1, 1), 1, (0, 0)),
            (name + ".branch7x7dbl_2", c7, c7, (7, 1), 1, (3, 0)), (name + ".branch7x7dbl_3", c7, c7, (1, 7), 1, (0, 3)),
            (name + ".branch7x7dbl_4", c7, c7, (7, 1), 1, (3, 0)),
            (name + ".branch7x7dbl_5", c7, 192, (1, 7), 1, (0, 3)),
            (name + ".branch_pool", 768, 192, (1, 1), 1, (0, 0))]


def _d(name):
    return [(name + ".branch3x3_1", 768, 192, (1, 1), 1, (0, 0)), (name + ".branch3x3_2", 192, 320, (3, 3), 2, (0, 0)),
            (name + ".branch7x7x3_1", 768, 192, (1, 1), 1, (0, 0)),
            (name + ".branch7x7x3_2", 192, 192, (1, 7), 1, (0, 3)),
            (name + ".branch7x7x3_3", 192, 192, (7, 1), 1, (3, 0)),
            (name + ".branch7x7x3_4", 192, 192, (3, 3), 2, (0, 0))]


def _e(name, cin):
    return [(name + ".branch1x1", cin, 320, (1, 1), 1, (0, 0)),
            (name + ".branch3x3_1", cin, 384, (1, 1), 1, (0, 0)),
            (name + ".branch3x3_2a", 384, 384, (1, 3), 1, (0, 1)), (name + ".branch3x3_2b", 384, 384, (3, 1), 1, (1, 0)),
            (name + ".branch3x3dbl_1", cin, 448, (1, 1), 1, (0, 0)),
            (name + ".branch3x3dbl_2", 448, 384, (3, 3), 1, (1, 1)),
            (name + ".branch3x3dbl_3a", 384, 384, (1, 3), 1, (0, 1)),
            (name + ".branch3x3dbl_3b", 384, 384, (3, 1), 1, (1, 0)),
            (name + ".branch_pool", cin, 192, (1, 1), 1, (0, 0))]


UNITS = (STEM + _a("Mixed_5b", 192, 32) + _a("Mixed_5c", 256, 64) + _a("Mixed_5d", 288, 64) + _b("Mixed_6a", 288)
         + _c("Mixed_6b", 128) + _c("Mixed_6c", 160) + _c("Mixed_6d", 160) + _c("Mixed_6e", 192) + _d("Mixed_7a")
         + _e("Mixed_7b", 1280) + _e("Mixed_7c", 2048))
SPEC = {u[0]: u for u in UNITS}


def random_weights(seed=0, dtype=torch.float32):
    """name -> dict(W [kh,kw,cin,cout] (TF HWIO), beta, mean, var [cout]).  He-normal kernels and near-identity BN
    statistics keep the activations O(1) through all 94 layers, so relative errors are meaningful at pool_3."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, cin, cout, (kh, kw), _, _ in UNITS:
        std = math.sqrt(2.0 / (kh * kw * cin))
        out[name] = dict(W=(torch.randn(kh, kw, cin, cout, generator=g) * std).to(dtype),
                         beta=(torch.randn(cout, generator=g) * 0.1).to(dtype),
                         mean=(torch.randn(cout, generator=g) * 0.1).to(dtype),
                         var=(torch.rand(cout, generator=g) * 0.5 + 0.75).to(dtype))
    return out


def resize_bilinear_tf1(x, out_h, out_w):
    """TF-1.x ResizeBilinear, align_corners=False (no half-pixel centres).  x [b,H,W,C] -> [b,out_h,out_w,C]."""
    b, H, W, C = x.shape

    def axis(n_in, n_out):
        src = torch.arange(n_out, dtype=x.dtype) * (n_in / n_out)
        lo = src.floor().long()
        hi = torch.clamp(lo + 1, max=n_in - 1)
        return lo, hi, src - lo.to(x.dtype)

    y0, y1, fy = axis(H, out_h)
    x0, x1, fx = axis(W, out_w)
    top = x[:, y0][:, :, x0] + (x[:, y0][:, :, x1] - x[:, y0][:, :, x0]) * fx[None, None, :, None]
    bot = x[:, y1][:, :, x0] + (x[:, y1][:, :, x1] - x[:, y1][:, :, x0]) * fx[None, None, :, None]
    return top + (bot - top) * fy[None, :, None, None]


class InceptionOracle:
    def __init__(self, weights, variant="fid", dtype=torch.float64, eps=BN_EPS):
        assert variant in ("fid", "torchvision")
        self.w = {k: {n: t.to(dtype) for n, t in v.items()} for k, v in weights.items()}
        self.fid, self.dtype, self.eps = variant == "fid", dtype, eps
        self.taps = {}                                   # name -> NHWC output of that conv unit (for layer parity)

    def unit(self, name, x):
        _, cin, cout, (kh, kw), s, (ph, pw) = SPEC[name]
        p = self.w[name]
        y = F.conv2d(x, p["W"].permute(3, 2, 0, 1), None, stride=s, padding=(ph, pw))
        gamma = p.get("gamma")
        inv = 1.0 / torch.sqrt(p["var"] + self.eps)
        if gamma is not None:
            inv = inv * gamma
        y = (y - p["mean"][None, :, None, None]) * inv[None, :, None, None] + p["beta"][None, :, None, None]
        y = F.relu(y)
        self.taps[name] = y.permute(0, 2, 3, 1)
        return y

    def _avg3(self, x):
        return F.avg_pool2d(x, 3, stride=1, padding=1, count_include_pad=not self.fid)

    def block_a(self, n, x):
        u = self.unit
        return torch.cat([u(n + ".branch1x1", x), u(n + ".branch5x5_2", u(n + ".branch5x5_1", x)),
                          u(n + ".branch3x3dbl_3", u(n + ".branch3x3dbl_2", u(n + ".branch3x3dbl_1", x))),
                          u(n + ".branch_pool", self._avg3(x))], 1)

    def block_b(self, n, x):
        u = self.unit
        return torch.cat([u(n + ".branch3x3", x),
                          u(n + ".branch3x3dbl_3", u(n + ".branch3x3dbl_2", u(n + ".branch3x3dbl_1", x))),
                          F.max_pool2d(x, 3, stride=2)], 1)

    def block_c(self, n, x):
        u = self.unit
        b7 = u(n + ".branch7x7_3", u(n + ".branch7x7_2", u(n + ".branch7x7_1", x)))
        d = x
        for i in range(1, 6):
            d = u(n + ".branch7x7dbl_%d" % i, d)
        return torch.cat([u(n + ".branch1x1", x), b7, d, u(n + ".branch_pool", self._avg3(x))], 1)

    def block_d(self, n, x):
        u = self.unit
        b7 = x
        for i in range(1, 5):
            b7 = u(n + ".branch7x7x3_%d" % i, b7)
        return torch.cat([u(n + ".branch3x3_2", u(n + ".branch3x3_1", x)), b7, F.max_pool2d(x, 3, stride=2)], 1)

    def block_e(self, n, x, max_pool):
        u = self.unit
        b3 = u(n + ".branch3x3_1", x)
        b3 = torch.cat([u(n + ".branch3x3_2a", b3), u(n + ".branch3x3_2b", b3)], 1)
        d = u(n + ".branch3x3dbl_2", u(n + ".branch3x3dbl_1", x))
        d = torch.cat([u(n + ".branch3x3dbl_3a", d), u(n + ".branch3x3dbl_3b", d)], 1)
        pool = F.max_pool2d(x, 3, stride=1, padding=1) if max_pool else self._avg3(x)
        return torch.cat([u(n + ".branch1x1", x), b3, d, u(n + ".branch_pool", pool)], 1)

    def features(self, x):
        """x [b,3,299,299] already normalised -> pool_3 [b,2048]."""
        for n in ("Conv2d_1a_3x3", "Conv2d_2a_3x3", "Conv2d_2b_3x3"):
            x = self.unit(n, x)
        x = F.max_pool2d(x, 3, stride=2)
        x = self.unit("Conv2d_4a_3x3", self.unit("Conv2d_3b_1x1", x))
        x = F.max_pool2d(x, 3, stride=2)
        for n in ("Mixed_5b", "Mixed_5c", "Mixed_5d"):
            x = self.block_a(n, x)
        x = self.block_b("Mixed_6a", x)
        for n in ("Mixed_6b", "Mixed_6c", "Mixed_6d", "Mixed_6e"):
            x = self.block_c(n, x)
        x = self.block_d("Mixed_7a", x)
        x = self.block_e("Mixed_7b", x, False)
        x = self.block_e("Mixed_7c", x, self.fid)
        self.taps["Mixed_7c"] = x.permute(0, 2, 3, 1)
        return x.mean(dim=(2, 3))

    def preprocess(self, images):
        """[b,H,W,3] in 0..255 -> [b,3,299,299] in about (-1,1)."""
        x = resize_bilinear_tf1(torch.as_tensor(images).to(self.dtype), 299, 299)
        self.taps["input"] = (x - 128.0) / 128.0
        return self.taps["input"].permute(0, 3, 1, 2)

    def __call__(self, images):
        with torch.no_grad():
            return self.features(self.preprocess(images))
