"""CPU restatement of the LittleGAN train-step / predict arithmetic (PyTorch-CPU).

TEST INFRASTRUCTURE - never imported by `littlegan_b200/`.  PARITY UNPINNED
(see oracle/__init__.py): the reference has no golden vectors and TensorFlow
is not installable here, so this file restates the TF-1.15 semantics the
reference relies on and is checked by definition-level loops in
tests/test_oracle.py.

All activations are NHWC, all kernels are in TensorFlow layouts:
  Conv2D kernel           [kh, kw, in, out]          (HWIO)
  Conv2DTranspose kernel  [kh, kw, out, in]
  Dense kernel            [in, out]
Gradients come from torch.autograd on this restatement (the reference uses
tf.GradientTape, eager_trainer.py:133-163).
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------- #
# hyper-parameters (sample.config.json:1-53)
# --------------------------------------------------------------------------- #
SAMPLE_CONFIG = dict(
    batch_size=32, image_channel=3, image_dim=128, noise_dim=93, init_dim=8,
    conv_filter=[384, 256, 128, 64, 32], kernel_size=5, leaky_alpha=0.3,
    dropout_rate=0.5, l1_lambda=0.02, lr=5e-5, beta_1=0.5, beta_2=0.9,
    use_gp=False, use_clip=True, clip_range=0.5, use_partition=True,
    partition_interval=4, train_adj=True, cond_dim=7,
)


def make_args(**over):
    d = dict(SAMPLE_CONFIG)
    d.update(over)
    return SimpleNamespace(**d)


def soft(x):
    """utils.py:47-48"""
    return 0.96 * x + 0.02


# --------------------------------------------------------------------------- #
# primitive ops with TF semantics
# --------------------------------------------------------------------------- #
def _same_pads(size, k, s):
    """TF 'SAME': out = ceil(size/s); total = max((out-1)*s + k - size, 0);
    before = total // 2, after = total - before."""
    out = -(-size // s)
    total = max((out - 1) * s + k - size, 0)
    return total // 2, total - total // 2


def conv2d_same(x, w, b, stride):
    """tf.layers.Conv2D(filters, k, stride, 'same') (model.py:15).
    x [N,H,W,Cin], w [kh,kw,Cin,Cout], b [Cout].
    y[n,oy,ox,co] = b[co] + sum x[n, s*oy+ky-pt, s*ox+kx-pl, ci] * w[ky,kx,ci,co]
    with (pt, pl) = TF SAME 'before' pads (1 for s=2,k=5, even H)."""
    kh, kw = w.shape[0], w.shape[1]
    pt, pb = _same_pads(x.shape[1], kh, stride)
    pl, pr = _same_pads(x.shape[2], kw, stride)
    xn = F.pad(x.permute(0, 3, 1, 2), (pl, pr, pt, pb))
    y = F.conv2d(xn, w.permute(3, 2, 0, 1), b, stride=stride)
    return y.permute(0, 2, 3, 1)


def conv2d_transpose_same(x, w, b, stride):
    """tf.layers.Conv2DTranspose(filters, k, stride, 'same') (model.py:39,86).
    x [N,H,W,Cin], w [kh,kw,Cout,Cin], b [Cout]; output [N, s*H, s*W, Cout].
    TF implements it as Conv2DBackpropInput of the SAME forward conv that maps
    (s*H) -> H, i.e. the full transposed output (index s*i + ky) cropped by the
    forward conv's 'before' pad:
        y[n,Y,X,co] = b[co] + sum_{s*i+ky-pt == Y} x[n,i,j,ci] * w[ky,kx,co,ci]."""
    kh, kw = w.shape[0], w.shape[1]
    H, W = x.shape[1], x.shape[2]
    pt, _ = _same_pads(H * stride, kh, stride)
    pl, _ = _same_pads(W * stride, kw, stride)
    full = F.conv_transpose2d(x.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), None, stride=stride)
    y = full[:, :, pt:pt + H * stride, pl:pl + W * stride]
    return y.permute(0, 2, 3, 1) + b


def instance_norm(x, gamma, beta, eps=1e-3):
    """InstanceNormalization(axis=None) (instance.py:105-128): per-sample
    statistics over ALL non-batch axes, scalar gamma/beta of shape (1,),
    epsilon added to the standard deviation (instance.py:115)."""
    axes = tuple(range(1, x.dim()))
    mean = x.mean(dim=axes, keepdim=True)
    std = ((x - mean) ** 2).mean(dim=axes, keepdim=True).sqrt() + eps
    return (x - mean) / std * gamma.reshape([1] * x.dim()) + beta.reshape([1] * x.dim())


def leaky(x, alpha):
    """tf.nn.leaky_relu (model.py:24,50,100,130); grad at 0 is alpha."""
    return F.leaky_relu(x, alpha)


def bce(target, output):
    """tf.keras.losses.binary_crossentropy (TF 1.15 eager branch) followed by
    tf.reduce_mean (eager_trainer.py:87-101): clip to [1e-7, 1-1e-7], then
    -(t*log(p+1e-7) + (1-t)*log(1-p+1e-7)), mean over the last axis, then mean
    over the batch.  Targets may lie outside [0,1] (soft labels are -0.94/0.98)."""
    eps = 1e-7
    one_m = torch.tensor(1.0, dtype=output.dtype) - torch.tensor(eps, dtype=output.dtype)
    p = torch.clamp(output, eps, float(one_m))
    v = -(target * torch.log(p + eps) + (1 - target) * torch.log(1 - p + eps))
    return v.mean(dim=-1).mean()


# --------------------------------------------------------------------------- #
# weights
# --------------------------------------------------------------------------- #
def _glorot(shape, fan_in, fan_out, gen, dtype):
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return (torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1).mul_(lim).to(dtype)


def init_weights(args, seed=0, dtype=torch.float32):
    """Keras/TF-layers defaults: Glorot-uniform kernels, zero biases, gamma=1,
    beta=0.  Returns dict(D=[20], G=[22], A=[4]) in `model.weights` order
    (SURVEY 8 a3/a4/a6): conv_i(k,b) norm_i(g,b) ...; dense(k,b)."""
    gen = torch.Generator().manual_seed(seed)
    k = args.kernel_size
    cf = args.conv_filter
    ch = args.image_channel
    flat = args.init_dim ** 2 * cf[0]
    one = lambda: torch.ones(1, dtype=dtype)
    zero = lambda n=1: torch.zeros(n, dtype=dtype)

    D = []
    cin = ch
    for i in range(1, 5):
        cout = cf[4 - i]
        D += [_glorot((k, k, cin, cout), k * k * cin, k * k * cout, gen, dtype), zero(cout), one(), zero()]
        cin = cout
    D += [_glorot((flat, 1), flat, 1, gen, dtype), zero(1)]
    D += [_glorot((flat, args.cond_dim), flat, args.cond_dim, gen, dtype), zero(args.cond_dim)]

    gin = args.noise_dim + args.cond_dim
    G = [_glorot((gin, flat), gin, flat, gen, dtype), zero(flat), one(), zero()]
    cin = cf[0]
    for i in range(1, 5):
        cout = cf[i]
        # Conv2DTranspose kernel [k,k,out,in]; Keras fans: in = k*k*in_ch ... computed on
        # the kernel shape as (receptive * shape[-2], receptive * shape[-1]).
        G += [_glorot((k, k, cout, cin), k * k * cout, k * k * cin, gen, dtype), zero(cout), one(), zero()]
        cin = cout
    G += [_glorot((k, k, ch, cin), k * k * ch, k * k * cin, gen, dtype), zero(ch)]

    A = [_glorot((args.cond_dim, flat), args.cond_dim, flat, gen, dtype), zero(flat), one(), zero()]
    return dict(D=D, G=G, A=A)


# --------------------------------------------------------------------------- #
# models (model.py)
# --------------------------------------------------------------------------- #
class _RoundBF16(torch.autograd.Function):
    """Storage emulation for the bf16-mode parity tests: the value AND the gradient that flows back through this
    point are rounded to bf16 (round-to-nearest-even), as a tensor that the product keeps in HBM as bf16 is - the
    forward map it stores, and the gradient with respect to it that the backward pass stores."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


class _RoundGradBF16(torch.autograd.Function):
    """Identity forward; the gradient is rounded to bf16 (a gradient the product stores without ever storing the
    forward value at that point, e.g. the pre-tanh gradient of the final layer)."""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


def _q(args):
    """args.store == "bf16": emulate the product's bf16 map storage (see _RoundBF16); default: exact arithmetic."""
    return _RoundBF16.apply if getattr(args, "store", None) == "bf16" else (lambda x: x)


def _qg(args):
    return _RoundGradBF16.apply if getattr(args, "store", None) == "bf16" else (lambda x: x)


def _tap(taps, name, x):
    """Record a per-layer activation for the parity tests (`taps`: dict or None)."""
    if taps is not None:
        taps[name] = x.detach()


def encoder(args, x, W, taps=None, tag="enc"):
    """Encoder.call (model.py:18-27). W = 16 tensors. Dropout is identity
    (tf.layers.dropout default training=False)."""
    outs = []
    q = _q(args)
    for i in range(4):
        k, b, g, be = W[4 * i:4 * i + 4]
        x = q(conv2d_same(x, k, b, 2))
        x = instance_norm(x, g, be)
        x = q(leaky(x, args.leaky_alpha))
        _tap(taps, "%s%d" % (tag, i + 1), x)
        outs.append(x)
    return outs


def decoder(args, x, add, W, taps=None, tag="dec"):
    """Decoder.call (model.py:43-51). W = 16 tensors."""
    q = _q(args)
    for i in range(4):
        if add[i] is not None:
            x = q(x + add[i])          # storage emulation: the product stores the SUM (one rounding)
        k, b, g, be = W[4 * i:4 * i + 4]
        x = q(conv2d_transpose_same(x, k, b, 2))
        x = instance_norm(x, g, be)
        x = leaky(x, args.leaky_alpha)
        if i == 3 or add[i + 1] is None:
            x = q(x)
        _tap(taps, "%s%d" % (tag, i + 1), x)
    return x


def discriminator(args, x, WD, taps=None, tag="d_"):
    """Discriminator.call (model.py:66-73) -> (pr [B,1], c [B,cond])."""
    feats = encoder(args, x, WD[:16], taps, tag + "enc")
    f = feats[-1].reshape(feats[-1].shape[0], -1)
    pr = torch.sigmoid(f @ WD[16] + WD[17])
    c = torch.sigmoid(f @ WD[18] + WD[19])
    _tap(taps, tag + "pr", pr)
    _tap(taps, tag + "c", c)
    return pr, c


def generator(args, noise, cond, WG, taps=None, tag="g_"):
    """Generator.call (model.py:90-105)."""
    x = torch.cat([noise, cond], dim=-1) @ WG[0] + WG[1]
    x = leaky(x, args.leaky_alpha)
    x = x.reshape(-1, args.init_dim, args.init_dim, args.conv_filter[0])
    x = _q(args)(instance_norm(x, WG[2], WG[3]))
    _tap(taps, tag + "head", x)
    x = decoder(args, x, [None] * 4, WG[4:20], taps, tag + "dec")
    return _q(args)(torch.tanh(_qg(args)(conv2d_transpose_same(x, WG[20], WG[21], 1))))


def adjuster(args, image, cond, WD, WG, WA, taps=None, tag="a_"):
    """Adjuster.call (model.py:126-136): shared encoder (D's), own dense+norm,
    shared decoder and final conv (G's), reversed encoder maps as skips."""
    enc = encoder(args, image, WD[:16], taps, tag + "enc")
    c = cond @ WA[0] + WA[1]
    c = leaky(c, args.leaky_alpha)
    c = instance_norm(c, WA[2], WA[3])
    c = c.reshape(-1, args.init_dim, args.init_dim, args.conv_filter[0])
    _tap(taps, tag + "head", c)
    x = decoder(args, c, enc[::-1], WG[4:20], taps, tag + "dec")
    return _q(args)(torch.tanh(_qg(args)(conv2d_transpose_same(x, WG[20], WG[21], 1))))


# --------------------------------------------------------------------------- #
# losses (eager_trainer.py:85-102)
# --------------------------------------------------------------------------- #
def discriminator_loss(real_true_c, real_predict_c, real_predict_pr, fake_predict_pr):
    return (bce(real_true_c, real_predict_c) * 2
            + bce(soft(torch.ones_like(real_predict_pr)), real_predict_pr)
            + bce(soft(torch.zeros_like(fake_predict_pr)), fake_predict_pr))


def generator_loss(args, cond_ori, cond_disc, pr_disc, image_ori, image_gen):
    return (bce(soft(torch.ones_like(pr_disc)), pr_disc)
            + bce(cond_ori, cond_disc)
            + args.l1_lambda * (image_ori - image_gen).abs().mean())


adjuster_loss = generator_loss  # eager_trainer.py:98-102, same form


# --------------------------------------------------------------------------- #
# TF-1.x Adam (tf.compat.v1.train.AdamOptimizer, eager_trainer.py:28-30)
# --------------------------------------------------------------------------- #
class TFAdam:
    """beta-power accumulators are per-optimiser and advance once per
    apply_gradients call, also for variables that are not in this call's list;
    epsilon (1e-8) sits outside the bias-corrected sqrt."""

    def __init__(self, lr, beta1=0.9, beta2=0.999, eps=1e-8):
        self.lr, self.b1, self.b2, self.eps = lr, beta1, beta2, eps
        self.t = 0
        self.m, self.v = {}, {}

    def apply(self, grads_and_vars):
        self.t += 1
        lr_t = self.lr * math.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
        with torch.no_grad():
            for g, p in grads_and_vars:
                key = id(p)
                if key not in self.m:
                    self.m[key] = torch.zeros_like(p)
                    self.v[key] = torch.zeros_like(p)
                m, v = self.m[key], self.v[key]
                m.mul_(self.b1).add_(g, alpha=1 - self.b1)
                v.mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
                p.sub_(lr_t * m / (v.sqrt() + self.eps))


# --------------------------------------------------------------------------- #
# trainer (eager_trainer.py:104-169, 265-298)
# --------------------------------------------------------------------------- #
PART_GROUPS = {
    "G": [range(0, 4), range(4, 8), range(8, 22)],
    "D": [range(0, 12), range(12, 16), range(16, 20)],
    "A": [range(0, 4)],   # adjuster.weights[16:20] == its own 4 tensors
}


class OracleTrainer:
    def __init__(self, args, weights, dtype=torch.float32):
        self.args = args
        self.dtype = dtype
        self.W = {k: [w.detach().clone().to(dtype).requires_grad_(True) for w in v]
                  for k, v in weights.items()}
        self.opt = {
            "G": TFAdam(args.lr, args.beta_1, args.beta_2),
            "D": TFAdam(args.lr, args.beta_1, args.beta_2),
            "A": TFAdam(args.lr),
        }

    def _train_idx(self, name, batch_no):
        """_get_train_weight (eager_trainer.py:104-113)."""
        a = self.args
        n = len(self.W[name])
        if a.use_partition and batch_no % (a.partition_interval + 1) == 0:
            groups = PART_GROUPS[name]
            return list(groups[(batch_no // (a.partition_interval + 1)) % len(groups)])
        return list(range(n))

    def train_step(self, batch_no, real_image_1, real_cond_1, real_image_2, real_cond_2,
                   noise, new_image=None, return_grads=False, taps=None):
        """_train_step (eager_trainer.py:115-169) with the random draws
        (noise, augmented new_image) injected so runs are reproducible.
        `taps`: a dict that receives every per-layer activation of the step (g_*: generator; dr_* / df_*:
        discriminator on new_image / fake_image; a_*: adjuster; da_*: discriminator on the adjusted images)
        plus `g_fake_via_D` / `g_adj_via_D`, the gradients of the adversarial + attribute terms of gen_loss /
        adj_loss with respect to the generated / adjusted images."""
        a, W = self.args, self.W
        cast = lambda t: t.to(self.dtype)
        real_image_1, real_cond_1, real_image_2, real_cond_2, noise = map(
            cast, (real_image_1, real_cond_1, real_image_2, real_cond_2, noise))
        new_image = real_image_1 if new_image is None else cast(new_image)
        if a.use_gp:
            raise NotImplementedError("GP didn't implemented on eager mode")

        fake_image = generator(a, noise, real_cond_2, W["G"], taps)
        real_pr, real_c = discriminator(a, new_image, W["D"], taps, "dr_")
        fake_pr, fake_c = discriminator(a, fake_image, W["D"], taps, "df_")
        if taps is not None:
            adv = bce(soft(torch.ones_like(fake_pr)), fake_pr) + bce(real_cond_2, fake_c)
            taps["g_fake_via_D"] = torch.autograd.grad(adv, fake_image, retain_graph=True)[0].detach()
        disc_loss = discriminator_loss(real_cond_1, real_c, real_pr, fake_pr)
        gen_loss = generator_loss(a, real_cond_2, fake_c, fake_pr, real_image_2, fake_image)

        d_idx = self._train_idx("D", batch_no)
        g_idx = self._train_idx("G", batch_no)
        d_vars = [W["D"][i] for i in d_idx]
        g_vars = [W["G"][i] for i in g_idx]
        gD = torch.autograd.grad(disc_loss, d_vars, retain_graph=True)
        if a.use_clip:
            gD = [g.clamp(-a.clip_range, a.clip_range) for g in gD]
        gG = torch.autograd.grad(gen_loss, g_vars, retain_graph=False)

        adj_image, adj_loss, gA, a_idx = None, None, None, None
        if a.train_adj and batch_no > 10:
            a_idx = self._train_idx("A", batch_no)
            a_vars = [W["A"][i] for i in a_idx]
            fake_const = fake_image.detach()
            adj_input_cond = (torch.cat([real_cond_2, real_cond_1], 0) + 1) * 0.5
            adj_target_cond = torch.cat([real_cond_2, real_cond_1], 0)
            adj_input_image = torch.cat([real_image_1, fake_const], 0)
            adj_target_image = torch.cat([real_image_2, real_image_1], 0)
            adj_image = adjuster(a, adj_input_image, adj_input_cond, W["D"], W["G"], W["A"], taps)
            adj_pr, adj_c = discriminator(a, adj_image, W["D"], taps, "da_")
            if taps is not None:
                adv = bce(soft(torch.ones_like(adj_pr)), adj_pr) + bce(adj_target_cond, adj_c)
                taps["g_adj_via_D"] = torch.autograd.grad(adv, adj_image, retain_graph=True)[0].detach()
            adj_loss = adjuster_loss(a, adj_target_cond, adj_c, adj_pr, adj_target_image, adj_image)
            gA = torch.autograd.grad(adj_loss, a_vars)
            self.opt["A"].apply(zip(gA, a_vars))
        self.opt["D"].apply(zip(gD, d_vars))
        self.opt["G"].apply(zip(gG, g_vars))

        out = dict(fake_image=fake_image.detach(), adj_image=None if adj_image is None else adj_image.detach(),
                   gen_loss=gen_loss.detach(), disc_loss=disc_loss.detach(),
                   adj_loss=None if adj_loss is None else adj_loss.detach())
        if return_grads:
            out["grads"] = dict(D=dict(zip(d_idx, gD)), G=dict(zip(g_idx, gG)),
                                A=None if gA is None else dict(zip(a_idx, gA)))
        return out

    @torch.no_grad()
    def forward_losses(self, batch_no, real_image_1, real_cond_1, real_image_2, real_cond_2, noise,
                       new_image=None, weights=None, args=None):
        """The three losses of `_train_step` (eager_trainer.py:134-140, 152-161) WITHOUT the update, plus for each
        the sum of the absolute values of its terms (`*_scale`: soft labels lie outside [0,1], so a loss is a
        difference of O(1) terms that may pass through zero - its own value is no scale for a relative error).
        `weights` / `args` override self.W / self.args (the parity tests evaluate a bf16-storage-emulating copy
        on bf16-representable weights)."""
        a = args or self.args
        W = weights or self.W
        cast = lambda t: t.to(self.dtype)
        i1, c1, i2, c2, noise = map(cast, (real_image_1, real_cond_1, real_image_2, real_cond_2, noise))
        new_image = i1 if new_image is None else cast(new_image)
        one = lambda p: soft(torch.ones_like(p))
        zero = lambda p: soft(torch.zeros_like(p))
        fake = generator(a, noise, c2, W["G"])
        real_pr, real_c = discriminator(a, new_image, W["D"])
        fake_pr, fake_c = discriminator(a, fake, W["D"])
        out = {}
        t = [2 * bce(c1, real_c), bce(one(real_pr), real_pr), bce(zero(fake_pr), fake_pr)]
        out["disc_loss"], out["disc_scale"] = float(sum(t)), float(sum(x.abs() for x in t))
        t = [bce(one(fake_pr), fake_pr), bce(c2, fake_c), a.l1_lambda * (i2 - fake).abs().mean()]
        out["gen_loss"], out["gen_scale"] = float(sum(t)), float(sum(x.abs() for x in t))
        out["adj_loss"] = out["adj_scale"] = None
        if a.train_adj and batch_no > 10:
            cin = (torch.cat([c2, c1], 0) + 1) * 0.5
            ct = torch.cat([c2, c1], 0)
            adj = adjuster(a, torch.cat([i1, fake], 0), cin, W["D"], W["G"], W["A"])
            apr, ac = discriminator(a, adj, W["D"])
            t = [bce(one(apr), apr), bce(ct, ac), a.l1_lambda * (torch.cat([i2, i1], 0) - adj).abs().mean()]
            out["adj_loss"], out["adj_scale"] = float(sum(t)), float(sum(x.abs() for x in t))
        return out

    @torch.no_grad()
    def predict(self, noise, cond, image):
        """predict (eager_trainer.py:265-298) without the file I/O."""
        a, W = self.args, self.W
        noise, cond, image = (t.to(self.dtype) for t in (noise, cond, image))
        gen_image = generator(a, noise, cond, W["G"])
        save = {}
        save["real_pr"], save["real_c"] = discriminator(a, image, W["D"])
        save["fake_pr"], save["fake_c"] = discriminator(a, gen_image, W["D"])
        mse = lambda t, p: ((t - p) ** 2).mean(dim=-1).mean(dim=0)
        save["real_pr_mse"] = float(mse(soft(1.0), save["real_pr"]))
        save["real_c_mse"] = float(mse(cond, save["real_c"]))
        save["fake_pr_mse"] = float(mse(soft(0.0), save["fake_pr"]))
        save["fake_c_mse"] = float(mse(cond, save["fake_c"]))
        adj_real = adj_fake = None
        if a.train_adj:
            adj_real = adjuster(a, image, cond, W["D"], W["G"], W["A"])
            adj_fake = adjuster(a, gen_image, cond, W["D"], W["G"], W["A"])
        return gen_image, save, adj_real, adj_fake


# --------------------------------------------------------------------------- #
# synthetic inputs (SURVEY 8 d)
# --------------------------------------------------------------------------- #
def synthetic_batch(args, batch, seed=0, dtype=torch.float32):
    """images U(-1,1) [B,H,W,C] x2, labels soft(+-1) in {-0.94, 0.98} x2,
    noise N(0,1) [B, noise_dim]."""
    g = torch.Generator().manual_seed(seed)
    H = args.init_dim * 16
    img = lambda: (torch.rand(batch, H, H, args.image_channel, generator=g) * 2 - 1).to(dtype)
    lab = lambda: soft((torch.rand(batch, args.cond_dim, generator=g) < 0.5).to(dtype) * 2 - 1)
    i1, c1, i2, c2 = img(), lab(), img(), lab()
    noise = torch.randn(batch, args.noise_dim, generator=g).to(dtype)
    return i1, c1, i2, c2, noise


# --------------------------------------------------------------------------- #
# input augmentation (eager_trainer.py:127-131), random draws injected
# --------------------------------------------------------------------------- #
def adjust_hue(x, delta):
    """tf.image.adjust_hue on [..., 3] RGB: rotate the hue by `delta` turns keeping each pixel's max and min
    (TF's fused AdjustHue kernel works on (hue, v_min, v_max), which is defined for any value range - the
    reference feeds images in [-1, 1]).  Per-pixel Python-free restatement of the hexcone model."""
    r, g, b = x[..., 0], x[..., 1], x[..., 2]
    vmax = torch.maximum(r, torch.maximum(g, b))
    vmin = torch.minimum(r, torch.minimum(g, b))
    rng = vmax - vmin
    safe = torch.where(rng > 0, rng, torch.ones_like(rng))
    h = torch.where(r == vmax, (g - b) / safe,
                    torch.where(g == vmax, 2.0 + (b - r) / safe, 4.0 + (r - g) / safe))
    h = torch.remainder(h + 6.0 * delta, 6.0)
    xm = vmin + rng * (1.0 - torch.abs(torch.remainder(h, 2.0) - 1.0))
    sec = torch.clamp(torch.floor(h), 0, 5).long()
    table = [(vmax, xm, vmin), (xm, vmax, vmin), (vmin, vmax, xm), (vmin, xm, vmax), (xm, vmin, vmax), (vmax, vmin, xm)]
    out = torch.stack([torch.stack([table[k][c] for k in range(6)], -1).gather(-1, sec[..., None])[..., 0]
                       for c in range(3)], -1)
    return torch.where((rng > 0)[..., None], out, x)


def augment(x, flips, db, fc, dh, noise):
    """new_image of eager_trainer.py:127-131 with the draws given: flips [N] (1 = flip left-right, per image:
    tf.image.random_flip_left_right), db brightness delta (random_brightness(0.02): one scalar per batch), fc
    contrast factor (random_contrast(0.75, 1.003): (x - mean_HW) * fc + mean_HW per image and channel), dh hue
    delta in turns (random_hue(x, 0.03, -0.03): max_delta 0.03 - the third argument is the seed), noise
    [N,H,W,3] = 0.1 * N(0, 0.2) already scaled."""
    x = torch.where(flips.reshape(-1, 1, 1, 1) > 0.5, x.flip(2), x)
    x = x + db
    mean = x.mean(dim=(1, 2), keepdim=True)
    x = (x - mean) * fc + mean
    x = adjust_hue(x, dh)
    return x + noise
