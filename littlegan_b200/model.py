"""Model builders with the reference's surface (`model.py:1-136`): Encoder, Decoder,
Discriminator, Generator, Adjuster - same constructor signatures, call conventions, attribute
names and `.weights` ordering (eager_trainer.py:48-63 indexes into it).

Differences a caller can see: tensors are torch CUDA tensors (NHWC, `args.dtype` storage) instead
of EagerTensors, and variables are created at construction (all shapes follow from `args`)
instead of at first call.  The arithmetic runs in the CUDA kernels of `csrc/`; there is no CPU
path.
"""
import math

import numpy as np
import torch

from . import engine as E
from . import kernels as K
from .instance import InstanceNormalization, _device

_INIT_SEED = [0]


def set_init_seed(seed):
    """Seed of the Glorot-uniform initialiser stream (Keras default init, unseeded there)."""
    _INIT_SEED[0] = int(seed)
    _GEN.clear()               # restart the stream: equal seeds give equal weights


_GEN = {}


def _glorot(shape, fan_in, fan_out):
    gen = _GEN.get("g")
    if gen is None or _GEN.get("seed") != _INIT_SEED[0]:
        gen = torch.Generator().manual_seed(_INIT_SEED[0])
        _GEN["g"], _GEN["seed"] = gen, _INIT_SEED[0]
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    w = (torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1) * lim
    return w.to(torch.float32).to(_device())


def _check_geometry(kernel_size, strides):
    """Every CUDA path of the conv family is written for the reference's 5x5 kernels (sample.config.json:14;
    the C-ABI carries no kernel-size argument) at stride 1 or 2: refuse anything else up front instead of
    letting a kernel read 25 taps out of a smaller buffer."""
    if int(kernel_size) != 5:
        raise ValueError("littlegan_b200 implements kernel_size 5 only (got %r): the CUDA kernels and the C-ABI "
                         "are specialised for the reference's 5x5 convolutions" % (kernel_size,))
    if int(strides) not in (1, 2):
        raise ValueError("littlegan_b200 implements strides 1 and 2 only (got %r)" % (strides,))


class Conv2D:
    """tf.layers.Conv2D(filters, k, strides, 'same'): kernel [k,k,in,out], bias [out]."""

    def __init__(self, in_channels, filters, kernel_size, strides):
        _check_geometry(kernel_size, strides)
        k = kernel_size
        self.filters, self.kernel_size, self.strides = filters, k, strides
        self.kernel = _glorot((k, k, in_channels, filters), k * k * in_channels, k * k * filters)
        self.bias = torch.zeros(filters, dtype=torch.float32, device=_device())
        self.wpack = None
        self.wpack_rows = None      # weight operand of the row-streaming kernel (engine.refresh_packs)
        self.wpack_drows = None     # ... of the row-streaming dgrad kernel

    @property
    def weights(self):
        return [self.kernel, self.bias]


class Conv2DTranspose:
    """tf.layers.Conv2DTranspose(filters, k, strides, 'same'): kernel [k,k,out,in], bias [out]."""

    def __init__(self, in_channels, filters, kernel_size, strides, activation=None):
        _check_geometry(kernel_size, strides)
        k = kernel_size
        self.filters, self.kernel_size, self.strides, self.activation = filters, k, strides, activation
        self.kernel = _glorot((k, k, filters, in_channels), k * k * filters, k * k * in_channels)
        self.bias = torch.zeros(filters, dtype=torch.float32, device=_device())
        self.wpack = None
        self.wpack_rows = None      # weight operand of the row-streaming kernel (engine.refresh_packs)
        self.wpack_drows = None     # ... of the row-streaming dgrad kernel

    @property
    def weights(self):
        return [self.kernel, self.bias]


class Dense:
    """tf.layers.Dense(units, activation): kernel [in,out], bias [out]."""

    def __init__(self, in_features, units, activation=None):
        self.units, self.activation = units, activation
        self.kernel = _glorot((in_features, units), in_features, units)
        self.bias = torch.zeros(units, dtype=torch.float32, device=_device())

    @property
    def weights(self):
        return [self.kernel, self.bias]


def _as_input(rt, x, dtype=None):
    """numpy / CPU / CUDA input -> contiguous CUDA tensor of the requested dtype."""
    if not torch.cuda.is_available():
        raise K._lib.LittleGANError("littlegan_b200 has no CPU path: a CUDA device is required")
    from .utils import upload
    return upload(x, rt.device).to(dtype or rt.act_dtype).contiguous()


class _Model:
    def __init__(self, args):
        self.args = args
        self._rt = None

    @property
    def rt(self):
        if self._rt is None:
            self._rt = E.Runtime(self.args)
        return self._rt

    def conv_layers(self):
        return []

    def __call__(self, inputs, training=None, mask=None):
        E.refresh_packs(self.rt, self.conv_layers())
        return self.call(inputs, training, mask)


class Encoder(_Model):
    def __init__(self, args):
        super().__init__(args)
        k = args.kernel_size
        cin = args.image_channel
        for i in range(1, 5):
            setattr(self, "conv" + str(i), Conv2D(cin, args.conv_filter[4 - i], k, 2))
            setattr(self, "norm" + str(i), InstanceNormalization())
            cin = args.conv_filter[4 - i]

    @property
    def convs(self):
        return [getattr(self, "conv" + str(i)) for i in range(1, 5)]

    @property
    def norms(self):
        return [getattr(self, "norm" + str(i)) for i in range(1, 5)]

    def conv_layers(self):
        return self.convs

    @property
    def weights(self):
        w = []
        for c, n in zip(self.convs, self.norms):
            w += c.weights + n.weights
        return w

    def call(self, inputs, training=None, mask=None):
        outs, _ = E.encoder_forward(self.rt, self, _as_input(self.rt, inputs))
        return outs


class Decoder(_Model):
    def __init__(self, args):
        super().__init__(args)
        k = args.kernel_size
        cin = args.conv_filter[0]
        for i in range(1, 5):
            setattr(self, "conv" + str(i), Conv2DTranspose(cin, args.conv_filter[i], k, 2))
            setattr(self, "norm" + str(i), InstanceNormalization())
            cin = args.conv_filter[i]

    convs = Encoder.convs
    norms = Encoder.norms
    weights = Encoder.weights

    def conv_layers(self):
        return self.convs

    def call(self, inputs, training=None, mask=None):
        x, add = inputs
        x = _as_input(self.rt, x)
        add = [None if a is None else _as_input(self.rt, a) for a in add]
        if add[0] is not None:
            x = x + add[0]                      # the fused paths fold this into the producer of x
        out, _ = E.decoder_forward(self.rt, self, x, add[1:4])
        return out


class Discriminator(_Model):
    def __init__(self, args, encoder):
        super().__init__(args)
        self.encoder = encoder
        flat = args.init_dim ** 2 * args.conv_filter[0]
        self.dense_pr = Dense(flat, 1, "sigmoid")
        self.dense_cond = Dense(flat, args.cond_dim, "sigmoid")

    def conv_layers(self):
        return self.encoder.convs

    @property
    def weights(self):
        return self.encoder.weights + self.dense_pr.weights + self.dense_cond.weights

    def call(self, inputs, training=None, mask=None):
        outs, _ = E.encoder_forward(self.rt, self.encoder, _as_input(self.rt, inputs))
        return E.disc_heads_forward(self.rt, self, outs[-1])


class Generator(_Model):
    def __init__(self, args, decoder):
        super().__init__(args)
        flat = args.init_dim ** 2 * args.conv_filter[0]
        self.dense = Dense(args.noise_dim + args.cond_dim, flat)
        self.norm = InstanceNormalization()
        self.decoder = decoder
        self.conv = Conv2DTranspose(args.conv_filter[4], args.image_channel, args.kernel_size, 1, "tanh")

    def conv_layers(self):
        return self.decoder.convs + [self.conv]

    @property
    def weights(self):
        return self.dense.weights + self.norm.weights + self.decoder.weights + self.conv.weights

    def forward_ctx(self, noise, cond, out=None):
        rt, a = self.rt, self.args
        xin = torch.cat([noise, cond], dim=-1).contiguous()
        x0, hctx = E.head_forward(rt, self.dense, self.norm, xin,
                                  (xin.shape[0], a.init_dim, a.init_dim, a.conv_filter[0]))
        x4, dctx = E.decoder_forward(rt, self.decoder, x0)
        img = E.final_conv_forward(rt, self.conv, x4, out)
        return img, (hctx, dctx, x4)

    def call(self, inputs, training=None, mask=None):
        noise, cond = inputs
        noise = _as_input(self.rt, noise, torch.float32)
        cond = _as_input(self.rt, cond, torch.float32)
        return self.forward_ctx(noise, cond)[0]


class Adjuster(_Model):
    def __init__(self, args, discriminator, generator):
        super().__init__(args)
        self.encoder = discriminator.encoder
        flat = args.init_dim ** 2 * args.conv_filter[0]
        self.dense = Dense(args.cond_dim, flat)
        self.norm = InstanceNormalization()
        self.decoder = generator.decoder
        self.conv = generator.conv

    def conv_layers(self):
        return self.encoder.convs + self.decoder.convs + [self.conv]

    @property
    def weights(self):
        return (self.encoder.weights + self.dense.weights + self.norm.weights + self.decoder.weights
                + self.conv.weights)

    def forward_ctx(self, image, cond, out=None, enc=None):
        """`enc`: the four maps of self.encoder(image) if the caller already has them (the train step runs the
        shared encoder once for the discriminator and the adjuster)."""
        rt, a = self.rt, self.args
        if enc is None:
            enc, _ = E.encoder_forward(rt, self.encoder, image)
        c0, hctx = E.head_forward(rt, self.dense, self.norm, cond,
                                  (cond.shape[0], a.init_dim, a.init_dim, a.conv_filter[0]), skip=enc[3])
        x4, dctx = E.decoder_forward(rt, self.decoder, c0, (enc[2], enc[1], enc[0]))
        img = E.final_conv_forward(rt, self.conv, x4, out)
        return img, (hctx, dctx, x4)

    def call(self, inputs, training=None, mask=None):
        image, cond = inputs
        return self.forward_ctx(_as_input(self.rt, image), _as_input(self.rt, cond, torch.float32))[0]
