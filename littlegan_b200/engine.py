"""Forward / hand-written backward passes of the LittleGAN networks on the C-ABI kernels.

The reference differentiates with tf.GradientTape (eager_trainer.py:133-163); here every pass is
an explicit sequence of kernel launches that computes exactly the gradients the train step needs
(13 encoder + 7 decoder conv passes per step).  A pass returns its outputs plus a
`ctx` list of (layer input, pre-norm conv output z, per-sample statistics) for the backward.

Every tensor is NHWC-contiguous; activations are `rt.act_dtype` (bf16 or fp32), the dense heads,
statistics, losses, weight gradients and optimiser state are fp32/fp64.
"""
import torch

from . import kernels as K


class Runtime:
    """Per-process execution settings derived from the config (`dtype`)."""

    def __init__(self, args):
        mode = getattr(args, "dtype", "bf16")
        if mode not in ("bf16", "fp32"):
            raise ValueError("dtype must be 'bf16' or 'fp32'")
        self.act_dtype = torch.bfloat16 if mode == "bf16" else torch.float32
        self.alpha = float(args.leaky_alpha)
        self.device = torch.device("cuda")
        self.want_tc = mode == "bf16" and getattr(args, "tensor_cores", True)
        self._tc_cache = {}
        self._heads_ws = None
        # weight-gradient launches run on a side stream: nothing in the backward chain consumes them, so they
        # overlap with the bandwidth-bound norm kernels that follow on the main stream (also inside a captured
        # CUDA graph, where the fork / join become parallel branches)
        # per-step arena for the small zero-initialised buffers (statistics, reductions, logit gradients): one
        # memset per step instead of one fill launch per buffer
        self._zarena, self._zoff, self._zactive = None, 0, False
        self.overlap = bool(getattr(args, "overlap_wgrad", True))
        # one side stream per launching stream: the train step runs its independent backward chains on
        # separate streams (EagerTrainer._step_body), each with its own weight-gradient branch
        self._sides = {}

    def use_tc(self, op, N, Hb, Wb, A, B, s):
        if not self.want_tc:
            return False
        key = (op, N, Hb, Wb, A, B, s)
        r = self._tc_cache.get(key)
        if r is None:
            r = K.tc_available() and K.tc_supported(op, N, Hb, Wb, A, B, s)
            self._tc_cache[key] = r
        return r

    def fuse_norm_bwd(self, op, N, Hb, Wb, A, B, s):
        """May the backward conv launch (op, geometry) carry the fused InstanceNorm-backward epilogue?"""
        if not self.use_tc(op, N, Hb, Wb, A, B, s):
            return False
        key = ("nb", op, N, Hb, Wb, A, B, s)
        r = self._tc_cache.get(key)
        if r is None:
            r = self._tc_cache[key] = K.norm_bwd_supported(op, N, Hb, Wb, A, B, s)
        return r

    def rows_ok(self, N, Hb, Wb, A, B, s):
        """Does the row-streaming fprop kernel (csrc/tc_rowconv.cu) take this geometry?  Returns the channel
        count its input must be stored with (8 for RGB maps, A otherwise) or 0."""
        if not self.want_tc:
            return 0
        key = ("rows", N, Hb, Wb, A, B, s)
        r = self._tc_cache.get(key)
        if r is None:
            A_big = 8 if A <= 8 else A
            r = A_big if (K.tc_available() and K.fprop_rows_supported(N, Hb, Wb, A_big, A, B, s)) else 0
            self._tc_cache[key] = r
        return r

    def drows_ok(self, N, Hb, Wb, A, B, s):
        """Does the row-streaming dgrad kernel (csrc/tc_rowdgrad.cu) take this geometry?"""
        if not self.want_tc:
            return False
        key = ("drows", N, Hb, Wb, A, B, s)
        r = self._tc_cache.get(key)
        if r is None:
            r = self._tc_cache[key] = K.tc_available() and K.dgrad_rows_supported(N, Hb, Wb, A, B, s)
        return r

    def on_side(self, fn, *keep):
        """Run the launches of `fn` on the side stream, ordered after everything issued so far on the current
        stream.  `keep`: tensors the launches read - held until join_side() so the caching allocator cannot
        hand their memory to later main-stream work while the side stream is still reading."""
        if not self.overlap:
            fn()
            return
        main = torch.cuda.current_stream()
        ent = self._sides.get(main.cuda_stream)
        if ent is None:
            ent = self._sides[main.cuda_stream] = [torch.cuda.Stream(), []]
        ent[0].wait_stream(main)
        with torch.cuda.stream(ent[0]):
            fn()
        ent[1].append(keep)

    def join_side(self):
        """Make the current stream wait for its side-stream launches (before their results are consumed)."""
        main = torch.cuda.current_stream()
        ent = self._sides.get(main.cuda_stream)
        if ent is not None and ent[1]:
            main.wait_stream(ent[0])
            ent[1] = []

    def heads_workspace(self, N):
        """Scratch of the fused discriminator-heads forward (self-cleaning, allocated once per size and per
        launching stream: two chains of the train step may run the heads at the same time)."""
        if self._heads_ws is None:
            self._heads_ws = {}                 # never freed: captured CUDA graphs hold the pointers
        key = (N, torch.cuda.current_stream().cuda_stream)
        ws = self._heads_ws.get(key)
        if ws is None:
            ws = self._heads_ws[key] = K.dense_heads_workspace(N, self.device)
        return ws

    def empty(self, *shape, dtype=None):
        return torch.empty(*shape, dtype=dtype or self.act_dtype, device=self.device)

    ZARENA_BYTES = 1 << 20

    def begin_step(self):
        """Start of a train step: zero the arena (one memset) and hand out zeros() from it until end_step()."""
        if self._zarena is None:
            self._zarena = torch.empty(self.ZARENA_BYTES, dtype=torch.uint8, device=self.device)
        self._zarena.zero_()
        self._zoff, self._zactive = 0, True

    def end_step(self):
        self._zactive = False

    def zeros(self, *shape, dtype=torch.float64):
        if self._zactive:
            n = 1
            for d in shape:
                n *= int(d)
            nbytes = n * torch.empty(0, dtype=dtype).element_size()
            if self._zoff + nbytes <= self.ZARENA_BYTES:
                t = self._zarena[self._zoff:self._zoff + nbytes].view(dtype).view(*shape)
                self._zoff += (nbytes + 255) & ~255
                return t
        return torch.zeros(*shape, dtype=dtype, device=self.device)


def _padded_for_tc(rt, x, op, B, s):
    """3-channel maps cannot be fetched by TMA (6-byte pixels): give the tcgen05 path a copy padded
    to 16 channels (one cheap pass over a tensor that is 1/20th of the layer's output)."""
    N, Hb, Wb, A = x.shape
    if A >= 16 or not rt.want_tc or rt.use_tc(op, N, Hb, Wb, A, B, s):   # the RGB kernels take the image as is
        return None
    if not rt.use_tc(op, N, Hb, Wb, 16, B, s):
        return None
    return K.pad_channels(x, rt.empty(N, Hb, Wb, 16))


def _fprop(rt, conv, x, bias, out, stats, s, nb=None, xrows=None):
    """Conv2D forward / Conv2DTranspose input-gradient of `conv` on the best kernel: the row-streaming kernel
    for the 128-pixel-wide maps with <= 32 channels (RGB maps through an 8-channel padded copy, `xrows` if the
    caller already has one), else the generic tensor-core / SIMT path.  Returns the padded copy it used."""
    N, Hb, Wb, A = x.shape
    B = out.shape[3]
    A_big = rt.rows_ok(N, Hb, Wb, A, B, s) if conv.wpack_rows is not None else 0
    if A_big:
        if A_big != A and xrows is None:
            xrows = K.pad_channels(x, rt.empty(N, Hb, Wb, A_big))
        K.conv2d_fprop_rows(xrows if A_big != A else x, conv.wpack_rows, bias, out, stats, s, A, norm_bwd=nb)
        return xrows
    K.conv2d_fprop(x, conv.kernel, bias, out, stats, s, conv.wpack, rt.use_tc(K.OP_FPROP, N, Hb, Wb, A, B, s),
                   norm_bwd=nb)
    return None


def _fprop_fuses_norm_bwd(rt, conv, N, Hb, Wb, A, B, s):
    if conv.wpack_rows is not None and rt.rows_ok(N, Hb, Wb, A, B, s):
        return True
    return rt.fuse_norm_bwd(K.OP_FPROP, N, Hb, Wb, A, B, s)


def _grad(p):
    """Gradient slot of a parameter (a view into the trainer's flat gradient arena)."""
    g = getattr(p, "lg_grad", None)
    if g is None:
        raise RuntimeError("parameter has no gradient slot; construct an EagerTrainer first")
    return g


def refresh_packs(rt, conv_layers):
    """Re-derive the bf16 operand copies of the conv kernels (after any weight change)."""
    if not rt.want_tc or not K.tc_available():
        return
    for layer in conv_layers:
        if layer.wpack is None:
            A, B = layer.kernel.shape[2], layer.kernel.shape[3]
            layer.wpack = torch.empty(K.pack_conv_weights_bytes(A, B), dtype=torch.uint8, device=rt.device)
    K.pack_conv_weights_multi([layer.kernel for layer in conv_layers], [layer.wpack for layer in conv_layers])
    for layer in conv_layers:
        A, B = layer.kernel.shape[2], layer.kernel.shape[3]
        A_big = 8 if A <= 8 else A
        if K.fprop_rows_supported(1, 128, 128, A_big, A, B, layer.strides):
            layer.wpack_rows = K.pack_rowconv_weights(layer.kernel, A_big, layer.strides, layer.wpack_rows)
        if K.dgrad_rows_supported(1, 128, 128, A, B, layer.strides):
            layer.wpack_drows = K.pack_rowdgrad_weights(layer.kernel, layer.wpack_drows)


# --------------------------------------------------------------------------------------------
# encoder (model.py:18-27): 4 x [Conv2D s2 -> IN -> LeakyReLU]; dropout is the identity
# --------------------------------------------------------------------------------------------
def encoder_forward(rt, enc, x):
    N = x.shape[0]
    stats = rt.zeros(4, N, 2)
    outs, ctx = [], []
    for i in range(4):
        conv, norm = enc.convs[i], enc.norms[i]
        _, Hb, Wb, A = x.shape
        B = conv.filters
        z = rt.empty(N, Hb // 2, Wb // 2, B)
        xpad = _padded_for_tc(rt, x, K.OP_FPROP, B, 2)
        if xpad is not None:
            K.conv2d_fprop(xpad, conv.kernel, conv.bias, z, stats[i], 2, conv.wpack, True)
        else:
            _fprop(rt, conv, x, conv.bias, z, stats[i], 2)
        a = rt.empty(N, Hb // 2, Wb // 2, B)
        K.instnorm_act_fwd(z, stats[i], norm.gamma, norm.beta, None, a, norm.epsilon, 1.0, rt.alpha)
        ctx.append((x, z, stats[i], xpad))
        outs.append(a)
        x = a
    return outs, ctx


def _norm_act_fwd(z, stats, norm, skip, out, a1, a2):
    """IN + activation (+ additive skip) of a batch.  `skip` may be a tuple of tensors that together cover the
    batch along N (the train step keeps the encoder maps of [real_image_1 | ... | fake] in one buffer where the
    adjuster's batch [real_image_1 ; fake] is not contiguous): one launch per piece - the norm is per sample."""
    if not isinstance(skip, (tuple, list)):
        K.instnorm_act_fwd(z, stats, norm.gamma, norm.beta, skip, out, norm.epsilon, a1, a2)
        return
    lo = 0
    for sk in skip:
        hi = lo + sk.shape[0]
        K.instnorm_act_fwd(z[lo:hi], stats[lo:hi], norm.gamma, norm.beta, sk, out[lo:hi], norm.epsilon, a1, a2)
        lo = hi
    assert lo == z.shape[0], "skip pieces must cover the batch"


class EncoderPass:
    """encoder_forward over one batch buffer, run in slices: the norm is per sample, so the slices of a batch are
    independent passes that may be issued at different times and on different streams (the train step starts the
    slice that holds the real images while the generator is still producing the fake ones).  All outputs are
    allocated up front on the constructing stream; run(lo, hi) fills rows [lo, hi) on the CURRENT stream;
    result() is encoder_forward's (outs, ctx) over the whole batch."""

    def __init__(self, rt, enc, x):
        self.rt, self.enc, self.x = rt, enc, x
        N, H, W, _ = x.shape
        self.stats = rt.zeros(4, N, 2)
        self.z, self.a = [], []
        for i in range(4):
            H, W = H // 2, W // 2
            self.z.append(rt.empty(N, H, W, enc.convs[i].filters))
            self.a.append(rt.empty(N, H, W, enc.convs[i].filters))

    @staticmethod
    def sliceable(rt, enc, x):
        """False when layer 1 needs a channel-padded copy of the whole image batch (kept in ctx for its wgrad)."""
        N, Hb, Wb, A = x.shape
        return A >= 16 or not rt.want_tc or rt.use_tc(K.OP_FPROP, N, Hb, Wb, A, enc.convs[0].filters, 2)

    def run(self, lo, hi):
        rt, enc = self.rt, self.enc
        x = self.x[lo:hi]
        for i in range(4):
            conv, norm = enc.convs[i], enc.norms[i]
            z, a, st = self.z[i][lo:hi], self.a[i][lo:hi], self.stats[i][lo:hi]
            _fprop(rt, conv, x, conv.bias, z, st, 2)
            K.instnorm_act_fwd(z, st, norm.gamma, norm.beta, None, a, norm.epsilon, 1.0, rt.alpha)
            x = a

    def result(self):
        xs = [self.x] + self.a[:3]
        return self.a, [(xs[i], self.z[i], self.stats[i], None) for i in range(4)]


def _norm_act_bwd(rt, g, z, stats, norm, red, conv, wgrad, dy_ready):
    """IN + LeakyReLU backward of one conv layer -> dz; with `wgrad` also d(gamma), d(beta), d(bias)."""
    dz = torch.empty_like(z)
    fuse_db = wgrad and K.bias_grad_fusable(z)
    K.instnorm_act_bwd(g.view_as(z), z, stats, norm.gamma, norm.beta, red, dz,
                       _grad(norm.gamma) if wgrad else None, _grad(norm.beta) if wgrad else None,
                       norm.epsilon, 1.0, rt.alpha, dy_ready=dy_ready,
                       dbias=_grad(conv.bias) if fuse_db else None)
    if wgrad and not fuse_db:
        K.bias_grad(dz, _grad(conv.bias))
    return dz


def encoder_backward(rt, enc, ctx, g, wgrad, input_grad):
    """g: gradient w.r.t. the last encoder output.  wgrad: accumulate kernel/bias/gamma/beta
    gradients.  Returns the gradient w.r.t. the encoder input if `input_grad`.
    The dgrad launch of layer i produces the gradient w.r.t. layer i-1's activation: where the
    tensor-core kernel supports it, that launch also does pass 1 of layer i-1's norm backward
    (csrc/norm_bwd.cuh) and hands over dy instead of g."""
    N = g.shape[0]
    red = rt.zeros(4, N, 2)
    dy_ready = False
    for i in (3, 2, 1, 0):
        conv, norm = enc.convs[i], enc.norms[i]
        x, z, stats, xpad = ctx[i]
        dz = _norm_act_bwd(rt, g, z, stats, norm, red[i], conv, wgrad, dy_ready)
        _, Hb, Wb, A = x.shape
        B = conv.filters
        if wgrad:
            if xpad is not None and rt.use_tc(K.OP_WGRAD, N, Hb, Wb, 16, B, 2):
                rt.on_side(lambda: K.conv2d_wgrad_padded(xpad, dz, _grad(conv.kernel), 2), xpad, dz)
            else:
                tcw = rt.use_tc(K.OP_WGRAD, N, Hb, Wb, A, B, 2)
                rt.on_side(lambda: K.conv2d_wgrad(x, dz, _grad(conv.kernel), 2, tcw), x, dz)
        dy_ready = False
        if i > 0 or input_grad:
            g = torch.empty_like(x)
            nb = None
            if i > 0 and rt.fuse_norm_bwd(K.OP_DGRAD, N, Hb, Wb, A, B, 2):
                _, zp, sp, _ = ctx[i - 1]
                nb = K.norm_bwd_desc(zp, sp, enc.norms[i - 1].gamma, enc.norms[i - 1].beta, red[i - 1],
                                     enc.norms[i - 1].epsilon, rt.alpha)
                dy_ready = True
            K.conv2d_dgrad(dz, conv.kernel, None, g, None, 2, K.ACT_NONE, conv.wpack,
                           rt.use_tc(K.OP_DGRAD, N, Hb, Wb, A, B, 2), norm_bwd=nb)
        else:
            g = None
    rt.join_side()
    return g


# --------------------------------------------------------------------------------------------
# decoder (model.py:43-51): 4 x [(+skip) -> Conv2DTranspose s2 -> IN -> LeakyReLU]
# `skips_after[i]` is added to the OUTPUT of layer i (i.e. it is the reference's add[i+1]);
# add[0] is folded into whatever produced the decoder input.
# --------------------------------------------------------------------------------------------
def decoder_forward(rt, dec, x, skips_after=(None, None, None)):
    N = x.shape[0]
    stats = rt.zeros(4, N, 2)
    ctx = []
    for i in range(4):
        conv, norm = dec.convs[i], dec.norms[i]
        _, Hs, Ws, B = x.shape
        A = conv.filters
        z = rt.empty(N, 2 * Hs, 2 * Ws, A)
        if conv.wpack_drows is not None and rt.drows_ok(N, 2 * Hs, 2 * Ws, A, B, 2):
            K.conv2d_dgrad_rows(x, conv.wpack_drows, conv.bias, z, stats[i], 2)
        else:
            tc = rt.use_tc(K.OP_DGRAD, N, 2 * Hs, 2 * Ws, A, B, 2)
            K.conv2d_dgrad(x, conv.kernel, conv.bias, z, stats[i], 2, K.ACT_NONE, conv.wpack, tc)
        a = torch.empty_like(z)
        skip = skips_after[i] if i < 3 else None
        _norm_act_fwd(z, stats[i], norm, skip, a, 1.0, rt.alpha)
        ctx.append((x, z, stats[i]))
        x = a
    return x, ctx


def decoder_backward(rt, dec, ctx, g, wgrad, red=None, dy_ready=False):
    """Returns the gradient w.r.t. the decoder input (always needed: the heads sit below).
    `red` / `dy_ready`: the producer of `g` (final_conv_backward) already ran pass 1 of the last layer's
    norm backward into red[3]."""
    N = g.shape[0]
    if red is None:
        red = rt.zeros(4, N, 2)
        dy_ready = False
    for i in (3, 2, 1, 0):
        conv, norm = dec.convs[i], dec.norms[i]
        x, z, stats = ctx[i]
        dz = _norm_act_bwd(rt, g, z, stats, norm, red[i], conv, wgrad, dy_ready)
        _, Hb, Wb, A = z.shape
        B = x.shape[3]
        if wgrad:
            tcw = rt.use_tc(K.OP_WGRAD, N, Hb, Wb, A, B, 2)
            rt.on_side(lambda: K.conv2d_wgrad(dz, x, _grad(conv.kernel), 2, tcw), dz, x)
        g = torch.empty_like(x)
        nb = None
        dy_ready = False
        if i > 0 and _fprop_fuses_norm_bwd(rt, conv, N, Hb, Wb, A, B, 2):
            _, zp, sp = ctx[i - 1]
            nb = K.norm_bwd_desc(zp, sp, dec.norms[i - 1].gamma, dec.norms[i - 1].beta, red[i - 1],
                                 dec.norms[i - 1].epsilon, rt.alpha)
            dy_ready = True
        _fprop(rt, conv, dz, None, g, None, 2, nb)
    rt.join_side()
    return g


# --------------------------------------------------------------------------------------------
# final Conv2DTranspose(image_channel, k, stride 1, tanh) (model.py:86,104)
# --------------------------------------------------------------------------------------------
def final_conv_forward(rt, conv, x, out=None):
    N, H, W, B = x.shape
    A = conv.filters
    if out is None:
        out = rt.empty(N, H, W, A)
    K.conv2d_dgrad(x, conv.kernel, conv.bias, out, None, 1, K.ACT_TANH, conv.wpack,
                   rt.use_tc(K.OP_DGRAD, N, H, W, A, B, 1))
    return out


def final_conv_backward(rt, conv, x, dpre, wgrad, nb=None):
    """dpre: gradient w.r.t. the pre-tanh output.  Returns the gradient w.r.t. x (with `nb`, a
    lg_norm_bwd_t of the decoder's last layer: dy of that layer, pass 1 of its norm backward done)."""
    N, H, W, B = x.shape
    A = conv.filters
    dpad = _padded_for_tc(rt, dpre, K.OP_FPROP, B, 1)
    if wgrad:
        def _wg():
            K.bias_grad(dpre, _grad(conv.bias))
            if dpad is not None and rt.use_tc(K.OP_WGRAD, N, H, W, 16, B, 1):
                K.conv2d_wgrad_padded(dpad, x, _grad(conv.kernel), 1)
            else:
                K.conv2d_wgrad(dpre, x, _grad(conv.kernel), 1, rt.use_tc(K.OP_WGRAD, N, H, W, A, B, 1))
        rt.on_side(_wg, dpre, dpad, x)
    g = torch.empty_like(x)
    if dpad is not None:
        K.conv2d_fprop(dpad, conv.kernel, None, g, None, 1, conv.wpack, True, norm_bwd=nb)
    else:
        _fprop(rt, conv, dpre, None, g, None, 1, nb)
    return g


def generator_tail_backward(rt, dec, conv, dctx, x4, dpre, wgrad):
    """final conv backward + decoder backward (Generator / Adjuster share the tail, model.py:103-104,
    135-136); the final conv's input-gradient launch carries pass 1 of the last decoder norm backward."""
    N, H, W, B = x4.shape
    A = conv.filters
    red, nb = None, None
    Ain = 16 if (_would_pad(rt, N, H, W, A, B)) else A
    if _fprop_fuses_norm_bwd(rt, conv, N, H, W, A, B, 1) or rt.fuse_norm_bwd(K.OP_FPROP, N, H, W, Ain, B, 1):
        red = rt.zeros(4, N, 2)
        _, z3, s3 = dctx[3]
        nb = K.norm_bwd_desc(z3, s3, dec.norms[3].gamma, dec.norms[3].beta, red[3], dec.norms[3].epsilon, rt.alpha)
    g = final_conv_backward(rt, conv, x4, dpre, wgrad, nb)
    return decoder_backward(rt, dec, dctx, g, wgrad, red, nb is not None)


def _would_pad(rt, N, H, W, A, B):
    """Mirrors _padded_for_tc's decision for a [N,H,W,A] gradient feeding the stride-1 fprop."""
    if A >= 16 or not rt.want_tc or rt.use_tc(K.OP_FPROP, N, H, W, A, B, 1):
        return False
    return rt.use_tc(K.OP_FPROP, N, H, W, 16, B, 1)


# --------------------------------------------------------------------------------------------
# dense -> LeakyReLU -> IN head of Generator / Adjuster (model.py:98-102, 129-132); fp32 inside
# --------------------------------------------------------------------------------------------
def head_forward(rt, dense, norm, xin, out_shape, skip=None):
    N, Kin = xin.shape
    F = dense.units
    h = rt.empty(N, F, dtype=torch.float32)
    K.gemm(xin, dense.kernel, h, N, F, Kin, bias=dense.bias)
    stats = rt.zeros(N, 2)
    K.rowstats(h, stats, rt.alpha)
    out = rt.empty(*out_shape)
    _norm_act_fwd(h, stats, norm, skip, out, rt.alpha, 1.0)
    return out, (xin, h, stats)


def head_backward(rt, dense, norm, ctx, g):
    xin, h, stats = ctx
    N, Kin = xin.shape
    F = dense.units
    red = rt.zeros(N, 2)
    dh = torch.empty_like(h)
    K.instnorm_act_bwd(g.view(N, F), h, stats, norm.gamma, norm.beta, red, dh, _grad(norm.gamma), _grad(norm.beta),
                       norm.epsilon, rt.alpha, 1.0)
    K.bias_grad(dh, _grad(dense.bias))
    K.gemm(xin, dh, _grad(dense.kernel), Kin, F, N, transA=True, accumulate=True)


# --------------------------------------------------------------------------------------------
# discriminator heads (model.py:62-63,70-72): flatten(HWC) -> Dense(1, sigmoid), Dense(cond, sigmoid)
# --------------------------------------------------------------------------------------------
def disc_heads_forward(rt, disc, feat):
    N = feat.shape[0]
    f = feat.view(N, -1)
    F = f.shape[1]
    U0, U1 = disc.dense_pr.units, disc.dense_cond.units
    if K.dense_heads_supported(F, U0, U1):
        pr = rt.empty(N, U0, dtype=torch.float32)
        c = rt.empty(N, U1, dtype=torch.float32)
        K.dense_heads_fwd(f, disc.dense_pr.kernel, disc.dense_pr.bias, disc.dense_cond.kernel, disc.dense_cond.bias,
                          pr, c, K.ACT_SIGMOID, rt.heads_workspace(N))
        return pr, c
    outs = []
    for dense in (disc.dense_pr, disc.dense_cond):
        o = rt.zeros(N, dense.units, dtype=torch.float32)
        K.gemm(f, dense.kernel, o, N, dense.units, F, accumulate=True)
        K.bias_act(o, dense.bias, K.ACT_SIGMOID)
        outs.append(o)
    return outs[0], outs[1]


def disc_heads_backward(rt, disc, feat, dl_pr, dl_c, wgrad):
    """dl_*: gradients w.r.t. the pre-sigmoid logits (fp32).  Returns d(feat)."""
    N = feat.shape[0]
    f = feat.view(N, -1)
    F = f.shape[1]
    dp, dc = disc.dense_pr, disc.dense_cond
    if K.dense_heads_supported(F, dp.units, dc.units):
        df = torch.empty_like(f)
        if wgrad:
            K.dense_heads_bwd(f, dl_pr, dl_c, dp.kernel, dc.kernel, df, _grad(dp.kernel), _grad(dc.kernel),
                              _grad(dp.bias), _grad(dc.bias))
        else:
            K.dense_heads_bwd(None, dl_pr, dl_c, dp.kernel, dc.kernel, df, None, None, None, None)
        return df.view_as(feat)
    df = rt.zeros(N, F, dtype=torch.float32)
    for dense, dl in ((dp, dl_pr), (dc, dl_c)):
        if dl is None:
            continue
        U = dense.units
        if wgrad:
            K.gemm(f, dl, _grad(dense.kernel), F, U, N, transA=True, accumulate=True)
            K.bias_grad(dl, _grad(dense.bias))
        K.gemm(dl, dense.kernel, df, N, F, U, transB=True, accumulate=True)
    if rt.act_dtype == torch.float32:
        return df.view_as(feat)
    return K.cast(df, torch.empty_like(feat))
