"""Thin torch-tensor wrappers over the C-ABI (one Python function per entry point).

PyTorch is plumbing here: device memory, streams, distributed.  Every function launches
hand-written sm_100a kernels on torch's current CUDA stream and returns immediately.
"""
import torch

from . import _lib
from ._lib import ACT_NONE, ACT_SIGMOID, ACT_TANH, BF16, F32, OP_DGRAD, OP_FPROP, OP_WGRAD  # noqa: F401

# every C-ABI compute call launches exactly one kernel of this library; counted for bench.py
_LAUNCHES = [0]
_LAST_CAPTURE = [0]


def check(ret, what=""):
    _LAUNCHES[0] += 1
    return _lib.check(ret, what)


def launch_count():
    return _LAUNCHES[0]


def note_capture(n):
    _LAST_CAPTURE[0] = n


def launch_count_of_last_capture():
    return _LAST_CAPTURE[0]

_DT = {torch.float32: F32, torch.bfloat16: BF16}

# which kernel family each conv launch took: ("fprop"|"dgrad"|"wgrad", "tc"|"rows"|"simt") -> launches
_PATHS = {}
_WARNED = set()


def path_counts(reset=False):
    """Conv launches so far by (operation, kernel family); tests assert that a bf16 run never took "simt"."""
    out = dict(_PATHS)
    if reset:
        _PATHS.clear()
    return out


def _note_path(op, family, t, geom):
    key = (op, family)
    _PATHS[key] = _PATHS.get(key, 0) + 1
    if family == "simt" and t.dtype == torch.bfloat16 and (op, geom) not in _WARNED:
        # bf16 maps are the tensor-core mode: a geometry the tcgen05 planner rejects still computes correctly on
        # the fp32-FMA kernel, but ~10x slower - say so once per geometry instead of silently
        _WARNED.add((op, geom))
        import warnings
        warnings.warn("littlegan_b200: %s with geometry (N,Hb,Wb,A,B,stride)=%s has no tcgen05 form (channel counts "
                      "must be multiples of 16 and the maps large enough for TMA tiles); running it on the much "
                      "slower SIMT fp32-FMA kernel" % (op, geom), RuntimeWarning, stacklevel=3)


def _p(t):
    return None if t is None else t.data_ptr()


def _st():
    return torch.cuda.current_stream().cuda_stream


def _cuda(*ts):
    for t in ts:
        if t is not None:
            if not t.is_cuda:
                raise _lib.LittleGANError("littlegan_b200 ops need CUDA tensors (no CPU fallback)")
            if not t.is_contiguous():
                raise _lib.LittleGANError("littlegan_b200 ops need contiguous (NHWC) tensors")


def dt(t):
    return _DT[t.dtype]


def tc_supported(op, N, Hb, Wb, A, B, stride):
    return bool(_lib.load().lg_conv2d_tc_supported(op, N, Hb, Wb, A, B, stride))


def tc_available():
    return bool(_lib.load().lg_tensor_core_path_available())


def set_cta_pairs(mode):
    """0 never / 1 auto / 2 always use CTA pairs (cta_group::2) in the generic conv kernels; returns the old mode."""
    return _lib.load().lg_set_cta_pairs(int(mode))


# ------------------------------------------------------------------------------------------- conv
def norm_bwd_supported(op, N, Hb, Wb, A, B, stride):
    return bool(_lib.load().lg_conv2d_norm_bwd_supported(op, N, Hb, Wb, A, B, stride))


def norm_bwd_desc(z, stats, gamma, beta, red, eps, alpha):
    """lg_norm_bwd_t for a backward conv launch whose output is dL/d(LeakyReLU(IN(z)))."""
    _cuda(z, stats, gamma, beta, red)
    if z.dtype != torch.bfloat16:
        raise _lib.LittleGANError("the fused norm-backward epilogue needs bf16 activations")
    d = _lib.NormBwd(_p(z), _p(stats), _p(gamma), _p(beta), _p(red), eps, alpha)
    d._keep = (z, stats, gamma, beta, red)      # the descriptor holds raw pointers: keep the tensors alive
    return d


def _nb(desc):
    import ctypes
    return None if desc is None else ctypes.byref(desc)


def conv2d_fprop(big, W, bias, out, stats, stride, wpack=None, use_tc=False, norm_bwd=None):
    """small = conv(big): big [N,Hb,Wb,A], W [5,5,A,B] fp32, out [N,Hb/s,Wb/s,B]."""
    _cuda(big, W, bias, out, stats, wpack)
    N, Hb, Wb, A = big.shape
    B = out.shape[3]
    _note_path("fprop", "tc" if use_tc else "simt", big, (N, Hb, Wb, A, B, stride))
    check(_lib.load().lg_conv2d_fprop(_p(big), _p(W), _p(wpack), _p(bias), _p(out), _p(stats), N, Hb, Wb, A, B,
                                      stride, dt(big), int(use_tc), _nb(norm_bwd), _st()), "lg_conv2d_fprop")
    return out


def fprop_rows_supported(N, Hb, Wb, A_big, A, B, stride):
    return bool(_lib.load().lg_conv2d_fprop_rows_supported(N, Hb, Wb, A_big, A, B, stride))


def pack_rowconv_weights(W, A_big, stride, wpack=None):
    """bf16 weight operand of the row-streaming kernel for W [5,5,A,B] (refresh after every weight update)."""
    A, B = W.shape[2], W.shape[3]
    lib = _lib.load()
    if wpack is None:
        nbytes = int(_lib.check(lib.lg_pack_rowconv_weights(None, None, A_big, A, B, stride, None)))
        wpack = torch.empty(nbytes, dtype=torch.uint8, device=W.device)
    _cuda(W, wpack)
    check(lib.lg_pack_rowconv_weights(_p(W), _p(wpack), A_big, A, B, stride, _st()), "lg_pack_rowconv_weights")
    return wpack


def conv2d_fprop_rows(big, wpack, bias, out, stats, stride, A, norm_bwd=None):
    """Row-streaming tcgen05 fprop: big [N,Hb,128,A_big] bf16 (A_big = 8 holds a zero-padded RGB image),
    wpack from pack_rowconv_weights(W [5,5,A,B])."""
    _cuda(big, wpack, bias, out, stats)
    N, Hb, Wb, A_big = big.shape
    B = out.shape[3]
    _note_path("fprop", "rows", big, (N, Hb, Wb, A, B, stride))
    check(_lib.load().lg_conv2d_fprop_rows(_p(big), A_big, _p(wpack), _p(bias), _p(out), _p(stats), N, Hb, Wb, A, B,
                                           stride, _nb(norm_bwd), _st()), "lg_conv2d_fprop_rows")
    return out


def conv2d_dgrad(small, W, bias, out, stats, stride, act=ACT_NONE, wpack=None, use_tc=False, norm_bwd=None):
    """big = conv_transpose(small): small [N,Hs,Ws,B], W [5,5,A,B] fp32, out [N,s*Hs,s*Ws,A]."""
    _cuda(small, W, bias, out, stats, wpack)
    N, Hb, Wb, A = out.shape
    B = small.shape[3]
    _note_path("dgrad", "tc" if use_tc else "simt", small, (N, Hb, Wb, A, B, stride))
    check(_lib.load().lg_conv2d_dgrad(_p(small), _p(W), _p(wpack), _p(bias), _p(out), _p(stats), N, Hb, Wb, A, B,
                                      stride, act, dt(small), int(use_tc), _nb(norm_bwd), _st()), "lg_conv2d_dgrad")
    return out


def dgrad_rows_supported(N, Hb, Wb, A, B, stride):
    return bool(_lib.load().lg_conv2d_dgrad_rows_supported(N, Hb, Wb, A, B, stride))


def pack_rowdgrad_weights(W, wpack=None):
    """bf16 weight operand of the row-streaming dgrad kernel for W [5,5,32,64]."""
    A, B = W.shape[2], W.shape[3]
    lib = _lib.load()
    if wpack is None:
        nbytes = int(_lib.check(lib.lg_pack_rowdgrad_weights(None, None, A, B, None)))
        wpack = torch.empty(nbytes, dtype=torch.uint8, device=W.device)
    _cuda(W, wpack)
    check(lib.lg_pack_rowdgrad_weights(_p(W), _p(wpack), A, B, _st()), "lg_pack_rowdgrad_weights")
    return wpack


def conv2d_dgrad_rows(small, wpack, bias, out, stats, stride):
    """Row-streaming tcgen05 dgrad (decoder conv4 forward): small [N,Hs,64,64] -> out [N,2Hs,128,32]."""
    _cuda(small, wpack, bias, out, stats)
    N, Hb, Wb, A = out.shape
    B = small.shape[3]
    _note_path("dgrad", "rows", small, (N, Hb, Wb, A, B, stride))
    check(_lib.load().lg_conv2d_dgrad_rows(_p(small), _p(wpack), _p(bias), _p(out), _p(stats), N, Hb, Wb, A, B,
                                           stride, _st()), "lg_conv2d_dgrad_rows")
    return out


def dgrad_rgb_supported(N, Hb, Wb, A, B, stride):
    return bool(_lib.load().lg_conv2d_dgrad_rgb_supported(N, Hb, Wb, A, B, stride))


def conv2d_dgrad_rgb(small, W, bias, out, out_pad8, stats, stride, act=ACT_NONE):
    """RGB transposed conv on the row-streaming kernel; out_pad8 (or None): [N,H,W,8] zero-padded copy."""
    _cuda(small, W, bias, out, out_pad8, stats)
    N, Hb, Wb, A = out.shape
    B = small.shape[3]
    _note_path("dgrad", "rows", small, (N, Hb, Wb, A, B, stride))
    check(_lib.load().lg_conv2d_dgrad_rgb(_p(small), _p(W), _p(bias), _p(out), _p(out_pad8), _p(stats), N, Hb, Wb,
                                          A, B, stride, act, _st()), "lg_conv2d_dgrad_rgb")
    return out


def conv2d_wgrad(big, small, dW, stride, use_tc=False):
    """dW[5,5,A,B] += correlation(big, small)."""
    _cuda(big, small, dW)
    N, Hb, Wb, A = big.shape
    B = small.shape[3]
    _note_path("wgrad", "tc" if use_tc else "simt", big, (N, Hb, Wb, A, B, stride))
    check(_lib.load().lg_conv2d_wgrad(_p(big), _p(small), _p(dW), N, Hb, Wb, A, B, stride, dt(big), int(use_tc),
                                      _st()), "lg_conv2d_wgrad")


def conv2d_wgrad_padded(big_padded, small, dW, stride):
    """tcgen05 wgrad with a channel-padded `big` ([N,Hb,Wb,A_big], A_big >= dW.shape[2])."""
    _cuda(big_padded, small, dW)
    N, Hb, Wb, A_big = big_padded.shape
    A, B = dW.shape[2], dW.shape[3]
    _note_path("wgrad", "tc", big_padded, (N, Hb, Wb, A, B, stride))
    check(_lib.load().lg_conv2d_wgrad_padded(_p(big_padded), _p(small), _p(dW), N, Hb, Wb, A_big, A, B, stride,
                                             _st()), "lg_conv2d_wgrad_padded")


def pad_channels(src, dst):
    _cuda(src, dst)
    C, Cp = src.shape[-1], dst.shape[-1]
    check(_lib.load().lg_pad_channels(_p(src), _p(dst), src.numel() // C, C, Cp, dt(src), _st()), "lg_pad_channels")
    return dst


def pack_conv_weights_bytes(A, B):
    return int(_lib.check(_lib.load().lg_pack_conv_weights(None, None, A, B, None)))


def pack_conv_weights(W, wpack):
    _cuda(W, wpack)
    A, B = W.shape[2], W.shape[3]
    check(_lib.load().lg_pack_conv_weights(_p(W), _p(wpack), A, B, _st()), "lg_pack_conv_weights")


def pack_conv_weights_multi(Ws, wpacks):
    """pack_conv_weights for several layers in one launch."""
    import ctypes
    lib = _lib.load()
    for i in range(0, len(Ws), _lib.PACK_MAX):
        W, P = Ws[i:i + _lib.PACK_MAX], wpacks[i:i + _lib.PACK_MAX]
        _cuda(*W, *P)
        n = len(W)
        wp = (ctypes.c_void_p * n)(*[w.data_ptr() for w in W])
        pp = (ctypes.c_void_p * n)(*[q.data_ptr() for q in P])
        aa = (ctypes.c_int * n)(*[w.shape[2] for w in W])
        bb = (ctypes.c_int * n)(*[w.shape[3] for w in W])
        check(lib.lg_pack_conv_weights_multi(wp, pp, aa, bb, n, _st()), "lg_pack_conv_weights_multi")


def bias_grad(g, db):
    _cuda(g, db)
    C = g.shape[-1]
    check(_lib.load().lg_bias_grad(_p(g), _p(db), g.numel() // C, C, dt(g), _st()), "lg_bias_grad")


# ------------------------------------------------------------------------------------------- norm
def rowstats(z, stats, alpha_pre=1.0):
    _cuda(z, stats)
    N = z.shape[0]
    check(_lib.load().lg_rowstats(_p(z), _p(stats), N, z.numel() // N, alpha_pre, dt(z), _st()), "lg_rowstats")


def instnorm_act_fwd(z, stats, gamma, beta, skip, out, eps, alpha_pre, alpha_post):
    _cuda(z, stats, gamma, beta, skip, out)
    N = z.shape[0]
    check(_lib.load().lg_instnorm_act_fwd(_p(z), _p(stats), _p(gamma), _p(beta), _p(skip), _p(out), N,
                                          z.numel() // N, eps, alpha_pre, alpha_post, dt(z), dt(out), _st()),
          "lg_instnorm_act_fwd")
    return out


def bias_grad_fusable(z):
    """Can lg_instnorm_act_bwd_apply also produce the bias gradient of the conv that wrote z [N,...,C]?"""
    N = z.shape[0]
    return bool(_lib.load().lg_instnorm_bias_grad_fusable(z.numel() // N, z.shape[-1], dt(z)))


def instnorm_act_bwd(g, z, stats, gamma, beta, red, dz, dgamma, dbeta, eps, alpha_pre, alpha_post, dy_ready=False,
                     dbias=None):
    """Two passes: per-sample reductions into `red` (zeroed by the caller), then dz.  dy_ready: pass 1 was
    done by the conv launch that produced `g` (norm_bwd epilogue).  dbias: also accumulate the per-channel
    sums of dz (needs bias_grad_fusable(z))."""
    _cuda(g, z, stats, gamma, beta, red, dz, dgamma, dbeta, dbias)
    N = z.shape[0]
    M = z.numel() // N
    lib = _lib.load()
    if not dy_ready:
        check(lib.lg_instnorm_act_bwd_reduce(_p(g), _p(z), _p(stats), _p(gamma), _p(beta), _p(red), N, M, eps,
                                             alpha_pre, alpha_post, dt(z), dt(g), _st()),
              "lg_instnorm_act_bwd_reduce")
    check(lib.lg_instnorm_act_bwd_apply(_p(g), _p(z), _p(stats), _p(red), _p(gamma), _p(beta), _p(dz), _p(dgamma),
                                        _p(dbeta), _p(dbias), z.shape[-1], int(dy_ready), N, M, eps, alpha_pre,
                                        alpha_post, dt(z), dt(g), _st()), "lg_instnorm_act_bwd_apply")
    return dz


# ------------------------------------------------------------------------------------------ dense
def gemm(A, Bm, C, M, N, K, bias=None, transA=False, transB=False, accumulate=False):
    _cuda(A, Bm, C, bias)
    check(_lib.load().lg_gemm(_p(A), _p(Bm), _p(bias), _p(C), M, N, K, int(transA), int(transB), int(accumulate),
                              dt(A), dt(C), _st()), "lg_gemm")
    return C


def dense_heads_supported(F, U0, U1):
    return U0 + U1 <= 48 and F % 64 == 0


def dense_heads_workspace(N, device):
    """Zero-filled scratch of the fused heads forward (the kernel leaves it zeroed after every call)."""
    nbytes = int(_lib.check(_lib.load().lg_dense_heads_workspace_bytes(N)))
    return torch.zeros(nbytes, dtype=torch.uint8, device=device)


def dense_heads_fwd(feat, W0, b0, W1, b1, out0, out1, act, workspace):
    """out_h = act(feat @ W_h + b_h) for two Dense layers sharing `feat` [N,F] (model.py:62-63,70-72)."""
    _cuda(feat, W0, b0, W1, b1, out0, out1, workspace)
    N, F = feat.shape
    check(_lib.load().lg_dense_heads_fwd(_p(feat), _p(W0), _p(b0), W0.shape[1], _p(W1), _p(b1), W1.shape[1],
                                         _p(out0), _p(out1), N, F, act, dt(feat), _p(workspace), _st()),
          "lg_dense_heads_fwd")


def dense_heads_bwd(feat, dl0, dl1, W0, W1, dfeat, dW0, dW1, db0, db1):
    _cuda(feat, dl0, dl1, W0, W1, dfeat, dW0, dW1, db0, db1)
    ref = feat if feat is not None else dfeat
    N, F = ref.shape
    check(_lib.load().lg_dense_heads_bwd(_p(feat), _p(dl0), _p(dl1), _p(W0), W0.shape[1], _p(W1), W1.shape[1],
                                         _p(dfeat), _p(dW0), _p(dW1), _p(db0), _p(db1), N, F, dt(ref), _st()),
          "lg_dense_heads_bwd")


def bias_act(x, bias, act):
    _cuda(x, bias)
    rows, cols = x.shape
    check(_lib.load().lg_bias_act(_p(x), _p(bias), rows, cols, act, _st()), "lg_bias_act")
    return x


# ----------------------------------------------------------------------------------------- losses
def bce_sigmoid(p, target, weight, loss_accum, dlogit):
    """target: tensor [rows, cols] or python float (constant target)."""
    rows, cols = p.shape
    tt = target if torch.is_tensor(target) else None
    tc = 0.0 if tt is not None else float(target)
    _cuda(p, tt, loss_accum, dlogit)
    check(_lib.load().lg_bce_sigmoid(_p(p), _p(tt), tc, rows, cols, weight, _p(loss_accum), _p(dlogit), _st()),
          "lg_bce_sigmoid")


def bce_sigmoid_multi(terms):
    """Several bce_sigmoid terms in one launch; terms = [(p, target, weight, loss_accum, dlogit), ...]."""
    lib = _lib.load()
    for i in range(0, len(terms), _lib.BCE_MAX):
        chunk = terms[i:i + _lib.BCE_MAX]
        items = (_lib.BceItem * len(chunk))()
        for it, (p, target, weight, loss_accum, dlogit) in zip(items, chunk):
            tt = target if torch.is_tensor(target) else None
            _cuda(p, tt, loss_accum, dlogit)
            it.p, it.target, it.loss_accum, it.dlogit = _p(p), _p(tt), _p(loss_accum), _p(dlogit)
            it.target_const = 0.0 if tt is not None else float(target)
            it.weight, it.n = float(weight), p.numel()
        check(lib.lg_bce_sigmoid_multi(items, len(chunk), _st()), "lg_bce_sigmoid_multi")


def l1_tanh_bwd(y, t, g_in, dpre, weight, loss_accum):
    _cuda(y, t, g_in, dpre, loss_accum)
    check(_lib.load().lg_l1_tanh_bwd(_p(y), _p(t), _p(g_in), _p(dpre), y.numel(), weight, _p(loss_accum), dt(y),
                                     _st()), "lg_l1_tanh_bwd")


# -------------------------------------------------------------------------------------- optimiser
def adam_advance(state, lr, beta1, beta2):
    _cuda(state)
    check(_lib.load().lg_adam_advance(_p(state), lr, beta1, beta2, _st()), "lg_adam_advance")


def adam_apply(p, g, m, v, state, beta1, beta2, eps, clip):
    _cuda(p, g, m, v, state)
    check(_lib.load().lg_adam_apply(_p(p), _p(g), _p(m), _p(v), p.numel(), _p(state), beta1, beta2, eps, clip,
                                    _st()), "lg_adam_apply")


def cast(src, dst):
    _cuda(src, dst)
    check(_lib.load().lg_cast(_p(src), _p(dst), src.numel(), dt(src), dt(dst), _st()), "lg_cast")
    return dst


def scale_shift(src, dst, scale=1.0, shift=0.0):
    """dst = src * scale + shift (fp32 arithmetic; fp32 / bf16 either side).  scale 1, shift 0 = a device copy."""
    _cuda(src, dst)
    if src.numel() != dst.numel():
        raise _lib.LittleGANError("scale_shift: size mismatch")
    check(_lib.load().lg_scale_shift(_p(src), _p(dst), src.numel(), float(scale), float(shift), dt(src), dt(dst),
                                     _st()), "lg_scale_shift")
    return dst


def normal_state(seed, device):
    """{seed, step, ticket} of lg_normal_fill's Philox stream, on the device."""
    return torch.tensor([int(seed), 0, 0], dtype=torch.int64, device=device)


def normal_fill(out, state):
    """out (fp32, CUDA) ~ N(0, 1); advances state[1] on the device."""
    _cuda(out, state)
    if out.dtype != torch.float32:
        raise _lib.LittleGANError("normal_fill: needs an fp32 buffer")
    check(_lib.load().lg_normal_fill(_p(out), out.numel(), _p(state), _st()), "lg_normal_fill")
    return out


def u8_rescale(src, dst):
    """data_rescale (utils.py:51-52) of decoded image bytes on the device: dst = src / 127.5 - 1."""
    _cuda(src, dst)
    if src.dtype != torch.uint8 or src.numel() != dst.numel() or not (src.is_contiguous() and dst.is_contiguous()):
        raise _lib.LittleGANError("u8_rescale: needs a contiguous uint8 source and a destination of the same size")
    check(_lib.load().lg_u8_rescale(_p(src), _p(dst), src.numel(), dt(dst), _st()), "lg_u8_rescale")
    return dst


# -------------------------------------------------------------------------------------------- FID
def fid_accumulate(X, shift, S1, S2):
    _cuda(X, shift, S1, S2)
    n, d = X.shape
    check(_lib.load().lg_fid_accumulate(_p(X), _p(shift), _p(S1), _p(S2), n, d, _st()), "lg_fid_accumulate")


def fid_finalize(S1, S2, shift, mu, sigma, n):
    _cuda(S1, S2, shift, mu, sigma)
    check(_lib.load().lg_fid_finalize(_p(S1), _p(S2), _p(shift), _p(mu), _p(sigma), n, S1.numel(), _st()),
          "lg_fid_finalize")


def dgemm(A, B, C, alpha=1.0, diag=0.0):
    """C = alpha * A @ B + diag * I on fp64 [n,n] matrices (csrc/fid.cu)."""
    _cuda(A, B, C)
    n = A.shape[0]
    if not (A.dtype == B.dtype == C.dtype == torch.float64 and A.shape == B.shape == C.shape == (n, n)):
        raise _lib.LittleGANError("dgemm: needs three fp64 [n,n] matrices")
    check(_lib.load().lg_dgemm(_p(A), _p(B), _p(C), n, float(alpha), float(diag), _st()), "lg_dgemm")
    return C


def dmat_stats(A, out3):
    """out3 (3 device doubles) = [trace(A), ||A||_F^2, max |A - A^T|]."""
    _cuda(A, out3)
    check(_lib.load().lg_dmat_stats(_p(A), A.shape[0], _p(out3), _st()), "lg_dmat_stats")
    return out3


def dmat_scale_shift(src, dst, alpha=1.0, diag=0.0, symmetrise=False):
    """dst = alpha * (src | (src + src^T)/2 | 0) + diag * I."""
    _cuda(src, dst)
    check(_lib.load().lg_dmat_scale_shift(_p(src), _p(dst), dst.shape[0], float(alpha), float(diag), int(symmetrise),
                                          _st()), "lg_dmat_scale_shift")
    return dst


# ------------------------------------------------------------------------------------ augmentation
AUG_MAX_BRIGHTNESS, AUG_CONTRAST, AUG_MAX_HUE, AUG_NOISE_STD = 0.02, (0.75, 1.003), 0.03, 0.1 * 0.2


def augment_state(seed, device):
    """{seed, step} of the augmentation's Philox streams, on the device (the step advances inside the launches)."""
    return torch.tensor([int(seed), 0], dtype=torch.int64, device=device)


def augment(x, out, params, state=None, noise=None, draw=True):
    """new_image of eager_trainer.py:127-131: x [N,H,W,3] fp32 -> out (same shape, fp32 or bf16).
    params: float32 [4 + 4 N] scratch (draw=False: [0:3] and the flip flags at [7::4] hold the caller's draws).
    noise: None = N(0, 0.02) from Philox(state), else a [N,H,W,3] fp32 tensor added as is."""
    _cuda(x, out, params, state, noise)
    N, H, W, C = x.shape
    if C != 3 or x.dtype != torch.float32:
        raise _lib.LittleGANError("augment: needs an fp32 RGB batch")
    lib = _lib.load()
    check(lib.lg_augment_prepare(_p(x), N, H, W, _p(state), _p(params), AUG_MAX_BRIGHTNESS, AUG_CONTRAST[0],
                                 AUG_CONTRAST[1], AUG_MAX_HUE, int(draw), _st()), "lg_augment_prepare")
    check(lib.lg_augment_apply(_p(x), _p(params), _p(noise), _p(state), AUG_NOISE_STD, _p(out), N, H, W, dt(out),
                               _st()), "lg_augment_apply")
    return out


# ------------------------------------------------------------------- Inception pool_3 forward (FID)
POOL_MAX, POOL_AVG_VALID, POOL_AVG_PADDED = 0, 1, 2


def pack_conv_bn_weights(W):
    """bf16 tensor-core operand of a unit's kernel W [kh,kw,cin,cout] (fp32, CUDA), or None when the geometry has no
    tensor-core form."""
    _cuda(W)
    kh, kw, ci, co = W.shape
    lib = _lib.load()
    nbytes = lib.lg_pack_conv_bn_weights(None, None, ci, kh, kw, co, None)
    if nbytes <= 0:
        return None
    wpack = torch.empty(nbytes, dtype=torch.uint8, device=W.device)
    check(lib.lg_pack_conv_bn_weights(_p(W), _p(wpack), ci, kh, kw, co, _st()), "lg_pack_conv_bn_weights")
    return wpack


def conv2d_bn_relu(x, W, scale, shift, y, y_off, stride=1, pad=(0, 0), x_off=0, cin=None, relu=True, wpack=None):
    """One conv + folded-BN + ReLU unit: x [N,H,W,Cx] (channels [x_off, x_off+cin)), W fp32 [kh,kw,cin,cout] ->
    channels [y_off, y_off+cout) of y [N,Ho,Wo,Cy].  bf16 maps with `wpack` run on the tensor cores."""
    _cuda(x, W, scale, shift, y, wpack)
    N, H, Wd, Cx = x.shape
    kh, kw, ci, co = W.shape
    cin = Cx - x_off if cin is None else cin
    Ho, Wo = (H + 2 * pad[0] - kh) // stride + 1, (Wd + 2 * pad[1] - kw) // stride + 1
    if ci != cin or tuple(y.shape[:3]) != (N, Ho, Wo) or x.dtype != y.dtype or W.dtype != torch.float32:
        raise _lib.LittleGANError("conv2d_bn_relu: shapes / dtypes do not match the geometry")
    check(_lib.load().lg_conv2d_bn_relu(_p(x), _p(W), _p(wpack), _p(scale), _p(shift), _p(y), N, H, Wd, cin, Cx, x_off,
                                        kh, kw, stride, pad[0], pad[1], co, y.shape[3], y_off, int(relu), dt(x), _st()),
          "lg_conv2d_bn_relu")
    return y


def pool2d(x, y, y_off, k, stride, pad, mode):
    _cuda(x, y)
    N, H, Wd, C = x.shape
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (Wd + 2 * pad - k) // stride + 1
    if tuple(y.shape[:3]) != (N, Ho, Wo) or x.dtype != y.dtype:
        raise _lib.LittleGANError("pool2d: output shape / dtype does not match the geometry")
    check(_lib.load().lg_pool2d(_p(x), _p(y), N, H, Wd, C, C, 0, k, stride, pad, mode, y.shape[3], y_off, dt(x), _st()),
          "lg_pool2d")
    return y


def global_avgpool(x, y):
    _cuda(x, y)
    N, H, Wd, C = x.shape
    check(_lib.load().lg_global_avgpool(_p(x), _p(y), N, H * Wd, C, dt(x), _st()), "lg_global_avgpool")
    return y


def resize_bilinear_norm(x, y, sub, mul):
    """TF-1.x ResizeBilinear of x [N,H,W,C] (uint8 or fp32, 0..255) to y's spatial size, then (v - sub) * mul; y may
    have more channels than x (zero filled)."""
    _cuda(x, y)
    if x.dtype not in (torch.uint8, torch.float32):
        raise _lib.LittleGANError("resize_bilinear_norm: source must be uint8 or fp32")
    N, H, Wd, C = x.shape
    if y.shape[0] != N or y.shape[3] < C:
        raise _lib.LittleGANError("resize_bilinear_norm: output batch / channels do not fit the source")
    check(_lib.load().lg_resize_bilinear_norm(_p(x), _p(y), N, H, Wd, C, y.shape[3], y.shape[1], y.shape[2], sub, mul,
                                              int(x.dtype == torch.uint8), dt(y), _st()), "lg_resize_bilinear_norm")
    return y
