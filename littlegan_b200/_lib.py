"""ctypes binding of the C-ABI in include/littlegan_b200.h.

The shared object is built in-tree by `littlegan_b200/csrc/build.py` (called from
`__graft_entry__.build()`).  There is NO CPU fallback: importing this module without the
library, or calling an op without a CUDA device, raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblittlegan_b200.so")

F32, BF16 = 0, 1
ACT_NONE, ACT_TANH, ACT_SIGMOID = 0, 1, 2
OP_FPROP, OP_DGRAD, OP_WGRAD = 0, 1, 2

_vp, _i, _i64, _f, _d = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double

# name -> (restype, argtypes); mirrors include/littlegan_b200.h line by line
SIGNATURES = {
    "lg_abi_version": (_i, []),
    "lg_last_error": (C.c_char_p, []),
    "lg_tensor_core_path_available": (_i, []),
    "lg_conv2d_tc_supported": (_i, [_i] * 7),
    "lg_set_cta_pairs": (_i, [_i]),
    "lg_conv2d_norm_bwd_supported": (_i, [_i] * 7),
    "lg_conv2d_fprop": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "lg_conv2d_fprop_rows_supported": (_i, [_i, _i, _i, _i, _i, _i, _i]),
    "lg_pack_rowconv_weights": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "lg_conv2d_fprop_rows": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "lg_conv2d_dgrad": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "lg_conv2d_dgrad_rows_supported": (_i, [_i, _i, _i, _i, _i, _i]),
    "lg_pack_rowdgrad_weights": (_i, [_vp, _vp, _i, _i, _vp]),
    "lg_conv2d_dgrad_rows": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "lg_conv2d_dgrad_rgb_supported": (_i, [_i, _i, _i, _i, _i, _i]),
    "lg_conv2d_dgrad_rgb": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "lg_conv2d_wgrad": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "lg_conv2d_wgrad_padded": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "lg_pad_channels": (_i, [_vp, _vp, _i64, _i, _i, _i, _vp]),
    "lg_conv2d_transpose_fprop": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "lg_conv2d_transpose_dgrad": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "lg_pack_conv_weights": (_i64, [_vp, _vp, _i, _i, _vp]),
    "lg_pack_conv_weights_multi": (_i, [_vp, _vp, _vp, _vp, _i, _vp]),
    "lg_bias_grad": (_i, [_vp, _vp, _i64, _i, _i, _vp]),
    "lg_rowstats": (_i, [_vp, _vp, _i, _i64, _f, _i, _vp]),
    "lg_instnorm_act_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i64, _f, _f, _f, _i, _i, _vp]),
    "lg_instnorm_act_bwd_reduce": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i64, _f, _f, _f, _i, _i, _vp]),
    "lg_instnorm_bias_grad_fusable": (_i, [_i64, _i, _i]),
    "lg_instnorm_act_bwd_apply": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i64, _f, _f, _f, _i,
                                       _i, _vp]),
    "lg_gemm": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "lg_dense_heads_workspace_bytes": (_i64, [_i]),
    "lg_dense_heads_fwd": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _i, _vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "lg_dense_heads_bwd": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "lg_bias_act": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "lg_bce_sigmoid": (_i, [_vp, _vp, _f, _i, _i, _f, _vp, _vp, _vp]),
    "lg_bce_sigmoid_multi": (_i, [_vp, _i, _vp]),
    "lg_l1_tanh_bwd": (_i, [_vp, _vp, _vp, _vp, _i64, _f, _vp, _i, _vp]),
    "lg_adam_advance": (_i, [_vp, _d, _d, _d, _vp]),
    "lg_adam_apply": (_i, [_vp, _vp, _vp, _vp, _i64, _vp, _f, _f, _f, _f, _vp]),
    "lg_augment_prepare": (_i, [_vp, _i, _i, _i, _vp, _vp, _f, _f, _f, _f, _i, _vp]),
    "lg_augment_apply": (_i, [_vp, _vp, _vp, _vp, _f, _vp, _i, _i, _i, _i, _vp]),
    "lg_cast": (_i, [_vp, _vp, _i64, _i, _i, _vp]),
    "lg_scale_shift": (_i, [_vp, _vp, _i64, _f, _f, _i, _i, _vp]),
    "lg_normal_fill": (_i, [_vp, _i64, _vp, _vp]),
    "lg_u8_rescale": (_i, [_vp, _vp, _i64, _i, _vp]),
    "lg_conv2d_bn_relu": (_i, [_vp] * 6 + [_i] * 16 + [_vp]),
    "lg_pack_conv_bn_weights": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "lg_pool2d": (_i, [_vp, _vp] + [_i] * 13 + [_vp]),
    "lg_global_avgpool": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "lg_resize_bilinear_norm": (_i, [_vp, _vp] + [_i] * 7 + [_f, _f, _i, _i, _vp]),
    "lg_fid_accumulate": (_i, [_vp, _vp, _vp, _vp, _i64, _i, _vp]),
    "lg_fid_finalize": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _i, _vp]),
    "lg_dgemm": (_i, [_vp, _vp, _vp, _i, _d, _d, _vp]),
    "lg_dmat_stats": (_i, [_vp, _i, _vp, _vp]),
    "lg_dmat_scale_shift": (_i, [_vp, _vp, _i, _d, _d, _i, _vp]),
}



class BceItem(C.Structure):
    """lg_bce_item_t"""
    _fields_ = [("p", _vp), ("target", _vp), ("loss_accum", _vp), ("dlogit", _vp), ("target_const", _f),
                ("weight", _f), ("n", _i), ("pad_", _i)]


PACK_MAX, BCE_MAX = 12, 8


class NormBwd(C.Structure):
    """lg_norm_bwd_t: fused InstanceNorm-backward epilogue descriptor of a backward conv launch."""
    _fields_ = [("z", _vp), ("stats", _vp), ("gamma", _vp), ("beta", _vp), ("red", _vp), ("eps", _f),
                ("alpha", _f)]


_lib = None


def load():
    """Load the shared library (once) and bind every symbol the header declares."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "littlegan_b200: %s is missing - build it with `python -m littlegan_b200.csrc.build` "
            "(or __graft_entry__.build()); there is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class LittleGANError(RuntimeError):
    pass


def check(ret, what=""):
    if ret < 0:
        msg = load().lg_last_error().decode("utf-8", "replace")
        raise LittleGANError("%s failed (%d): %s" % (what or "littlegan_b200 call", ret, msg))
    return ret
