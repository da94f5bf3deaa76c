"""Timing of the row-streaming fprop kernel against the kernels it replaces (CUDA events, L2 flushed).
usage: python scripts/row_bench.py [batch]"""
import sys
sys.path.insert(0, ".")
import torch
from littlegan_b200 import kernels as K

NB = int(sys.argv[1]) if len(sys.argv) > 1 else 128
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def timeit(fn, reps=5):
    tot = 0.0
    for i in range(2 + reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        if i >= 2:
            tot += e0.elapsed_time(e1)
    return tot / reps * 1e3


# name, A_big, A, B, stride
for name, A_big, A, B, s in [("final-dgrad", 8, 3, 32, 1), ("enc1-fwd", 8, 3, 64, 2), ("dec4-dgrad", 32, 32, 64, 2)]:
    N, Hb = NB, 128
    x = torch.randn(N, Hb, Hb, A, device="cuda").to(torch.bfloat16)
    xp = torch.zeros(N, Hb, Hb, A_big, device="cuda", dtype=torch.bfloat16)
    xp[..., :A] = x
    W = torch.randn(5, 5, A, B, device="cuda") * 0.05
    wr = K.pack_rowconv_weights(W, A_big, s)
    wp = torch.empty(K.pack_conv_weights_bytes(A, B), dtype=torch.uint8, device="cuda")
    K.pack_conv_weights(W, wp)
    out = torch.empty(N, Hb // s, Hb // s, B, device="cuda", dtype=torch.bfloat16)
    bias = torch.zeros(B, device="cuda")
    stats = torch.zeros(N, 2, dtype=torch.float64, device="cuda")
    z = torch.randn_like(out)
    K.rowstats(z, stats, 1.0)
    red = torch.zeros(N, 2, dtype=torch.float64, device="cuda")
    nb = K.norm_bwd_desc(z, stats, torch.ones(1, device="cuda"), torch.zeros(1, device="cuda"), red, 1e-3, 0.3)
    flops = 2.0 * 25 * A * B * N * (Hb // s) ** 2
    gb = (xp.numel() + out.numel()) * 2 / 1e9
    for label, fn in [
        ("rows", lambda: K.conv2d_fprop_rows(xp, wr, bias, out, stats, s, A)),
        ("rows+nb", lambda: K.conv2d_fprop_rows(xp, wr, None, out, None, s, A, norm_bwd=nb)),
        ("old", lambda: K.conv2d_fprop(x, W, bias, out, stats, s, wp, True)),
        ("old+nb", lambda: K.conv2d_fprop(x, W, None, out, None, s, wp, True, norm_bwd=nb)),
    ]:
        us = timeit(fn)
        extra = out.numel() * 2 / 1e9 if "nb" in label else 0.0
        print("%-12s %-8s N=%3d %8.1f us %7.1f TFLOP/s %7.0f GB/s" % (name, label, N, us, flops / us / 1e6,
                                                                    (gb + extra) / us * 1e6), flush=True)

# final Conv2DTranspose(3) forward: row-streaming kernel vs GEMM + col2im
N, Hb = NB, 128
x = torch.randn(N, Hb, Hb, 32, device="cuda").to(torch.bfloat16)
W = torch.randn(5, 5, 3, 32, device="cuda") * 0.05
wp = torch.empty(K.pack_conv_weights_bytes(3, 32), dtype=torch.uint8, device="cuda")
K.pack_conv_weights(W, wp)
out = torch.empty(N, Hb, Hb, 3, device="cuda", dtype=torch.bfloat16)
out8 = torch.empty(N, Hb, Hb, 8, device="cuda", dtype=torch.bfloat16)
bias = torch.zeros(3, device="cuda")
for label, fn in [("rows", lambda: K.conv2d_dgrad_rgb(x, W, bias, out, None, None, 1, K.ACT_TANH)),
                  ("rows+pad8", lambda: K.conv2d_dgrad_rgb(x, W, bias, out, out8, None, 1, K.ACT_TANH))]:
    us = timeit(fn)
    print("%-12s %-8s N=%3d %8.1f us %7.0f GB/s" % ("final-fwd", label, N, us, (x.numel() + out.numel()) * 2 / 1e3 / us), flush=True)

# decoder conv4 forward: row-streaming dgrad vs the four-phase implicit-GEMM kernel
x = torch.randn(N, 64, 64, 64, device="cuda").to(torch.bfloat16)
W = torch.randn(5, 5, 32, 64, device="cuda") * 0.05
wp = torch.empty(K.pack_conv_weights_bytes(32, 64), dtype=torch.uint8, device="cuda")
K.pack_conv_weights(W, wp)
wd = K.pack_rowdgrad_weights(W)
out = torch.empty(N, 128, 128, 32, device="cuda", dtype=torch.bfloat16)
bias = torch.zeros(32, device="cuda")
stats = torch.zeros(N, 2, dtype=torch.float64, device="cuda")
flops = 2.0 * 25 * 32 * 64 * N * 64 * 64
for label, fn in [("rows", lambda: K.conv2d_dgrad_rows(x, wd, bias, out, stats, 2)),
                  ("old", lambda: K.conv2d_dgrad(x, W, bias, out, stats, 2, K.ACT_NONE, wp, True))]:
    us = timeit(fn)
    print("%-12s %-8s N=%3d %8.1f us %7.1f TFLOP/s %7.0f GB/s" % ("dec4-fwd", label, N, us, flops / us / 1e6,
                                                                (x.numel() + out.numel()) * 2 / 1e3 / us), flush=True)

# encoder conv1 input gradient (64 -> 3 channels, stride 2): row-streaming kernel (routed through lg_conv2d_dgrad)
x = torch.randn(N, 64, 64, 64, device="cuda").to(torch.bfloat16)
W = torch.randn(5, 5, 3, 64, device="cuda") * 0.05
out = torch.empty(N, 128, 128, 3, device="cuda", dtype=torch.bfloat16)
us = timeit(lambda: K.conv2d_dgrad_rgb(x, W, None, out, None, None, 2, K.ACT_NONE))
print("%-12s %-8s N=%3d %8.1f us %7.0f GB/s" % ("enc1-dgrad", "rows", N, us, (x.numel() + out.numel()) * 2 / 1e3 / us), flush=True)

# decoder conv4 weight gradient: row-streaming kernel (routed through lg_conv2d_wgrad)
big = torch.randn(N, 128, 128, 32, device="cuda").to(torch.bfloat16)
small = torch.randn(N, 64, 64, 64, device="cuda").to(torch.bfloat16)
dW = torch.zeros(5, 5, 32, 64, device="cuda")
flops = 2.0 * 25 * 32 * 64 * N * 64 * 64
us = timeit(lambda: K.conv2d_wgrad(big, small, dW, 2, True))
print("%-12s %-8s N=%3d %8.1f us %7.1f TFLOP/s %7.0f GB/s" % ("dec4-wgrad", "rows", N, us, flops / us / 1e6,
                                                            (big.numel() + small.numel()) * 2 / 1e3 / us), flush=True)
