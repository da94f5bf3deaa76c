"""Per-layer timing of the conv-family kernels (CUDA events, L2 flushed between launches).
usage: python scripts/layer_bench.py [batch] [filter-substring] [cta-pair mode]"""
import sys
sys.path.insert(0, ".")
import torch
from littlegan_b200 import kernels as K

NB = int(sys.argv[1]) if len(sys.argv) > 1 else 128
flt = sys.argv[2] if len(sys.argv) > 2 else ""
if len(sys.argv) > 3:       # CTA-pair mode of the generic kernels: 0 never, 1 auto, 2 always
    K.set_cta_pairs(int(sys.argv[3]))
# name, Hb, A, B, stride   (big map Hb x Hb x A  <->  small map Hb/s x Hb/s x B)
LAYERS = [("enc1", 128, 3, 64, 2), ("enc2", 64, 64, 128, 2), ("enc3", 32, 128, 256, 2), ("enc4", 16, 256, 384, 2),
          ("dec4", 128, 32, 64, 2), ("final", 128, 3, 32, 1)]
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


for name, Hb, A, B, s in LAYERS:
    N = NB
    big = torch.randn(N, Hb, Hb, A, device="cuda").to(torch.bfloat16)
    small = torch.randn(N, Hb // s, Hb // s, B, device="cuda").to(torch.bfloat16)
    W = torch.randn(5, 5, A, B, device="cuda") * 0.05
    wp = torch.empty(K.pack_conv_weights_bytes(A, B), dtype=torch.uint8, device="cuda")
    K.pack_conv_weights(W, wp)
    stats = torch.zeros(N, 2, dtype=torch.float64, device="cuda")
    bA, bB = torch.zeros(A, device="cuda"), torch.zeros(B, device="cuda")
    dW = torch.zeros(5, 5, A, B, device="cuda")
    flops = 2.0 * 25 * A * B * N * (Hb // s) ** 2
    big16 = K.pad_channels(big, torch.empty(N, Hb, Hb, 16, dtype=torch.bfloat16, device="cuda")) if A < 16 else None
    runs = []
    if A >= 16:
        runs.append(("fprop", lambda: K.conv2d_fprop(big, W, bB, small, stats, s, wp, True)))
        runs.append(("wgrad", lambda: K.conv2d_wgrad(big, small, dW, s, True)))
    else:
        runs.append(("fprop(cin3)", lambda: K.conv2d_fprop(big, W, bB, small, stats, s, wp, True)))
        runs.append(("wgrad(cin3)", lambda: K.conv2d_wgrad(big, small, dW, s, True)))
        runs.append(("fprop(pad16)", lambda: K.conv2d_fprop(big16, W, bB, small, stats, s, wp, True)))
        runs.append(("wgrad(pad16)", lambda: K.conv2d_wgrad_padded(big16, small, dW, s)))
    runs.append(("dgrad", lambda: K.conv2d_dgrad(small, W, bA, big, stats, s, K.ACT_NONE, wp, True)))
    for op, fn in runs:
        if flt and flt not in name + op:
            continue
        us = timeit(fn)
        print("%-6s %-13s N=%3d  %8.1f us  %7.1f TFLOP/s" % (name, op, N, us, flops / us / 1e6), flush=True)
