"""Launches the bench's roofline kernel (tcgen05 dgrad, dec2 geometry, batch 128) a few times -
the command captured with `ncu --set full` for the per-launch DRAM traffic."""
import sys
sys.path.insert(0, ".")
import torch
from littlegan_b200 import kernels as K
which = sys.argv[1] if len(sys.argv) > 1 else "dgrad"
N, Hb, Wb, A, B, s = 128, 32, 32, 128, 256, 2
W = torch.randn(5, 5, A, B, device="cuda") * 0.05
wp = torch.empty(K.pack_conv_weights_bytes(A, B), dtype=torch.uint8, device="cuda")
K.pack_conv_weights(W, wp)
small = torch.randn(N, Hb // s, Wb // s, B, device="cuda").to(torch.bfloat16)
big = torch.randn(N, Hb, Wb, A, device="cuda").to(torch.bfloat16)
stats = torch.zeros(N, 2, dtype=torch.float64, device="cuda")
bias_a = torch.zeros(A, device="cuda"); bias_b = torch.zeros(B, device="cuda")
dW = torch.zeros(5, 5, A, B, device="cuda")
for _ in range(5):
    if which == "dgrad":
        K.conv2d_dgrad(small, W, bias_a, big, stats, s, K.ACT_NONE, wp, True)
    elif which == "fprop":
        K.conv2d_fprop(big, W, bias_b, small, stats, s, wp, True)
    else:
        K.conv2d_wgrad(big, small, dW, s, True)
torch.cuda.synchronize()
print("ok")
