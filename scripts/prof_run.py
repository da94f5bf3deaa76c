import sys; sys.path.insert(0, "."); import os
import torch
from littlegan_b200 import kernels as K
for name, Hb, A, B, s in [("enc2", 64, 64, 128, 2), ("enc3", 32, 128, 256, 2), ("enc4", 16, 256, 384, 2), ("dec4", 128, 32, 64, 2)]:
    N = 128
    big = torch.randn(N, Hb, Hb, A, device="cuda").to(torch.bfloat16)
    small = torch.randn(N, Hb // s, Hb // s, B, device="cuda").to(torch.bfloat16)
    W = torch.randn(5, 5, A, B, device="cuda") * 0.05
    wp = torch.empty(K.pack_conv_weights_bytes(A, B), dtype=torch.uint8, device="cuda"); K.pack_conv_weights(W, wp)
    stats = torch.zeros(N, 2, dtype=torch.float64, device="cuda")
    bA, bB = torch.zeros(A, device="cuda"), torch.zeros(B, device="cuda")
    for rep in range(2):
        print(name, "fprop", flush=True); K.conv2d_fprop(big, W, bB, small, stats, s, wp, True); torch.cuda.synchronize()
    for rep in range(2):
        print(name, "dgrad", flush=True); K.conv2d_dgrad(small, W, bA, big, stats, s, K.ACT_NONE, wp, True); torch.cuda.synchronize()
