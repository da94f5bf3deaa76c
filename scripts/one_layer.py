"""Launch one conv-family kernel a few times (for ncu): python scripts/one_layer.py <layer> <op> [batch]"""
import sys
sys.path.insert(0, ".")
import torch
from littlegan_b200 import kernels as K
LAYERS = {"enc1": (128, 3, 64, 2), "enc2": (64, 64, 128, 2), "enc3": (32, 128, 256, 2), "enc4": (16, 256, 384, 2),
          "dec4": (128, 32, 64, 2), "final": (128, 3, 32, 1)}
name, op = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 128
Hb, A, B, s = LAYERS[name]
big = torch.randn(N, Hb, Hb, A, device="cuda").to(torch.bfloat16)
small = torch.randn(N, Hb // s, Hb // s, B, device="cuda").to(torch.bfloat16)
W = torch.randn(5, 5, A, B, device="cuda") * 0.05
wp = torch.empty(K.pack_conv_weights_bytes(A, B), dtype=torch.uint8, device="cuda"); K.pack_conv_weights(W, wp)
stats = torch.zeros(N, 2, dtype=torch.float64, device="cuda")
bA, bB = torch.zeros(A, device="cuda"), torch.zeros(B, device="cuda")
dW = torch.zeros(5, 5, A, B, device="cuda")
for _ in range(4):
    if op == "fprop": K.conv2d_fprop(big, W, bB, small, stats, s, wp, True)
    elif op == "dgrad": K.conv2d_dgrad(small, W, bA, big, stats, s, K.ACT_NONE, wp, True)
    else: K.conv2d_wgrad(big, small, dW, s, True)
torch.cuda.synchronize(); print("ok")
