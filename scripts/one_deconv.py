"""One launch of the row-streaming RGB transposed conv for ncu.  usage: one_deconv.py [batch]"""
import sys
sys.path.insert(0, ".")
import torch
from littlegan_b200 import kernels as K
N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
x = torch.randn(N, 128, 128, 32, device="cuda").to(torch.bfloat16)
W = torch.randn(5, 5, 3, 32, device="cuda") * 0.05
out = torch.empty(N, 128, 128, 3, device="cuda", dtype=torch.bfloat16)
bias = torch.zeros(3, device="cuda")
for _ in range(3):
    K.conv2d_dgrad_rgb(x, W, bias, out, None, None, 1, K.ACT_TANH)
torch.cuda.synchronize()
print("ok")
