"""One launch set of the row-streaming dec4 wgrad kernel for ncu.  usage: one_wgrad.py [batch]"""
import sys
sys.path.insert(0, ".")
import torch
from littlegan_b200 import kernels as K
N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
big = torch.randn(N, 128, 128, 32, device="cuda").to(torch.bfloat16)
small = torch.randn(N, 64, 64, 64, device="cuda").to(torch.bfloat16)
dW = torch.zeros(5, 5, 32, 64, device="cuda")
for _ in range(3):
    K.conv2d_wgrad(big, small, dW, 2, True)
torch.cuda.synchronize()
print("ok")
