"""One launch of the row-streaming fprop kernel for ncu.  usage: one_row.py final|enc1|dec4 [nb] [batch]"""
import sys
sys.path.insert(0, ".")
import torch
from littlegan_b200 import kernels as K

name = sys.argv[1]
with_nb = len(sys.argv) > 2 and sys.argv[2] == "nb"
N = int(sys.argv[3]) if len(sys.argv) > 3 else 128
A_big, A, B, s = {"final": (8, 3, 32, 1), "enc1": (8, 3, 64, 2), "dec4": (32, 32, 64, 2)}[name]
Hb = 128
xp = torch.zeros(N, Hb, Hb, A_big, device="cuda", dtype=torch.bfloat16)
xp[..., :A] = torch.randn(N, Hb, Hb, A, device="cuda").to(torch.bfloat16)
W = torch.randn(5, 5, A, B, device="cuda") * 0.05
wr = K.pack_rowconv_weights(W, A_big, s)
out = torch.empty(N, Hb // s, Hb // s, B, device="cuda", dtype=torch.bfloat16)
bias = torch.zeros(B, device="cuda")
stats = torch.zeros(N, 2, dtype=torch.float64, device="cuda")
z = torch.randn_like(out)
K.rowstats(z, stats, 1.0)
red = torch.zeros(N, 2, dtype=torch.float64, device="cuda")
nb = K.norm_bwd_desc(z, stats, torch.ones(1, device="cuda"), torch.zeros(1, device="cuda"), red, 1e-3, 0.3)
for _ in range(3):
    if with_nb:
        K.conv2d_fprop_rows(xp, wr, None, out, None, s, A, norm_bwd=nb)
    else:
        K.conv2d_fprop_rows(xp, wr, bias, out, stats, s, A)
torch.cuda.synchronize()
print("ok")
