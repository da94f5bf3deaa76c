"""Launch the fused discriminator-heads kernels and the final-conv backward (norm-backward epilogue) a few
times at the bench sizes (for ncu): python scripts/prof_heads.py [batch]"""
import sys
sys.path.insert(0, ".")
import torch
from littlegan_b200 import kernels as K
N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
F, U1 = 24576, 40
f = torch.randn(N, F, device="cuda").to(torch.bfloat16)
W0 = torch.randn(F, 1, device="cuda") * 0.01; W1 = torch.randn(F, U1, device="cuda") * 0.01
b0 = torch.zeros(1, device="cuda"); b1 = torch.zeros(U1, device="cuda")
o0 = torch.empty(N, 1, device="cuda"); o1 = torch.empty(N, U1, device="cuda")
dl0 = torch.randn(N, 1, device="cuda"); dl1 = torch.randn(N, U1, device="cuda")
df = torch.empty_like(f)
dW0 = torch.zeros_like(W0); dW1 = torch.zeros_like(W1); db0 = torch.zeros(1, device="cuda"); db1 = torch.zeros(U1, device="cuda")
ws = K.dense_heads_workspace(N, "cuda")
# final conv backward with the fused epilogue
Hb, A, B = 128, 3, 32
dpre = torch.randn(N, Hb, Hb, A, device="cuda").to(torch.bfloat16)
z = torch.randn(N, Hb, Hb, B, device="cuda").to(torch.bfloat16)
g = torch.empty_like(z)
Wc = torch.randn(5, 5, A, B, device="cuda") * 0.05
wp = torch.empty(K.pack_conv_weights_bytes(A, B), dtype=torch.uint8, device="cuda"); K.pack_conv_weights(Wc, wp)
stats = torch.zeros(N, 2, dtype=torch.float64, device="cuda"); K.rowstats(z, stats, 1.0)
red = torch.zeros(N, 2, dtype=torch.float64, device="cuda")
gamma = torch.ones(1, device="cuda"); beta = torch.zeros(1, device="cuda")
nb = K.norm_bwd_desc(z, stats, gamma, beta, red, 1e-3, 0.3)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
for rep in range(3):
    ev[0].record()
    K.dense_heads_fwd(f, W0, b0, W1, b1, o0, o1, K.ACT_SIGMOID, ws)
    ev[1].record()
    K.dense_heads_bwd(f, dl0, dl1, W0, W1, df, dW0, dW1, db0, db1)
    ev[2].record()
    K.conv2d_fprop(dpre, Wc, None, g, None, 1, wp, True, norm_bwd=nb)
    ev[3].record()
    K.conv2d_fprop(dpre, Wc, None, g, None, 1, wp, True)
    ev[4].record()
    torch.cuda.synchronize()
    print("heads fwd %.1f us  bwd %.1f us  cin3 fprop+nb %.1f us  plain %.1f us" % tuple(
        ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(4)), flush=True)
