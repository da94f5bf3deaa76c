import sys; sys.path.insert(0, ".")
import torch
from oracle import littlegan_oracle as O
from littlegan_b200 import kernels as K
from tests.util import rel_err
torch.manual_seed(0)
for (N,Hb,Wb,A,B,s) in [(2,64,64,64,128,2),(2,32,32,128,256,2),(2,128,128,32,64,2)]:
    x = torch.randn(N,Hb,Wb,A); W = torch.randn(5,5,A,B)*0.05
    ref = O.conv2d_same(x.double(), W.double(), torch.zeros(B,dtype=torch.float64), s)
    out = torch.empty(N,Hb//s,Wb//s,B, device="cuda")
    K.conv2d_fprop(x.cuda(), W.cuda(), None, out, None, s)
    print("fprop", (N,Hb,Wb,A,B,s), rel_err(out, ref))
    # with structured (non-random) input: smooth gradient-like data
    x2 = x * 1e-3 + 5.0
    ref = O.conv2d_same(x2.double(), W.double(), torch.zeros(B,dtype=torch.float64), s)
    K.conv2d_fprop(x2.cuda(), W.cuda(), None, out, None, s)
    print("  offset input", rel_err(out, ref))
# IN backward at large M with near-cancelling inputs
for M in [131072, 524288]:
    N = 2
    z = torch.randn(N, M) * 0.3 + 2.0
    g = torch.randn(N, M) * 1e-3 + 0.05
    gamma = torch.tensor([1.0]); beta = torch.tensor([0.0])
    zr = z.double().requires_grad_(True)
    y = torch.nn.functional.leaky_relu(O.instance_norm(zr, gamma.double(), beta.double()), 0.3)
    (dref,) = torch.autograd.grad((y * g.double()).sum(), zr)
    stats = torch.zeros(N,2,dtype=torch.float64,device="cuda"); K.rowstats(z.cuda(), stats, 1.0)
    red = torch.zeros(N,2,dtype=torch.float64,device="cuda"); dz = torch.empty(N,M,device="cuda")
    K.instnorm_act_bwd(g.cuda(), z.cuda(), stats, gamma.cuda(), beta.cuda(), red, dz, None, None, 1e-3, 1.0, 0.3)
    print("in_bwd", M, rel_err(dz, dref), float(dref.abs().max()), float(g.abs().max()))
