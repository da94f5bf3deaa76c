import sys; sys.path.insert(0, ".")
import torch
from oracle import littlegan_oracle as O
from littlegan_b200 import kernels as K
from tests.util import rel_err
torch.manual_seed(0)
for M in [131072, 262144, 524288]:
    N = 2
    z = torch.randn(N, M) * 0.3 + 2.0
    g = torch.randn(N, M) * 1e-3 + 0.05
    gamma = torch.tensor([1.0]); beta = torch.tensor([0.0])
    zd = z.double()
    mu = zd.mean(1, keepdim=True); sig = ((zd-mu)**2).mean(1, keepdim=True).sqrt(); s = sig + 1e-3
    xh = (zd-mu)/s; y = xh
    dy = g.double() * torch.where(y > 0, 1.0, 0.3)
    zr = zd.clone().requires_grad_(True)
    yy = torch.nn.functional.leaky_relu(O.instance_norm(zr, gamma.double(), beta.double()), 0.3)
    (dref,) = torch.autograd.grad((yy * g.double()).sum(), zr)
    stats = torch.zeros(N,2,dtype=torch.float64,device="cuda"); K.rowstats(z.cuda(), stats, 1.0)
    print(M, "stats", stats.cpu().tolist(), "ref", zd.sum(1).tolist(), (zd**2).sum(1).tolist())
    red = torch.zeros(N,2,dtype=torch.float64,device="cuda"); dz = torch.empty(N,M,device="cuda")
    K.instnorm_act_bwd(g.cuda(), z.cuda(), stats, gamma.cuda(), beta.cuda(), red, dz, None, None, 1e-3, 1.0, 0.3)
    print("  red", red.cpu().tolist(), "ref", dy.sum(1).tolist(), (dy*xh).sum(1).tolist())
    err = (dz.cpu().double()-dref).abs()
    print("  err max", float(err.max()), "at", int(err.argmax()) % M, "frac bad", float((err > 1e-4).double().mean()))
    bad = (err > 1e-4).nonzero()
    if len(bad): print("  first bad", bad[:3].tolist(), "last bad", bad[-3:].tolist())
