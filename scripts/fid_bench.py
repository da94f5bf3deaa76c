"""FID statistics pass (config 5): N x 2048 fp32 features -> mean / covariance (fp64) on the GPU,
timed against np.mean/np.cov on the host; also the Frechet distance."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from littlegan_b200 import fid
from oracle import fid_oracle
N = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
d = 2048
g = torch.Generator(device="cuda").manual_seed(0)
X = torch.randn(N, d, device="cuda", generator=g) * 0.5 + 0.3
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    mu, sigma = fid.calculate_activation_statistics(X, None, batch_size=100, as_numpy=False)
    torch.cuda.synchronize(); t1 = time.perf_counter()
print("gpu stats: %.1f ms for N=%d (%.1f GFLOP fp64 SYRK-equivalent -> %.2f TFLOP/s fp64)" % (
    (t1 - t0) * 1e3, N, N * d * d / 1e9, N * d * d / (t1 - t0) / 1e12))
Xs = X[:10000].cpu().numpy()
t0 = time.perf_counter(); mu_r, sig_r = fid_oracle.activation_statistics(Xs); t1 = time.perf_counter()
print("numpy np.mean/np.cov on 10000 rows: %.1f ms (host)" % ((t1 - t0) * 1e3))
mu2, sig2 = fid.calculate_activation_statistics(X[:10000], None, batch_size=100)
print("parity vs numpy (10000 rows): mu %.2e sigma %.2e" % (np.abs(mu2 - mu_r).max(), np.abs(sig2 - sig_r).max() / np.abs(sig_r).max()))
t0 = time.perf_counter(); f = fid.calculate_frechet_distance(mu, sigma, mu2, sig2); torch.cuda.synchronize(); t1 = time.perf_counter()
print("frechet distance (2048x2048 fp64, GPU eigh): %.4f in %.1f ms" % (f, (t1 - t0) * 1e3))
