"""Accuracy (vs scipy.linalg.sqrtm, the reference's arithmetic) and time of the Newton-Schulz Frechet distance."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from littlegan_b200 import fid, kernels as K
from oracle import fid_oracle
rng = np.random.RandomState(0)
# dgemm alone
n = 2048
A = torch.randn(n, n, dtype=torch.float64, device="cuda"); B = torch.randn(n, n, dtype=torch.float64, device="cuda")
C = torch.empty_like(A)
K.dgemm(A, B, C, 0.5, 2.0); torch.cuda.synchronize()
ref = 0.5 * (A @ B) + 2.0 * torch.eye(n, dtype=torch.float64, device="cuda")
print("dgemm 2048 max rel err %.2e" % float((C - ref).abs().max() / ref.abs().max()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): K.dgemm(A, B, C)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("dgemm 2048^3: %.3f ms = %.1f TFLOP/s fp64" % (ms, 2 * n ** 3 / ms / 1e9))
for m in (37, 200):
    a = torch.randn(m, m, dtype=torch.float64, device="cuda"); b = torch.randn(m, m, dtype=torch.float64, device="cuda"); c = torch.empty_like(a)
    K.dgemm(a, b, c, 1.0, 0.0); print("dgemm %d max err %.2e" % (m, float((c - a @ b).abs().max())))
for name, d, n1, n2 in (("full rank d=2048", 2048, 6000, 5000), ("singular d=512 (n < d)", 512, 60, 90), ("small d=24", 24, 400, 300),
                        ("odd d=200", 200, 1000, 150)):
    x1 = (rng.randn(n1, d) @ (rng.randn(d, d) / np.sqrt(d))) * 0.5 + 0.3
    x2 = (rng.randn(n2, d) @ (rng.randn(d, d) / np.sqrt(d))) * 0.7 + 0.1
    m1, s1 = fid_oracle.activation_statistics(x1); m2, s2 = fid_oracle.activation_statistics(x2)
    t = time.perf_counter(); ref = fid_oracle.frechet_distance(m1, s1, m2, s2); t_ref = time.perf_counter() - t
    got = fid.calculate_frechet_distance(m1, s1, m2, s2); torch.cuda.synchronize()
    t = time.perf_counter(); got = fid.calculate_frechet_distance(m1, s1, m2, s2); torch.cuda.synchronize(); t_got = time.perf_counter() - t
    _, tr, it = fid._sqrtm_psd(torch.from_numpy(s1).cuda(), want_matrix=False)
    print("%-24s scipy %.10g (%.2f s)  ours %.10g (%.3f s)  rel diff %.2e   [sqrt(sigma1): %d iterations]" % (
        name, ref, t_ref, got, t_got, abs(got - ref) / abs(ref), it))
