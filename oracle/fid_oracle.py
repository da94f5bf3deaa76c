"""FID statistics / Frechet distance oracle (NumPy / SciPy, fp64).

TEST INFRASTRUCTURE - never imported by `littlegan_b200/`.  Unlike the train
step, this row is PINNED: NumPy and SciPy are installed, so these functions
are the reference's own arithmetic (`fid.py:186-187` np.mean / np.cov and
`fid.py:112-163` scipy.linalg.sqrtm), restated without the TF session.
"""
import warnings

import numpy as np
from scipy import linalg


def activation_statistics(act):
    """fid.py:185-188 on a precomputed activation matrix [N, d] (fp64)."""
    act = np.asarray(act, dtype=np.float64)
    mu = np.mean(act, axis=0)
    sigma = np.cov(act, rowvar=False)
    return mu, sigma


def used_rows(n_images, batch_size):
    """fid.py:89-94: batch_size is clamped to N and the N % batch tail is dropped."""
    if batch_size > n_images:
        batch_size = n_images
    return (n_images // batch_size) * batch_size


def frechet_distance(mu1, sigma1, mu2, sigma2, eps=1e-6):
    """fid.py:112-163."""
    mu1 = np.atleast_1d(mu1)
    mu2 = np.atleast_1d(mu2)
    sigma1 = np.atleast_2d(sigma1)
    sigma2 = np.atleast_2d(sigma2)
    assert mu1.shape == mu2.shape, "Training and test mean vectors have different lengths"
    assert sigma1.shape == sigma2.shape, "Training and test covariances have different dimensions"
    diff = mu1 - mu2
    cov_mean = linalg.sqrtm(sigma1.dot(sigma2))
    if isinstance(cov_mean, tuple):
        cov_mean = cov_mean[0]
    if not np.isfinite(cov_mean).all():
        warnings.warn("fid calculation produces singular product; adding %s to diagonal of cov estimates" % eps)
        offset = np.eye(sigma1.shape[0]) * eps
        cov_mean = linalg.sqrtm((sigma1 + offset).dot(sigma2 + offset))
    if np.iscomplexobj(cov_mean):
        if not np.allclose(np.diagonal(cov_mean).imag, 0, atol=1e-3):
            raise ValueError("Imaginary component {}".format(np.max(np.abs(cov_mean.imag))))
        cov_mean = cov_mean.real
    return diff.dot(diff) + np.trace(sigma1) + np.trace(sigma2) - 2 * np.trace(cov_mean)
