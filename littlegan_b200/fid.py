"""FID feature statistics and Frechet distance (reference `fid.py:73-188`).

`calculate_activation_statistics` keeps the reference's signature; the statistics pass is a
streaming fp64 accumulation of sum(x) and sum(x x^T) on the GPU (csrc/fid.cu) instead of
materialising the [N, 2048] fp64 activation matrix and calling np.mean / np.cov
(`fid.py:95,186-187`).  With torch.distributed initialised every rank accumulates its shard and
the (n, S1, S2) triple is all-reduced once (SURVEY 8 e).

The Inception-2015 pool_3 forward itself (`fid.py:36-106`) is `littlegan_b200.inception.InceptionPool3`: an instance
is the `sess` argument here (any callable mapping an image batch [b,H,W,3] in 0..255 to a CUDA fp32 feature matrix
[b,d] works); pass `sess=None` with a 2-D input to feed precomputed activations.  The file-based variants
(`fid.py:197-318`: `*_from_files`, `_handle_path`, `calculate_fid_given_paths`) decode on host threads one batch ahead
of the GPU and hand the forward image BYTES.
"""
import math
import warnings

import numpy as np
import torch

from . import kernels as K


class InvalidFIDException(Exception):
    pass


class FeatureStatistics:
    """Streaming accumulator: update(X[b,d] fp32 cuda) ... finalize() -> (mu, sigma) fp64."""

    def __init__(self, d, device=None, shift=None):
        dev = device or torch.device("cuda")
        self.d, self.n = d, 0
        self.S1 = torch.zeros(d, dtype=torch.float64, device=dev)
        self.S2 = torch.zeros(d, d, dtype=torch.float64, device=dev)
        self.shift = None if shift is None else shift.to(dev, torch.float64).contiguous()
        self._stage = None            # batches are pooled to >= 4096 rows per kernel launch
        self._fill = 0

    def update(self, X):
        """Queue a feature batch [b, d]; the outer-product kernel runs once per ~4096 rows (a launch per
        100-row Inception batch, `evaluate.py:54`, would leave the GPU mostly idle)."""
        if X.dim() != 2 or X.shape[1] != self.d:
            raise ValueError("expected features of shape [b, %d]" % self.d)
        X = X.to(torch.float32)
        if X.shape[0] >= 4096 and self._fill == 0:
            return self._accumulate(X.contiguous())
        if self._stage is None:
            self._stage = torch.empty(8192, self.d, dtype=torch.float32, device=self.S1.device)
        if self._fill + X.shape[0] > self._stage.shape[0]:
            self.flush()
        self._stage[self._fill:self._fill + X.shape[0]].copy_(X)
        self._fill += X.shape[0]
        if self._fill >= 4096:
            self.flush()

    def flush(self):
        if self._fill:
            self._accumulate(self._stage[:self._fill])
            self._fill = 0

    def _accumulate(self, X):
        if X.dim() != 2 or X.shape[1] != self.d:
            raise ValueError("expected features of shape [b, %d]" % self.d)
        X = X.to(torch.float32).contiguous()
        if self.shift is None:
            # centre on the first batch's mean: removes the cancellation in S2 - n mu mu^T
            self.shift = X.double().mean(dim=0).contiguous()
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                dist.broadcast(self.shift, src=0)
        K.fid_accumulate(X, self.shift, self.S1, self.S2)
        self.n += X.shape[0]

    def finalize(self):
        import torch.distributed as dist
        self.flush()
        n = self.n
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            cnt = torch.tensor([n], dtype=torch.float64, device=self.S1.device)
            dist.all_reduce(cnt)
            dist.all_reduce(self.S1)
            dist.all_reduce(self.S2)
            n = int(cnt.item())
        if n < 2:
            raise InvalidFIDException("need at least two samples")
        mu = torch.empty_like(self.S1)
        sigma = torch.empty_like(self.S2)
        K.fid_finalize(self.S1, self.S2, self.shift, mu, sigma, n)
        return mu, sigma


def get_activations(images, sess, batch_size=50, verbose=False):
    """fid.py:73-106 as a generator of CUDA feature batches; the N % batch tail is dropped as in
    the reference (`fid.py:89-94`)."""
    d0 = images.shape[0]
    if batch_size > d0:
        print("warning: batch size is bigger than the data size. setting batch size to data size")
        batch_size = d0
    n_batches = d0 // batch_size
    for i in range(n_batches):
        if verbose:
            print("\rPropagating batch %d/%d" % (i + 1, n_batches), end="", flush=True)
        batch = images[i * batch_size:(i + 1) * batch_size]
        if sess is None:
            feats = batch
        else:
            feats = sess(batch)
        if isinstance(feats, np.ndarray):
            feats = torch.from_numpy(feats)
        yield feats.reshape(batch_size, -1).to("cuda", torch.float32, non_blocking=True)
    if verbose:
        print(" done")


def calculate_activation_statistics(images, sess=None, batch_size=50, verbose=False, as_numpy=True):
    """fid.py:169-188 -> (mu [d], sigma [d,d]) in fp64."""
    acc = None
    for feats in get_activations(images, sess, batch_size, verbose):
        if acc is None:
            acc = FeatureStatistics(feats.shape[1], feats.device)
        acc.update(feats)
    mu, sigma = acc.finalize()
    if as_numpy:
        return mu.cpu().numpy(), sigma.cpu().numpy()
    return mu, sigma


_NS_MAX_ITERS = 60


def _sqrtm_psd(A, want_matrix=True, tol=1e-14):
    """Square root of a symmetric PSD fp64 matrix A [n,n] (CUDA) by the coupled Newton-Schulz iteration
        Y0 = A / c, Z0 = I;  T = (3 I - Z Y) / 2;  Y <- Y T, Z <- T Z;   Y -> (A / c)^(1/2),  c = ||A||_F >= lambda_max
    on this library's fp64 GEMM (csrc/fid.cu `lg_dgemm`): no eigen-solver, no library call.
    Returns (sqrt(A) or None, Tr sqrt(A), iterations).

    Each eigen-component converges monotonically from below - quadratically once lambda_k/c has grown to O(1),
    after ~log(c / lambda) / log(2.25) iterations - so Tr Y increases until every component above the rounding
    level has arrived.  Numerically-zero eigenvalues of a singular A are rounding noise of either sign; a negative
    one grows like -|noise| * 1.5^k, so the iteration stops as soon as the trace stops increasing by more than
    `tol` (relative) - long before such a component matters - and never runs past _NS_MAX_ITERS."""
    n = A.shape[0]
    dev = A.device
    stats = torch.empty(3, dtype=torch.float64, device=dev)
    K.dmat_stats(A, stats)
    tr_a, fro2, _ = stats.tolist()
    if not (math.isfinite(tr_a) and math.isfinite(fro2)):
        return None, float("nan"), 0
    c = math.sqrt(fro2)
    if c == 0.0:
        return (torch.zeros_like(A) if want_matrix else None), 0.0, 0
    Y, Z, T, Yn, Zn = (torch.empty_like(A) for _ in range(5))
    K.dmat_scale_shift(A, Y, 1.0 / c, 0.0)
    K.dmat_scale_shift(None, Z, 0.0, 1.0)
    prev, best, iters = None, None, 0
    for k in range(_NS_MAX_ITERS):
        K.dgemm(Z, Y, T, -0.5, 1.5)
        K.dgemm(Y, T, Yn)
        K.dgemm(T, Z, Zn)
        Y, Yn, Z, Zn = Yn, Y, Zn, Z
        K.dmat_stats(Y, stats)
        t = float(stats[0])
        iters = k + 1
        if not math.isfinite(t):
            return None, float("nan"), iters
        if prev is not None and t - prev <= tol * abs(t):
            if t < prev:                       # a rounding-noise component has started to grow: keep the peak
                Y, t = Yn, prev                # (Yn holds the previous iterate)
            best = t
            break
        prev = best = t
    root = math.sqrt(c)
    if want_matrix:
        K.dmat_scale_shift(Y, T, root, 0.0, symmetrise=True)
        return T, best * root, iters
    return None, best * root, iters


def _trace_sqrt_product(s1, s2):
    """Tr sqrtm(s1 s2) for symmetric PSD s1, s2 (fp64, CUDA): the spectrum of s1 s2 is that of the symmetric PSD
    matrix R s2 R with R = s1^(1/2); both square roots by Newton-Schulz on the library's own fp64 GEMM."""
    R, _, _ = _sqrtm_psd(s1.contiguous(), want_matrix=True)
    if R is None:
        return torch.tensor(float("nan"), dtype=torch.float64, device=s1.device)
    tmp, M = torch.empty_like(R), torch.empty_like(R)
    K.dgemm(R, s2.contiguous(), tmp)
    K.dgemm(tmp, R, M)
    Ms = torch.empty_like(M)
    K.dmat_scale_shift(M, Ms, 1.0, 0.0, symmetrise=True)
    _, tr, _ = _sqrtm_psd(Ms, want_matrix=False)
    return torch.tensor(tr, dtype=torch.float64, device=s1.device)


def calculate_frechet_distance(mu1, sigma1, mu2, sigma2, eps=1e-6):
    """fid.py:112-163: ||mu1-mu2||^2 + Tr s1 + Tr s2 - 2 Tr sqrtm(s1 s2), fp64 on the GPU.
    The reference's scipy.linalg.sqrtm (Schur) is replaced by the symmetric Newton-Schulz form above (hand-written
    fp64 GEMM kernels, no cuSOLVER); the non-finite fallback (add eps*I to both covariances, `fid.py:148-152`) is
    kept."""
    dev = torch.device("cuda")
    t = lambda x: torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x).to(dev, torch.float64)
    mu1, mu2 = torch.atleast_1d(t(mu1)), torch.atleast_1d(t(mu2))
    sigma1, sigma2 = torch.atleast_2d(t(sigma1)), torch.atleast_2d(t(sigma2))
    assert mu1.shape == mu2.shape, "Training and test mean vectors have different lengths"
    assert sigma1.shape == sigma2.shape, "Training and test covariances have different dimensions"
    diff = mu1 - mu2
    tr = _trace_sqrt_product(sigma1, sigma2)
    if not torch.isfinite(tr):
        warnings.warn("fid calculation produces singular product; adding %s to diagonal of cov estimates" % eps)
        offset = torch.eye(sigma1.shape[0], dtype=torch.float64, device=dev) * eps
        tr = _trace_sqrt_product(sigma1 + offset, sigma2 + offset)
    st = torch.empty(2, 3, dtype=torch.float64, device=dev)
    K.dmat_stats(sigma1.contiguous(), st[0])
    K.dmat_stats(sigma2.contiguous(), st[1])
    d2 = diff.cpu().numpy()                       # the reference's diff.dot(diff), fp64 on the host
    return float(d2.dot(d2) + float(st[0, 0]) + float(st[1, 0]) - 2 * float(tr))


# ------------------------------------------------------------------ file-based variants (fid.py:197-318)
def load_image_batch(files):
    """fid.py:197-204: the images of `files` as one array [n,H,W,3].  uint8 (the forward's input stage reads bytes;
    the reference widens the same values to float32)."""
    from PIL import Image
    out = []
    for fn in files:
        with Image.open(str(fn)) as im:
            out.append(np.array(im.convert("RGB"), dtype=np.uint8))
    return np.stack(out)


def create_inception_graph(pth=None, dtype="bf16", allow_random=False):
    """fid.py:36-42 stand-in: the pool_3 feature extractor.  `pth`: the reference's own model file
    (`classify_image_graph_def.pb` or the .tgz around it - read without TensorFlow, `inception.load_graphdef`), or a
    converted `.npz` (`inception.load_npz`).  Without a model file this raises: an 'FID' from a randomly
    initialised network is a plausible-looking but meaningless number.  `allow_random=True` (throughput runs and
    tests only) builds the network on random weights."""
    from .inception import InceptionPool3, load_graphdef, load_npz
    if pth is None:
        if not allow_random:
            raise RuntimeError("no Inception model file given (the reference downloads inception-2015-12-05.tgz, "
                               "fid.py:273-288; there is no network here): pass the path of "
                               "classify_image_graph_def.pb / the .tgz / a converted .npz, or allow_random=True "
                               "for a throughput run on random weights")
        weights = None
    elif str(pth).endswith(".npz"):
        weights = load_npz(str(pth))
    else:
        weights = load_graphdef(str(pth))
    return InceptionPool3(weights=weights, dtype=dtype)


def check_or_download_inception(inception_path):
    """fid.py:273-288 without the download (no network here): the model file under `inception_path` - the
    reference's `classify_image_graph_def.pb`, the .tgz, or a converted `.npz` - or an error saying what to put there."""
    import pathlib
    if inception_path is None:
        return None
    p = pathlib.Path(inception_path)
    if p.is_file():
        return str(p)
    for name in ("classify_image_graph_def.pb", "inception-2015-12-05.tgz", "inception_2015_pool3.npz"):
        if (p / name).exists():
            return str(p / name)
    raise RuntimeError("no Inception model file under %s (no network: place classify_image_graph_def.pb from "
                       "inception-2015-12-05.tgz there)" % p)


def get_activations_from_files(files, sess, batch_size=50, verbose=False):
    """fid.py:207-240 as a generator of CUDA feature batches; the next batch is decoded on a host thread while the
    GPU runs the current one; the N % batch tail is dropped as in the reference."""
    from concurrent.futures import ThreadPoolExecutor
    d0 = len(files)
    if batch_size > d0:
        print("warning: batch size is bigger than the data size. setting batch size to data size")
        batch_size = d0
    n_batches = d0 // batch_size
    with ThreadPoolExecutor(max_workers=1) as pool:
        nxt = pool.submit(load_image_batch, files[:batch_size]) if n_batches else None
        for i in range(n_batches):
            if verbose:
                print("\rPropagating batch %d/%d" % (i + 1, n_batches), end="", flush=True)
            batch = nxt.result()
            if i + 1 < n_batches:
                nxt = pool.submit(load_image_batch, files[(i + 1) * batch_size:(i + 2) * batch_size])
            feats = sess(batch)
            if isinstance(feats, np.ndarray):
                feats = torch.from_numpy(feats)
            yield feats.reshape(batch_size, -1).to("cuda", torch.float32, non_blocking=True)
    if verbose:
        print(" done")


def calculate_activation_statistics_from_files(files, sess, batch_size=50, verbose=False, as_numpy=True):
    """fid.py:243-260 -> (mu, sigma) in fp64, streaming."""
    acc = None
    for feats in get_activations_from_files(files, sess, batch_size, verbose):
        if acc is None:
            acc = FeatureStatistics(feats.shape[1], feats.device)
        acc.update(feats)
    if acc is None:
        raise InvalidFIDException("no images")
    mu, sigma = acc.finalize()
    return (mu.cpu().numpy(), sigma.cpu().numpy()) if as_numpy else (mu, sigma)


def _handle_path(path, sess, low_profile=False):
    """fid.py:291-304: statistics of an .npz file (`mu`, `sigma`) or of the *.jpg / *.png files of a directory."""
    import pathlib
    path = str(path)
    if path.endswith(".npz"):
        with np.load(path) as f:
            return f["mu"][:], f["sigma"][:]
    p = pathlib.Path(path)
    files = sorted(p.glob("*.jpg")) + sorted(p.glob("*.png"))
    if low_profile:
        return calculate_activation_statistics_from_files(files, sess)
    return calculate_activation_statistics(load_image_batch(files), sess)


def calculate_fid_given_paths(paths, inception_path, low_profile=False, sess=None):
    """fid.py:307-318.  `sess`: a ready feature extractor (skips loading weights from `inception_path`)."""
    import os
    for p in paths:
        if not os.path.exists(p):
            raise RuntimeError("Invalid path: %s" % p)
    if sess is None:
        sess = create_inception_graph(check_or_download_inception(inception_path))
    m1, s1 = _handle_path(paths[0], sess, low_profile=low_profile)
    m2, s2 = _handle_path(paths[1], sess, low_profile=low_profile)
    return calculate_frechet_distance(m1, s1, m2, s2)
