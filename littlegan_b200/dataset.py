"""Synthetic CelebA-shaped data source with the reference dataset's surface (`dataset.py:8-48`):
`.batches`, `.get_new_iterator()` -> object with `.get_next() -> (image, cond)`.

The reference's JPEG pipeline (tf.data glob/decode/shuffle, `dataset.py:15-31`) is host I/O and out
of scope; benchmarks and tests use this generator: images U(-1,1) [B,H,W,C] fp32 (what
`data_rescale` yields), labels soft(+-1) in {-0.94, 0.98} (`dataset.py:33`, `utils.py:47-48`).
Batches are produced in pinned host memory so the train step's host->device copy is asynchronous.
"""
import torch

from .eager_trainer import OutOfRangeError
from .utils import soft


class _Iterator:
    def __init__(self, owner):
        self.owner, self.i = owner, 0

    def get_next(self):
        if self.i >= self.owner.batches:
            raise OutOfRangeError()
        out = self.owner._batch(self.i)
        self.i += 1
        return out


class SyntheticCelebA:
    def __init__(self, args, batches=64, seed=0, pool=8, pin=None):
        self.args = args
        self.batches = batches
        g = torch.Generator().manual_seed(seed)
        H, C, B = args.init_dim * 16, args.image_channel, args.batch_size
        pin = torch.cuda.is_available() if pin is None else pin
        self._pool = []
        for _ in range(max(2, pool)):
            img = torch.rand(B, H, H, C, generator=g) * 2 - 1
            lab = soft((torch.rand(B, args.cond_dim, generator=g) < 0.5).float() * 2 - 1)
            if pin:
                img, lab = img.pin_memory(), lab.pin_memory()
            self._pool.append((img, lab))

    def _batch(self, i):
        return self._pool[i % len(self._pool)]

    def get_new_iterator(self):
        return _Iterator(self)
