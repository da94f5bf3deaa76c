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


class DevicePrefetcher:
    """Wraps an iterator of pinned host batches: the host->device copy of the next batches is
    issued on a side stream while the current step computes (the role of tf.data's prefetch,
    dataset.py:23).  get_next() returns CUDA tensors and makes the consumer stream wait for their
    copy; a buffer is recycled only after the consumer has read it."""

    def __init__(self, iterator, depth=4):
        self.it = iterator
        self.stream = torch.cuda.Stream()
        self.depth = depth
        self.ring = []            # (copy-done event, image_dev, cond_dev, consumed event)
        self.bufs = None
        self.slot = 0
        self.done = False
        self._prev = None         # consumed-event of the buffer handed out by the previous get_next
        for _ in range(depth):
            self._issue()

    def _issue(self):
        if self.done:
            return
        try:
            img, cond = self.it.get_next()
        except OutOfRangeError:
            self.done = True
            return
        if self.bufs is None:
            self.bufs = [(torch.empty(img.shape, dtype=img.dtype, device="cuda"),
                          torch.empty(cond.shape, dtype=cond.dtype, device="cuda"),
                          torch.cuda.Event()) for _ in range(self.depth + 2)]
        dimg, dcond, consumed = self.bufs[self.slot % len(self.bufs)]
        self.slot += 1
        self.stream.wait_event(consumed)        # no-op until the event has been recorded once
        with torch.cuda.stream(self.stream):
            dimg.copy_(img, non_blocking=True)
            dcond.copy_(cond, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.ring.append((ev, dimg, dcond, consumed))

    def get_next(self):
        cur = torch.cuda.current_stream()
        if self._prev is not None:
            self._prev.record(cur)              # the consumer's reads of the previous buffer are queued by now
            self._prev = None
        if not self.ring:
            raise OutOfRangeError()
        ev, dimg, dcond, consumed = self.ring.pop(0)
        cur.wait_event(ev)
        self._prev = consumed
        self._issue()
        return dimg, dcond
