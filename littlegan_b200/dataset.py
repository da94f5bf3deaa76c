"""Data sources with the reference dataset's surface (`dataset.py:8-48`): `.batches`, `.label`,
`.get_new_iterator()` -> object with `.get_next() -> (image, cond)`.

* `CelebA(args)` - the reference's file pipeline (`dataset.py:8-48`: glob + attribute list -> decode -> batch ->
  shuffle(prefetch) -> prefetch) re-planned for a B200 box: worker threads decode JPEGs straight into PINNED
  uint8 batch buffers, the batch crosses PCIe as bytes (4x less than the reference's fp32 tensors) and
  `data_rescale` (`utils.py:51-52`) runs on the device (`lg_u8_rescale`), either in `DevicePrefetcher` or in the
  train step's input staging.  Labels are `soft(float(attr))` fp32 (`dataset.py:33`).
* `SyntheticCelebA` - what benchmarks and tests use: images U(-1,1) [B,H,W,C] fp32 (what `data_rescale` yields),
  labels soft(+-1) in {-0.94, 0.98}, in pinned host memory.
* `DevicePrefetcher` - the role of tf.data's `prefetch` (`dataset.py:23`) for either source.
"""
import os
import random
from concurrent.futures import ThreadPoolExecutor
from glob import glob

import numpy as np
import torch

from .eager_trainer import OutOfRangeError
from .utils import data_rescale, soft


class _Iterator:
    def __init__(self, owner):
        self.owner, self.i = owner, 0

    def get_next(self):
        if self.i >= self.owner.batches:
            raise OutOfRangeError()
        out = self.owner._batch(self.i)
        self.i += 1
        return out


ALL_LABEL = ["有短髭", "柳叶眉", "有魅力", "有眼袋", "秃头", "有刘海", "大嘴唇", "大鼻子", "黑发", "金发", "睡眼惺松", "棕发", "浓眉",
             "丰满", "双下巴", "眼镜", "山羊胡", "白发", "浓妆", "高颧骨", "男性", "嘴轻微张开", "八字胡", "眯缝眼", "完全没有胡子",
             "鹅蛋脸", "白皮肤", "尖鼻子", "发际线高的", "脸红的", "有鬓脚", "微笑", "直发", "卷发", "戴耳环", "戴帽子", "涂口红",
             "戴项链", "戴领带", "年轻人"]


class _ShuffledBatches:
    """One epoch of the reference's `batch(B).shuffle(prefetch).prefetch(prefetch)` (`dataset.py:21-23`): batches
    are formed from consecutive files FIRST and then shuffled through a buffer of `prefetch` batches (tf.data's
    algorithm: fill the buffer, emit a random slot, refill it with the next batch).  Decoding runs `ahead`
    batches in front of the consumer on the dataset's thread pool."""

    def __init__(self, owner, seed):
        self.o = owner
        self.rng = random.Random(seed)
        n, B = len(owner._image_list), owner.args.batch_size
        self.spans = [(i, min(i + B, n)) for i in range(0, n, B)]      # the last batch may be short, as in tf.data
        self.next_span = 0
        self.buffer = []                                               # futures of decoded batches
        self.cap = max(1, int(owner.args.prefetch))
        self._fill()

    def _fill(self):
        while len(self.buffer) < self.cap and self.next_span < len(self.spans):
            self.buffer.append(self.o._submit(*self.spans[self.next_span]))
            self.next_span += 1

    def get_next(self):
        if not self.buffer:
            raise OutOfRangeError()
        k = self.rng.randrange(len(self.buffer))
        fut = self.buffer[k]
        if self.next_span < len(self.spans):
            self.buffer[k] = self.o._submit(*self.spans[self.next_span])
            self.next_span += 1
        else:
            self.buffer[k] = self.buffer[-1]
            self.buffer.pop()
        return fut.result()


class CelebA:
    """`dataset.py:8-48`.  `get_next()` yields (image uint8 [b,H,W,C] in pinned memory, cond fp32 [b,cond_dim]);
    the train step / `DevicePrefetcher` turn the bytes into `data_rescale`d activations on the device.
    `as_float=True` restores the reference's host-side fp32 tensors (x / 127.5 - 1)."""

    def __init__(self, args, seed=0, pin=None, as_float=False):
        print(" - Initializing Dataset...")
        self.args = args
        # the reference pairs glob order with attribute-file order by index (`dataset.py:19`); sorted so that the
        # pairing is the same on every filesystem
        self._image_list = sorted(glob(os.path.join(args.image_path, "*." + args.image_ext)))
        self._attributes_list = self._get_attr_list(args.attr_path, args.attr)
        if len(self._attributes_list) < len(self._image_list):
            raise ValueError("attribute list has %d rows for %d images" % (len(self._attributes_list),
                                                                             len(self._image_list)))
        self.batches = len(self._image_list) // args.batch_size
        self.all_label = ALL_LABEL
        self.label = [self.all_label[x] for x in args.attr] if args.attr is not None else list(ALL_LABEL)
        self._dim = getattr(args, "image_dim", args.init_dim * 16)
        self._pin = torch.cuda.is_available() if pin is None else pin
        self._as_float = as_float
        self._seed, self._epoch = seed, 0
        self._pool = ThreadPoolExecutor(max_workers=max(1, int(getattr(args, "threads", 8))))
        self._labels = soft(torch.tensor(np.asarray(self._attributes_list[:len(self._image_list)], dtype=np.float32)
                                         .reshape(len(self._image_list), -1)))
        if not hasattr(args, "prefetch"):
            args.prefetch = getattr(args, "prefetch_batch", 3) * args.batch_size      # config.py:40

    @staticmethod
    def _get_attr_list(attr_file, attr_filter):
        """`dataset.py:35-45`: every line is `<name> <attr> <attr> ...`; keeps the columns in `attr_filter`."""
        with open(attr_file) as f:
            rows = f.read().splitlines()
        out = []
        for item in rows:
            raw = item.split()[1:]
            out.append(raw if attr_filter is None else [raw[x] for x in attr_filter])
        return out

    def _decode_into(self, filename, dst):
        from PIL import Image
        with Image.open(filename) as im:
            im = im.convert("L" if self.args.image_channel == 1 else "RGB")
            a = np.array(im, dtype=np.uint8)
        if a.ndim == 2:
            a = a[:, :, None]
        if a.shape != tuple(dst.shape):        # tf's set_shape (`dataset.py:28`) rejects any other size
            raise ValueError("%s: decoded shape %s, expected %s" % (filename, a.shape, tuple(dst.shape)))
        dst.copy_(torch.from_numpy(a))

    def _load_batch(self, lo, hi):
        C = self.args.image_channel
        img = torch.empty(hi - lo, self._dim, self._dim, C, dtype=torch.uint8)
        if self._pin:
            img = img.pin_memory()
        for i in range(lo, hi):
            self._decode_into(self._image_list[i], img[i - lo])
        cond = self._labels[lo:hi].clone()
        if self._pin:
            cond = cond.pin_memory()
        if self._as_float:
            img = data_rescale(img.float())
            if self._pin:
                img = img.pin_memory()
        return img, cond

    def _submit(self, lo, hi):
        return self._pool.submit(self._load_batch, lo, hi)

    def get_new_iterator(self):
        self._epoch += 1
        return _ShuffledBatches(self, self._seed * 7919 + self._epoch)


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
    copy.

    Buffer recycling contract: the consumer must have QUEUED its reads of a batch on the stream that was
    current at get_next() before the second-next get_next() call (`_train_step` draws two batches and then
    reads both, eager_trainer.py:120-121).  The `consumed` event of a batch is therefore recorded one call
    late - at the start of the second get_next() after the one that handed it out - and the copy that reuses
    its slot (depth + 2 slots: issued at the earliest in that same call, after the record) waits for it."""

    def __init__(self, iterator, depth=4):
        self.it = iterator
        self.stream = torch.cuda.Stream()
        self.depth = depth
        self.ring = []            # (copy-done event, image_dev, cond_dev, consumed event)
        self.bufs = None
        self.slot = 0
        self.done = False
        self._handed = []         # consumed-events of the batches handed out by the last two get_next calls
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
        if img.shape[0] < dimg.shape[0]:        # the short last batch of an epoch (tf.data keeps it)
            dimg, dcond = dimg[:img.shape[0]], dcond[:img.shape[0]]
        self.stream.wait_event(consumed)        # no-op until the event has been recorded once
        with torch.cuda.stream(self.stream):
            dimg.copy_(img, non_blocking=True)
            dcond.copy_(cond, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.ring.append((ev, dimg, dcond, consumed))

    def get_next(self):
        cur = torch.cuda.current_stream()
        # the batch handed out two calls ago: its reads are queued by now (see the class docstring)
        while len(self._handed) >= 2:
            self._handed.pop(0).record(cur)
        if not self.ring:
            raise OutOfRangeError()
        ev, dimg, dcond, consumed = self.ring.pop(0)
        cur.wait_event(ev)
        self._handed.append(consumed)
        self._issue()
        return dimg, dcond
