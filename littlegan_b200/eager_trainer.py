"""EagerTrainer with the reference's surface (`eager_trainer.py:10-303`): losses, `_train_step`,
`train`, `predict`, checkpoint helpers - the arithmetic of one step is a fixed schedule of
hand-written CUDA kernels (engine.py), captured into CUDA graphs and replayed.

Data parallelism (new; the reference is single-device): when torch.distributed is initialised
each rank runs the step on its slice of the batch and the flat gradient arenas are all-reduced
(average) over NCCL before the value-clip and the Adam update, which is the gradient of the
global-batch mean - the reference's one-device semantics (SURVEY 8 e).
"""
import json
import os
import time

import numpy as np
import torch

from . import engine as E
from . import kernels as K
from .utils import data_rescale, save_image, soft, upload


class OutOfRangeError(Exception):
    """Raised by an iterator's get_next() at the end of an epoch (tf.errors.OutOfRangeError)."""


_ALIGN = 64  # elements; every parameter starts on a 256-byte boundary inside the arenas


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist
    return None


class LossValue:
    """A loss scalar of one train step (what the reference returns as a 0-d EagerTensor): the three losses of
    a step are copied to pinned host memory on a read-back stream right behind the step, and `float()` /
    `.item()` / `.numpy()` wait for THAT copy only - a loop that logs the losses of step i after submitting
    step i+1 never drains the GPU.  `.tensor` is the device value."""
    __slots__ = ("tensor", "_slot", "_gen", "_idx")

    def __init__(self, tensor, slot, gen, idx):
        self.tensor, self._slot, self._gen, self._idx = tensor, slot, gen, idx

    def __float__(self):
        host, event, gen = self._slot
        if gen[0] == self._gen:                 # the pinned slot still holds this step
            event.synchronize()
            if gen[0] == self._gen:
                return float(host[self._idx])
        return float(self.tensor)               # slot recycled: plain (stream-synchronising) device read

    def item(self):
        return float(self)

    def numpy(self):
        return np.float32(float(self))

    def cpu(self):
        return torch.tensor(float(self))

    def __array__(self, dtype=None, copy=None):
        return np.asarray(float(self), dtype=dtype or np.float32)

    def __format__(self, spec):
        return format(float(self), spec)

    def __repr__(self):
        return "LossValue(%r)" % float(self)


class EagerTrainer:
    def __init__(self, args, generator, discriminator, adjuster, dataset):
        self.args = args
        print(" - Initializing Trainer(Executor)...")
        self.dataset = dataset
        self.adjuster = adjuster
        self.discriminator = discriminator
        self.generator = generator
        self.models = [self.discriminator, self.generator, self.adjuster]
        self.global_epoch = 1
        self.rt = E.Runtime(args) if torch.cuda.is_available() else None
        for m in self.models:
            m._rt = self.rt

        # eager_trainer.py:48-63
        self.part_groups = {
            "Generator": [range(0, 4), range(4, 8), range(8, 22)],
            "Discriminator": [range(0, 12), range(12, 16), range(16, 20)],
            "Adjuster": [range(16, 20)],
        }
        self.part_weights = {}
        for model in self.models:
            name = model.__class__.__name__
            w = model.weights
            self.part_weights[name] = [[w[layer] for layer in group] for group in self.part_groups[name]]
        self.all_weights = {
            "Generator": self.generator.weights,
            "Discriminator": self.discriminator.weights,
            "Adjuster": [self.adjuster.weights[w] for w in range(16, 20)],
        }
        self._build_arenas()
        # tf.compat.v1.train.AdamOptimizer(lr, beta_1, beta_2) x2, AdamOptimizer(lr) (eager_trainer.py:28-30)
        self._hyper = {
            "Generator": (args.lr, args.beta_1, args.beta_2),
            "Discriminator": (args.lr, args.beta_1, args.beta_2),
            "Adjuster": (args.lr, 0.9, 0.999),
        }
        self._static = None
        self._graphs = {}
        self._seen = set()
        self._pool = None
        self._noise_state = None                 # {seed, step, ticket} of the generator-noise Philox stream (device)
        self._comm_stream = None
        self._reduced = []                       # this step's buckets in issue order: (optimiser, lo, hi, event)
        self._chain_streams = {}
        self._rb, self._rb_count = None, 0
        self._aug_state = None
        # config key `debug_taps` (tests): keep every per-layer activation of the last step in self.taps (the
        # references keep the buffers out of the allocator's reuse; the arithmetic and the launches are unchanged)
        self.taps = {} if getattr(args, "debug_taps", False) else None
        self.test_noise = self.test_cond = self.test_image = None
        if getattr(args, "result_dir", None) and getattr(args, "init_dirs", False):
            self._init_dir()
            self._restore_latest()

    # ------------------------------------------------------------------ parameter arenas
    def _build_arenas(self):
        """Re-home every trained tensor into one flat fp32 arena (D | G | A-own) with matching
        gradient / Adam-slot arenas, so clip + all-reduce + Adam are one launch per optimiser and
        a partition group is a contiguous range."""
        order = [("Discriminator", self.all_weights["Discriminator"]),
                 ("Generator", self.all_weights["Generator"]),
                 ("Adjuster", self.all_weights["Adjuster"])]
        off = 0
        self._offsets = {}
        for name, ws in order:
            offs = []
            for p in ws:
                offs.append(off)
                off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
            self._offsets[name] = offs + [off]
        dev = order[0][1][0].device
        self.P = torch.zeros(off, dtype=torch.float32, device=dev)
        self.Gd = torch.zeros(off, dtype=torch.float32, device=dev)
        self.M = torch.zeros(off, dtype=torch.float32, device=dev)
        self.V = torch.zeros(off, dtype=torch.float32, device=dev)
        self.adam_state = {name: torch.zeros(4, dtype=torch.float64, device=dev) for name, _ in order}
        for name, ws in order:
            for p, o in zip(ws, self._offsets[name]):
                n = p.numel()
                self.P[o:o + n].copy_(p.reshape(-1))
                p.set_(self.P.untyped_storage(), o, p.shape)
                p.lg_grad = self.Gd[o:o + n].view(p.shape)

    def _range(self, name, batch_no):
        """Flat [begin, end) of the tensors `_get_train_weight` selects for this step."""
        offs = self._offsets[name]
        a = self.args
        if a.use_partition and batch_no % (a.partition_interval + 1) == 0:
            groups = self.part_groups[name]
            grp = groups[(batch_no // (a.partition_interval + 1)) % len(groups)]
            base = 16 if name == "Adjuster" else 0
            return offs[grp[0] - base], offs[grp[-1] - base + 1]
        return offs[0], offs[-1]

    def _get_train_weight(self, model, batch_no):
        """eager_trainer.py:104-113."""
        name = model.__class__.__name__
        a = self.args
        if a.use_partition and batch_no % (a.partition_interval + 1) == 0:
            weights = self.part_weights[name]
            return weights[(batch_no // (a.partition_interval + 1)) % len(weights)]
        return self.all_weights[name]

    # ------------------------------------------------------------------ losses (public, forward only)
    @staticmethod
    def _bce_mean(target, p):
        p = p.float().contiguous()
        acc = torch.zeros(1, dtype=torch.float32, device=p.device)
        if torch.is_tensor(target):
            target = target.float().contiguous()
        K.bce_sigmoid(p, target, 1.0, acc, None)
        return acc[0]

    @staticmethod
    def discriminator_loss(real_true_c, real_predict_c, real_predict_pr, fake_predict_pr):
        """eager_trainer.py:85-91."""
        b = EagerTrainer._bce_mean
        return b(real_true_c, real_predict_c) * 2 + b(soft(1.0), real_predict_pr) + b(soft(0.0), fake_predict_pr)

    def _l1_mean(self, image_ori, image_gen):
        acc = torch.zeros(1, dtype=torch.float32, device=image_gen.device)
        y = image_gen.contiguous()
        K.l1_tanh_bwd(y, image_ori.to(y.dtype).contiguous(), None, None, 1.0, acc)
        return acc[0]

    def generator_loss(self, cond_ori, cond_disc, pr_disc, image_ori, image_gen):
        """eager_trainer.py:93-96."""
        b = EagerTrainer._bce_mean
        return b(soft(1.0), pr_disc) + b(cond_ori, cond_disc) + self.args.l1_lambda * self._l1_mean(image_ori, image_gen)

    def adjuster_loss(self, cond_ori, cond_disc, pr_disc, image_ori, image_adj):
        """eager_trainer.py:98-102."""
        return self.generator_loss(cond_ori, cond_disc, pr_disc, image_ori, image_adj)

    def _tap(self, **named):
        if self.taps is not None:
            self.taps.update(named)

    # ------------------------------------------------------------------ the step
    def _alloc_static(self):
        a, rt = self.args, self.rt
        B, H, C = a.batch_size, a.init_dim * 16, a.image_channel
        f32 = torch.float32
        S = {}
        S["in_img1"] = rt.empty(B, H, H, C, dtype=f32)      # H2D staging (fp32 as the reference feeds)
        S["in_img2"] = rt.empty(B, H, H, C, dtype=f32)
        S["in_new"] = rt.empty(B, H, H, C, dtype=f32)
        S["cond1"] = rt.empty(B, a.cond_dim, dtype=f32)
        S["cond2"] = rt.empty(B, a.cond_dim, dtype=f32)
        S["noise"] = rt.empty(B, a.noise_dim, dtype=f32)
        # One image buffer [real_image_1 | new_image | fake_image] and ONE set of encoder maps over it serve
        # D(new_image), D(fake) (the discriminator's batch is its last 2B images) AND the adjuster's
        # encoder([real_image_1 ; fake], eager_trainer.py:157): the adjuster shares D's encoder (model.py:119),
        # nothing has been updated yet and the norm is per sample, so encoder(fake) is the same tensor in both.
        # The two real thirds come first so that their encoder pass is one contiguous 2B batch that can run while
        # the generator is still producing `fake`.
        S["img3"] = rt.empty(3 * B, H, H, C)
        S["real1"] = S["img3"][:B]
        S["dimg"] = S["img3"][B:]                            # [new_image ; fake_image]
        S["fake"] = S["img3"][2 * B:]
        S["img2"] = rt.empty(B, H, H, C)
        S["aug_params"] = rt.empty(4 + 4 * B, dtype=f32)     # draws + per-image channel means of the augmentation
        S["aimg_t"] = rt.empty(2 * B, H, H, C)               # [real_image_2 ; real_image_1]
        S["acond_in"] = rt.empty(2 * B, a.cond_dim, dtype=f32)
        S["acond_t"] = rt.empty(2 * B, a.cond_dim, dtype=f32)
        S["loss"] = rt.zeros(4, dtype=f32)                   # gen, disc, adj
        S["adj"] = None
        return S

    def _conv_layers(self):
        return self.discriminator.encoder.convs + self.generator.decoder.convs + [self.generator.conv]

    def _prepare_inputs(self, S, aug):
        """Casts / concatenations of the step inputs (part of the captured step; this library's kernels only).
        aug: new_image is the reference's augmentation of real_image_1 (eager_trainer.py:127-131) instead of the
        staged `in_new`."""
        B = self.args.batch_size
        if aug:
            if self._aug_state is None:
                rank = _dist().get_rank() if _dist() is not None else 0
                self._aug_state = K.augment_state(int(getattr(self.args, "seed", 0)) * 7919 + 13 + rank, self.rt.device)
            K.augment(S["in_img1"], S["img3"][B:2 * B], S["aug_params"], state=self._aug_state)
        else:
            K.cast(S["in_new"], S["img3"][B:2 * B])
        K.cast(S["in_img2"], S["img2"])
        K.cast(S["in_img1"], S["real1"])
        K.cast(S["in_img2"], S["aimg_t"][:B])                # adj_target_image = [real_image_2 ; real_image_1]
        K.cast(S["in_img1"], S["aimg_t"][B:])
        # eager_trainer.py:155-156: adj_target_cond = [cond_2 ; cond_1], adj_input_cond = (adj_target_cond + 1) * 0.5
        K.scale_shift(S["cond2"], S["acond_t"][:B])
        K.scale_shift(S["cond1"], S["acond_t"][B:])
        K.scale_shift(S["acond_t"], S["acond_in"], 0.5, 0.5)

    def _step_body(self, S, adj_on, batch_no, aug=False, draw_noise=False):
        a, rt = self.args, self.rt
        B = a.batch_size
        G, D, A = self.generator, self.discriminator, self.adjuster
        f32 = torch.float32
        loss = S["loss"]
        l_gen, l_disc, l_adj = loss[0:1], loss[1:2], loss[2:3]

        # ---- prologue: the input staging / zeroing (stream P) runs beside the weight packing (this stream); the
        # generator forward needs only the latter
        main = torch.cuda.current_stream()
        sP = self._chain_stream("P", main)
        with torch.cuda.stream(sP):
            self.Gd.zero_()
            loss.zero_()
            self._prepare_inputs(S, aug)
        rt.begin_step()
        E.refresh_packs(rt, self._conv_layers())
        if draw_noise:
            # the generator noise ~ N(0,1) (eager_trainer.py:125) is drawn inside the step, on the device, from a
            # Philox stream whose step counter advances inside the launch: every replay of the graph draws anew
            if self._noise_state is None:
                rank = _dist().get_rank() if _dist() is not None else 0
                self._noise_state = K.normal_state(int(getattr(a, "seed", 0)) * 1000003 + rank, rt.device)
            K.normal_fill(S["noise"], self._noise_state)

        # ---- forward: G, and the encoder on [real_image_1 | new_image | fake] (eager_trainer.py:134-137, 157-160).
        # The real part of that batch does not depend on G: its encoder pass runs on a side stream while the
        # (small-batch, GPU-underfilling) generator forward produces `fake`; the fake third follows on this stream.
        off = B if adj_on else 0                                  # where D's batch [new_image ; fake] starts
        enc_in = S["img3"] if adj_on else S["dimg"]
        n_real = enc_in.shape[0] - B
        sliced = (bool(getattr(a, "overlap_chains", True)) and bool(getattr(a, "overlap_encoder", True))
                  and E.EncoderPass.sliceable(rt, D.encoder, enc_in))
        if sliced:
            ep = E.EncoderPass(rt, D.encoder, enc_in)
            sR = self._chain_stream("R", main)
            sR.wait_stream(sP)
            with torch.cuda.stream(sR):
                ep.run(0, n_real)
        fake, (g_hctx, g_dctx, g_x4) = G.forward_ctx(S["noise"], S["cond2"], out=S["fake"])
        main.wait_stream(sP)
        if sliced:
            ep.run(n_real, n_real + B)
            main.wait_stream(sR)
            outs3, ectx3 = ep.result()
        else:
            outs3, ectx3 = E.encoder_forward(rt, D.encoder, enc_in)
        outs = [o[off:] for o in outs3]
        ectx = [(x[off:], z[off:], st[off:], None if xp is None else xp[off:]) for (x, z, st, xp) in ectx3]
        self._tap(enc_off=off, enc=outs3, g_head=g_dctx[0][0], g_dec=[c[0] for c in g_dctx[1:]] + [g_x4], fake=fake)

        # The step is three chains that depend only on the forward pass and on the (read-only until Adam)
        # weights: the D-loss backward, the G-loss backward and the whole adjuster sub-step.  They run on
        # separate streams - parallel branches of the captured graph - so that the HBM-bound norm / loss / dense
        # kernels of one chain execute under the L2- and tensor-bound conv kernels of another.  Every chain
        # allocates its temporaries on its own stream (the caching allocator never hands a block to another
        # stream), and the tensors chains share stay referenced until the join below.
        # The adjuster sub-step (eager_trainer.py:152-164) is the longest chain: it forks first, on a
        # high-priority stream, and the other chains fill in around it.
        sA = self._chain_stream("A", main, high=True) if adj_on else None
        if adj_on:
            with torch.cuda.stream(sA):
                # the adjuster's batch is [real_image_1 ; fake]: rows [:B] and [2B:] of the encoder maps
                self._adjuster_chain(S, batch_no, [(o[:B], o[2 * B:]) for o in outs3])

        pr, c = E.disc_heads_forward(rt, D, outs[3])              # rows [:B] new_image, [B:] fake
        self._tap(pr=pr, c=c)

        # ---- losses + gradients w.r.t. the logits (eager_trainer.py:139-140)
        dl_pr_d = rt.empty(2 * B, 1, dtype=f32)
        dl_c_d = rt.zeros(2 * B, a.cond_dim, dtype=f32)          # fake half: no D-loss term on fake_c
        dl_pr_g = rt.empty(B, 1, dtype=f32)
        dl_c_g = rt.empty(B, a.cond_dim, dtype=f32)
        K.bce_sigmoid_multi([(c[:B], S["cond1"], 2.0, l_disc, dl_c_d[:B]),
                             (pr[:B], soft(1.0), 1.0, l_disc, dl_pr_d[:B]),
                             (pr[B:], soft(0.0), 1.0, l_disc, dl_pr_d[B:]),
                             (pr[B:], soft(1.0), 1.0, l_gen, dl_pr_g),
                             (c[B:], S["cond2"], 1.0, l_gen, dl_c_g)])

        # ---- disc_tape.gradient(disc_loss, D weights): both halves, no input gradient (:145)
        sD = self._chain_stream("D", main)
        with torch.cuda.stream(sD):
            g4d = E.disc_heads_backward(rt, D, outs[3], dl_pr_d, dl_c_d, wgrad=True)
            E.encoder_backward(rt, D.encoder, ectx, g4d, wgrad=True, input_grad=False)
            self._reduce_async("Discriminator", batch_no)    # overlaps with the G backward + adjuster step

        # ---- gen_tape.gradient(gen_loss, G weights): dgrad-only through D(fake), then G (:149)
        ectx_f = [(x[B:], z[B:], st[B:], None) for (x, z, st, _) in ectx]
        g4 = E.disc_heads_backward(rt, D, outs[3][B:], dl_pr_g, dl_c_g, wgrad=False)
        g_img = E.encoder_backward(rt, D.encoder, ectx_f, g4, wgrad=False, input_grad=True)
        self._tap(g_fake_via_D=g_img)
        dpre = torch.empty_like(fake)
        K.l1_tanh_bwd(fake, S["img2"], g_img, dpre, a.l1_lambda, l_gen)
        g = E.generator_tail_backward(rt, G.decoder, G.conv, g_dctx, g_x4, dpre, wgrad=True)
        self._reduce_async("Generator", batch_no, part=(4, 22))     # decoder + final conv: under the dense backward
        E.head_backward(rt, G.dense, G.norm, g_hctx, g)
        self._reduce_async("Generator", batch_no, part=(0, 4))
        if adj_on:
            # The collectives of one communicator execute in issue order, so the adjuster's bucket - produced by the
            # LONGEST chain - is issued last: issued from inside its chain (which is launched first) it made the
            # discriminator's and the generator's all-reduces wait for the whole adjuster sub-step and all four
            # ran, exposed, behind the last backward kernel (profiles/r2_dp2_timeline_before.txt).
            with torch.cuda.stream(sA):
                self._reduce_async("Adjuster", batch_no)

        main.wait_stream(sD)
        if adj_on:
            main.wait_stream(sA)

        # ---- apply (eager_trainer.py:164-168); D grads value-clipped (:146-148).  The reference applies A, D, G; the
        # three optimisers own disjoint tensors and every gradient was taken at the pre-update weights, so the order
        # is immaterial: the adjuster - whose all-reduce is the last collective - goes last, and each Adam waits only
        # for its own optimiser's all-reduce (the adjuster's runs under the Adams of D and G).
        names = ["Discriminator", "Generator"] + (["Adjuster"] if adj_on else [])
        for name in names:
            lr, b1, b2 = self._hyper[name]
            K.adam_advance(self.adam_state[name], lr, b1, b2)
        # data parallel: one Adam launch per reduced bucket, in the order the all-reduces were issued, each behind
        # its own all-reduce only (the generator's decoder range is updated while its dense-layer bucket is still
        # on the wire); single device: one launch per optimiser
        pieces, self._reduced = self._reduced, []
        if not pieces:
            pieces = [(name,) + self._range(name, batch_no) + (None,) for name in names]
        for name in names:                       # the buckets of an optimiser must tile its active range
            lo, hi = self._range(name, batch_no)
            mine = sorted((x[1], x[2]) for x in pieces if x[0] == name)
            assert mine and mine[0][0] == lo and mine[-1][1] == hi and all(
                x[1] == y[0] for x, y in zip(mine, mine[1:])), ("buckets do not tile the range", name, mine, lo, hi)
        for name, blo, bhi, ev in pieces:
            if ev is not None:
                torch.cuda.current_stream().wait_event(ev)
            _, b1, b2 = self._hyper[name]
            clip = a.clip_range if (name == "Discriminator" and a.use_clip) else 0.0
            K.adam_apply(self.P[blo:bhi], self.Gd[blo:bhi], self.M[blo:bhi], self.V[blo:bhi], self.adam_state[name],
                         b1, b2, 1e-8, clip)
        rt.end_step()

    def _chain_stream(self, name, main, high=False):
        """Stream of one independent chain of the step, ordered after everything issued so far on `main`."""
        if not bool(getattr(self.args, "overlap_chains", True)):
            return main
        s = self._chain_streams.get(name)
        if s is None:
            s = self._chain_streams[name] = torch.cuda.Stream(priority=-1 if high else 0)
        s.wait_stream(main)
        return s

    def _adjuster_chain(self, S, batch_no, enc):
        """eager_trainer.py:152-164 on the current stream: forward A on [image ; fake] (`enc`: the shared
        encoder's four maps of exactly that batch, already computed), D on the result, the gradient of adj_loss
        w.r.t. the adjuster's own dense + norm (dgrad-only through D and the decoder)."""
        a, rt = self.args, self.rt
        B = a.batch_size
        D, A = self.discriminator, self.adjuster
        f32 = torch.float32
        l_adj = S["loss"][2:3]
        adj, (a_hctx, a_dctx, a_x4) = A.forward_ctx(None, S["acond_in"], enc=enc)
        outs2, ectx2 = E.encoder_forward(rt, D.encoder, adj)
        apr, ac = E.disc_heads_forward(rt, D, outs2[3])
        dl_pr_a = rt.empty(2 * B, 1, dtype=f32)
        dl_c_a = rt.empty(2 * B, a.cond_dim, dtype=f32)
        K.bce_sigmoid_multi([(apr, soft(1.0), 1.0, l_adj, dl_pr_a), (ac, S["acond_t"], 1.0, l_adj, dl_c_a)])
        g4 = E.disc_heads_backward(rt, D, outs2[3], dl_pr_a, dl_c_a, wgrad=False)
        g_img = E.encoder_backward(rt, D.encoder, ectx2, g4, wgrad=False, input_grad=True)
        self._tap(a_head=a_dctx[0][0], a_dec=[c[0] for c in a_dctx[1:]] + [a_x4], adj=adj, da_enc=outs2, da_pr=apr,
                  da_c=ac, g_adj_via_D=g_img)
        dpre = torch.empty_like(adj)
        K.l1_tanh_bwd(adj, S["aimg_t"], g_img, dpre, a.l1_lambda, l_adj)
        g = E.generator_tail_backward(rt, A.decoder, A.conv, a_dctx, a_x4, dpre, wgrad=False)
        E.head_backward(rt, A.dense, A.norm, a_hctx, g)
        S["adj"] = adj

    def _bucket(self, name, batch_no, part=None):
        """Flat [lo, hi) of one gradient bucket: the optimiser's active range (`_range`), optionally cut down to
        the tensors [part[0], part[1]) of that optimiser; None when the cut is empty."""
        lo, hi = self._range(name, batch_no)
        if part is not None:
            offs = self._offsets[name]
            lo, hi = max(lo, offs[part[0]]), min(hi, offs[part[1]])
        return (lo, hi) if lo < hi else None

    @staticmethod
    def _all_reduce_mean(dist, buf):
        """Average `buf` over the ranks in place (NCCL: one AVG all-reduce; gloo has no AVG: SUM, then scale)."""
        if dist.get_backend() == "nccl":
            dist.all_reduce(buf, op=dist.ReduceOp.AVG)
        else:
            dist.all_reduce(buf, op=dist.ReduceOp.SUM)
            buf.div_(dist.get_world_size())

    def _reduce_async(self, name, batch_no, part=None):
        """Data parallel: average this optimiser's (active range of the) flat gradient arena over the
        ranks on a side stream, so that NCCL runs under the remaining backward work; the ranges of
        the three optimisers are disjoint and the main stream joins before the first Adam.
        part = (first, last+1) tensor indices of the optimiser: reduce only that bucket of the range (the
        generator's decoder gradients are complete while its dense-layer gradients are still being computed).
        Host arenas (the gloo protocol tests): the same bucket arithmetic, reduced synchronously."""
        dist = _dist()
        if dist is None:
            return
        rng = self._bucket(name, batch_no, part)
        if rng is None:
            return
        buf = self.Gd[rng[0]:rng[1]]
        if not buf.is_cuda:
            self._all_reduce_mean(dist, buf)
            self._reduced.append((name, rng[0], rng[1], None))
            return
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream()
        # (fp32 on the wire: sending the buckets as bf16 was measured at 8 GPUs - 3.44 ms either way; the ring
        # all-reduces of these 4-18 MB buckets are latency / rank-skew bound, not bandwidth bound - and dropped.)
        self._comm_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self._comm_stream):
            self._all_reduce_mean(dist, buf)
            ev = torch.cuda.Event()
            ev.record(self._comm_stream)
        self._reduced.append((name, rng[0], rng[1], ev))      # issue order: each Adam piece waits for its own bucket

    def _variant(self, batch_no):
        a = self.args
        adj_on = bool(a.train_adj and batch_no > 10)
        grp = None
        if a.use_partition and batch_no % (a.partition_interval + 1) == 0:
            grp = (batch_no // (a.partition_interval + 1)) % 3
        return adj_on, grp

    def _run_step(self, S, batch_no, aug=False, draw_noise=False):
        adj_on, grp = self._variant(batch_no)
        key = (adj_on, grp, aug) if not draw_noise else (adj_on, grp, aug, True)
        use_graph = bool(getattr(self.args, "cuda_graph", True))
        if not use_graph or key not in self._seen:
            # first occurrence of a variant runs eagerly (it also warms up lazy state before capture)
            self._seen.add(key)
            self._step_body(S, adj_on, batch_no, aug, draw_noise)
            return
        g = self._graphs.get(key)
        if g is None:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            if self._pool is None:
                self._pool = torch.cuda.graph_pool_handle()
            n0 = K.launch_count()
            with torch.cuda.graph(g, pool=self._pool):
                self._step_body(S, adj_on, batch_no, aug, draw_noise)
            K.note_capture(K.launch_count() - n0)
            self._graphs[key] = (g, S["adj"], None if self.taps is None else dict(self.taps))
        g, adj, taps = self._graphs[key]
        S["adj"] = adj
        if taps is not None:
            self.taps = dict(taps)              # the buffers THIS variant's graph writes
        g.replay()

    def _to_static(self, dst, src):
        if isinstance(src, np.ndarray):
            src = torch.from_numpy(src)
        if src.dtype == torch.uint8:
            # decoded image bytes (dataset.CelebA): data_rescale (utils.py:51-52) runs on the device
            K.u8_rescale(src.to(dst.device, non_blocking=True).contiguous(), dst)
            return
        dst.copy_(src.reshape(dst.shape), non_blocking=True)

    def _train_step(self, batch_no, iterator, noise=None, new_image=None):
        """eager_trainer.py:115-169.  Returns (True, fake_image, adj_image | None, gen_loss, disc_loss, adj_loss | None).
        `fake_image` / `adj_image` are VIEWS of buffers the next step overwrites (the reference returns fresh tensors):
        `.clone()` them to keep them across steps (`train()` writes its image dumps before the next step); the losses
        are `LossValue` objects that stay valid.
        `noise` / `new_image` may be injected for reproducible runs; by default
        noise ~ N(0,1) is drawn on the device and new_image is the reference's augmentation of real_image_1
        (eager_trainer.py:127-131: flip / brightness / contrast / hue / Gaussian noise, csrc/augment.cu) with
        device-side Philox draws; `augment: false` in the config makes new_image = real_image_1."""
        a = self.args
        try:
            real_image_1, real_cond_1 = iterator.get_next()
            real_image_2, real_cond_2 = iterator.get_next()
        except OutOfRangeError:
            return None,
        if not real_cond_1.shape[0] == real_cond_2.shape[0] == a.batch_size:
            return False,
        if a.use_gp:
            raise NotImplementedError("GP didn't implemented on eager mode")
        if self.rt is None:
            raise K._lib.LittleGANError("littlegan_b200 has no CPU path: a CUDA device is required")
        if self._static is None:
            self._static = self._alloc_static()
        S = self._static
        self._to_static(S["in_img1"], real_image_1)
        self._to_static(S["in_img2"], real_image_2)
        self._to_static(S["cond1"], real_cond_1)
        self._to_static(S["cond2"], real_cond_2)
        # new_image: injected, else the reference's augmentation of real_image_1 (config key `augment`, default
        # on as in the reference), else real_image_1 itself
        aug = new_image is None and bool(getattr(a, "augment", True))
        if not aug:
            self._to_static(S["in_new"], real_image_1 if new_image is None else new_image)
        if noise is not None:
            self._to_static(S["noise"], noise)
        adj_on, _ = self._variant(batch_no)
        self._run_step(S, batch_no, aug, draw_noise=noise is None)
        B = a.batch_size
        losses = self._read_back(S["loss"].clone())
        fake_image = S["fake"]
        adj_image = S["adj"] if adj_on else None
        return True, fake_image, adj_image, losses[0], losses[1], (losses[2] if adj_on else None)

    _RB_SLOTS = 16

    def _read_back(self, dev):
        """Start the device->host copy of this step's three loss scalars on the read-back stream."""
        if self._rb is None:
            self._rb = (torch.cuda.Stream(),
                        [(torch.empty(dev.numel(), dtype=torch.float32).pin_memory(), torch.cuda.Event(), [0])
                         for _ in range(self._RB_SLOTS)])
        stream, slots = self._rb
        self._rb_count += 1
        slot = slots[self._rb_count % self._RB_SLOTS]
        host, event, gen = slot
        stream.wait_stream(torch.cuda.current_stream())
        if gen[0]:
            event.synchronize()                 # a reader of the step that owned this slot may be mid-copy
        gen[0] = self._rb_count
        with torch.cuda.stream(stream):
            host.copy_(dev, non_blocking=True)
            event.record(stream)
        return [LossValue(dev[i], slot, self._rb_count, i) for i in range(3)]

    # ------------------------------------------------------------------ epoch loop (host glue)
    def _ckpt_dir(self):
        a = self.args
        d = os.path.join(a.result_dir, "checkpoint") if getattr(a, "result_dir", None) else None
        return d if d and os.path.isdir(d) else None

    def _restore_latest(self):
        """eager_trainer.py:37-43: with `restore`, load the checkpoint status.json names and resume at its epoch."""
        d = self._ckpt_dir()
        if not d or not getattr(self.args, "restore", False):
            return False
        status = os.path.join(d, "status.json")
        if not os.path.isfile(status):
            return False
        with open(status) as f:
            st = json.load(f)
        ck = os.path.join(d, st.get("latest", ""))
        if not os.path.isfile(ck):
            return False
        print("Loading Checkpoint...")
        self.load_checkpoint(ck)
        self.global_epoch = int(st["epoch"])
        return True

    def _write_status(self, name):
        with open(os.path.join(self._ckpt_dir(), "status.json"), "w") as f:
            json.dump({"epoch": self.global_epoch, "latest": name}, f)

    def _interrupted(self, signum, f_name):
        """eager_trainer.py:171-178: SIGINT saves a checkpoint + status.json and exits."""
        if self._ckpt_dir():
            torch.cuda.synchronize()
            self.save_checkpoint(os.path.join(self._ckpt_dir(), "interrupt.pt"))
            self._write_status("interrupt.pt")
            print("\n Checkpoint has been saved")
        print(signum, f_name)
        import sys
        sys.exit(1)

    def _init_test_data(self):
        """eager_trainer.py:65-83: the fixed (noise, cond, image) triple of the periodic `predict` dumps - re-used
        from test_data_<env>.npz with `reuse`, else the first batch of the dataset (saved there when possible)."""
        a = self.args
        npz = os.path.join(getattr(a, "test_data_dir", "") or "", "test_data_%s.npz" % getattr(a, "env", "sample"))
        if os.path.isfile(npz) and getattr(a, "reuse", False):
            data = np.load(npz)
            self.test_noise, self.test_cond, self.test_image = data["n"], data["c"], data["i"]
            return
        print("No reuse test data, generating...")
        image, cond = self.dataset.get_new_iterator().get_next()
        if torch.is_tensor(image) and image.dtype == torch.uint8:      # decoded bytes from dataset.CelebA
            image = data_rescale(image.float())
        to_np = lambda t: t.detach().float().cpu().numpy() if torch.is_tensor(t) else np.asarray(t, np.float32)
        self.test_image, self.test_cond = to_np(image), to_np(cond)
        self.test_noise = np.random.default_rng(int(getattr(a, "seed", 0))).standard_normal(
            (self.test_cond.shape[0], a.noise_dim)).astype(np.float32)
        if os.path.isdir(os.path.dirname(npz)):
            np.savez_compressed(npz, n=self.test_noise, c=self.test_cond, i=self.test_image)

    def train(self):
        """eager_trainer.py:180-229 (loss scalars are read back every `log_every` steps only)."""
        a = self.args
        log_every = int(getattr(a, "log_every", 50))
        import signal
        import threading
        if threading.current_thread() is threading.main_thread():
            signal.signal(signal.SIGINT, self._interrupted)
        dump = bool(getattr(a, "result_dir", None)) and os.path.isdir(os.path.join(a.result_dir, "test"))
        for e in range(self.global_epoch, a.epoch + 1):
            print("Experiment:", a.exp_name, "Epoch:", e, "Starting...")
            self.global_epoch = e
            iterator = self.dataset.get_new_iterator()
            start_time = time.time()
            for b in range(1, self.dataset.batches + 1):
                result = self._train_step(b, iterator)
                if result[0] is None:
                    break
                elif not result[0]:
                    continue
                if b % log_every == 0:
                    vals = [("LossG", result[3]), ("LossD", result[4]), ("LossA", result[5])]
                    print("  batch %d: " % b + " ".join("%s %.4f" % (k, float(v)) for k, v in vals if v is not None))
                if getattr(a, "result_dir", None) and os.path.isdir(os.path.join(a.result_dir, "train")):
                    if b % a.freq_gen == 0:
                        save_image(result[1], os.path.join(a.result_dir, "train", "gen", "%d-%d.jpg" % (e, b)))
                        if result[2] is not None:
                            save_image(result[2], os.path.join(a.result_dir, "train", "adj", "%d-%d.jpg" % (e, b)))
                if dump and b % a.freq_test == 0:
                    if self.test_image is None:
                        self._init_test_data()
                    r = a.result_dir
                    self.predict(self.test_noise, self.test_cond, self.test_image,
                                 os.path.join(r, "test", "gen", "%d-%d.jpg" % (e, b)),
                                 os.path.join(r, "test", "disc", "%d-%d.json" % (e, b)),
                                 os.path.join(r, "test", "adj", "%d-%d.jpg" % (e, b)))
            torch.cuda.synchronize()
            print("Time usage:", time.time() - start_time, "s")
            if self._ckpt_dir():
                self.global_epoch = e + 1                      # a restored run continues with the next epoch
                self.save_checkpoint(os.path.join(self._ckpt_dir(), "%d.pt" % e))
                self._write_status("%d.pt" % e)

    def _init_dir(self):
        a = self.args
        for item in [".", "train/gen", "train/adj", "test/adj", "test/gen", "test/disc", "checkpoint", "log",
                     "sample", "evaluate/gen", "evaluate/adj", "evaluate/disc", "model"]:
            os.makedirs(os.path.join(a.result_dir, item), exist_ok=True)
        with open(os.path.join(a.result_dir, "config.json"), "w") as f:
            json.dump({k: v for k, v in a.__dict__.items() if isinstance(v, (int, float, str, bool, list, type(None)))}, f)

    def save_checkpoint(self, path):
        torch.save({"P": self.P, "M": self.M, "V": self.V, "epoch": self.global_epoch,
                    "adam": {k: v for k, v in self.adam_state.items()}}, path)

    def load_checkpoint(self, path):
        ck = torch.load(path, map_location=self.P.device)
        self.P.copy_(ck["P"]); self.M.copy_(ck["M"]); self.V.copy_(ck["V"])
        for k, v in ck["adam"].items():
            self.adam_state[k].copy_(v)
        self.global_epoch = ck["epoch"]

    def export_model_checkpoint(self):
        """eager_trainer.py:300-303: weights only."""
        path = os.path.join(self.args.result_dir, "model", "model.pt")
        torch.save({"P": self.P, "offsets": self._offsets}, path)

    def plot(self):
        """eager_trainer.py:247-263: parameter summary (Keras plot_model has no equivalent here)."""
        lines = []
        for m in [self.discriminator.encoder, self.generator.decoder, self.discriminator, self.generator,
                  self.adjuster]:
            lines.append("%s: %d tensors, %d parameters" % (m.__class__.__name__, len(m.weights),
                                                            sum(w.numel() for w in m.weights)))
        return "\n".join(lines)

    # ------------------------------------------------------------------ predict
    @torch.no_grad()
    def predict(self, noise, cond, image, gen_image_save_path=None, json_save_path=None, adj_image_save_path=None):
        """eager_trainer.py:265-298: G(noise, cond); D(image), D(gen_image); the four MSE scalars; A(image, cond),
        A(gen_image, cond).  The reference runs 1 G + 2 D + 2 A forward passes; D and A share one encoder
        (model.py:119) and every norm is per sample, so ONE encoder pass over [image ; gen_image] serves both
        discriminator calls and both adjuster calls, and one decoder pass serves both adjuster calls."""
        G, D, A = self.generator, self.discriminator, self.adjuster
        rt, a = self.rt, self.args
        if rt is None:
            raise K._lib.LittleGANError("littlegan_b200 has no CPU path: a CUDA device is required")

        def dev(x, dtype):                     # host -> device in the source dtype, cast on the device
            return upload(x, rt.device).to(dtype).contiguous()

        noise_d, cond_d = dev(noise, torch.float32), dev(cond, torch.float32)
        image_d = dev(image, rt.act_dtype)
        B = image_d.shape[0]
        E.refresh_packs(rt, self._conv_layers())
        both = rt.empty(2 * B, *image_d.shape[1:])
        both[:B].copy_(image_d)
        gen_image = G.forward_ctx(noise_d, cond_d, out=both[B:])[0]
        if gen_image_save_path is not None:
            save_image(gen_image, gen_image_save_path)
        enc, _ = E.encoder_forward(rt, D.encoder, both)
        pr, c = E.disc_heads_forward(rt, D, enc[3])
        save = dict()
        save["real_cond"] = cond_d
        save["real_pr"], save["real_c"] = pr[:B], c[:B]
        save["fake_pr"], save["fake_c"] = pr[B:], c[B:]
        mse = lambda t, p: ((t - p) ** 2).mean(dim=-1).mean(dim=0)
        mses = torch.stack([mse(soft(1.0), save["real_pr"]), mse(cond_d, save["real_c"]),
                            mse(soft(0.0), save["fake_pr"]), mse(cond_d, save["fake_c"])]).tolist()
        save["real_pr_mse"], save["real_c_mse"], save["fake_pr_mse"], save["fake_c_mse"] = mses
        for x in ["real_cond", "real_pr", "real_c", "fake_c", "fake_pr"]:
            save[x] = torch.round(save[x] * 100).to(torch.int64).cpu().tolist()
        if json_save_path is not None:
            with open(json_save_path, "w") as f:
                json.dump(save, f)
        adj_fake_image, adj_real_image = None, None
        if a.train_adj:
            # note: raw cond here, as the reference (:292)
            adj = A.forward_ctx(both, torch.cat([cond_d, cond_d], 0), enc=enc)[0]
            adj_real_image, adj_fake_image = adj[:B], adj[B:]
            if adj_image_save_path is not None:
                save_image(adj, adj_image_save_path)
        return gen_image, save, adj_real_image, adj_fake_image
