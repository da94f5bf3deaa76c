#!/usr/bin/env python
"""Reference-held golden vectors: run the UNMODIFIED reference (model.py, instance.py, eager_trainer.py under
--reference, default /root/reference) on TensorFlow 1.15 and dump what the oracle must reproduce.

    python scripts/make_tf_goldens.py [--reference /root/reference] [--out tests/golden]
                                      [--inception /path/to/classify_image_graph_def.pb]

Runnable only where `tensorflow==1.15.x` (requirements.txt:2) is installed - NOT in this repo's build container
(no cp312 wheel, no network); a maintainer with a TF-1.15 environment runs it once and commits the two small
files it writes.  `tests/test_oracle.py::test_oracle_matches_tf_goldens` consumes them when present (skipped
otherwise); until then the train-step oracle stays "parity unpinned" (oracle/__init__.py).

What it does (nothing of the reference is edited or re-implemented):
  * builds the reference's Encoder / Decoder / Generator / Discriminator / Adjuster from `model.py` on the
    reduced architecture of tests/util.small_args (32x32 images: the files stay < 3 MB) and assigns them the
    oracle's seeded initial weights (`oracle.littlegan_oracle.init_weights`, stored in the file too);
  * creates an `EagerTrainer` WITHOUT running its `__init__` (which needs a dataset, a result directory and a git
    checkout) and gives it exactly the attributes `_train_step` reads (`eager_trainer.py:28-30,48-63`);
  * makes the step reproducible from OUTSIDE: `tf.random.normal` and the four `tf.image.random_*` calls of
    `_train_step` (`eager_trainer.py:125-131`) are replaced for the duration of the call by functions that
    return the injected noise / the un-augmented image, and the three optimisers' `apply_gradients` are wrapped
    to record the gradients they receive;
  * tf_step.npz: one step at batch_no 11 (adjuster on) and one at batch_no 15 (a partition step) from the initial
    weights: inputs, per-layer encoder / generator / adjuster activations, the three losses, every gradient, the
    updated weights;
  * tf_trajectory.json: the G / D / A losses of 100 consecutive steps (batches = synthetic_batch(seed=1000+b));
  * with --inception: tf_pool3.npz = pool_3:0 of the real 2015 graph on a seeded 4-image batch (fid.py:36-106).
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _small_args(**over):
    from tests.util import small_args
    a = small_args(**over)
    a.image_dim = a.init_dim * 16
    return a


def _build_reference(ref_dir, args, tf):
    sys.path.insert(0, ref_dir)
    import model as RM                                     # the reference's model.py, unmodified
    decoder, encoder = RM.Decoder(args), RM.Encoder(args)
    generator = RM.Generator(args, decoder)
    discriminator = RM.Discriminator(args, encoder)
    adjuster = RM.Adjuster(args, discriminator, generator)
    H = args.init_dim * 16
    B = 2
    # variables are created at first call (eager_trainer.py:70-75 does the same)
    discriminator(tf.zeros([B, H, H, args.image_channel]))
    generator([tf.zeros([B, args.noise_dim]), tf.zeros([B, args.cond_dim])])
    adjuster([tf.zeros([B, H, H, args.image_channel]), tf.zeros([B, args.cond_dim])])
    return generator, discriminator, adjuster


def _assign(generator, discriminator, adjuster, W):
    """Oracle weight lists (model.weights order) -> the reference's variables."""
    for var, w in zip(discriminator.weights, W["D"]):
        assert tuple(var.shape) == tuple(w.shape), (var.name, var.shape, w.shape)
        var.assign(w.numpy())
    for var, w in zip(generator.weights, W["G"]):
        assert tuple(var.shape) == tuple(w.shape), (var.name, var.shape, w.shape)
        var.assign(w.numpy())
    for var, w in zip([adjuster.weights[i] for i in range(16, 20)], W["A"]):
        assert tuple(var.shape) == tuple(w.shape), (var.name, var.shape, w.shape)
        var.assign(w.numpy())


def _make_trainer(ref_dir, args, generator, discriminator, adjuster, tf):
    sys.path.insert(0, ref_dir)
    import eager_trainer as RT                             # the reference's eager_trainer.py, unmodified
    tr = object.__new__(RT.EagerTrainer)                   # skip __init__: no dataset / result dir / git needed
    tr.args, tr.dataset = args, None
    tr.adjuster, tr.discriminator, tr.generator = adjuster, discriminator, generator
    tr.models = [discriminator, generator, adjuster]
    tr.generator_optimizer = tf.compat.v1.train.AdamOptimizer(args.lr, args.beta_1, args.beta_2)
    tr.discriminator_optimizer = tf.compat.v1.train.AdamOptimizer(args.lr, args.beta_1, args.beta_2)
    tr.adjuster_optimizer = tf.compat.v1.train.AdamOptimizer(args.lr)
    tr.part_groups = {"Generator": [range(0, 4), range(4, 8), range(8, 22)],
                      "Discriminator": [range(0, 12), range(12, 16), range(16, 20)],
                      "Adjuster": [range(16, 20)]}
    tr.part_weights = {}
    for m in tr.models:
        name = m.__class__.__name__
        tr.part_weights[name] = [[m.weights[layer] for layer in group] for group in tr.part_groups[name]]
    tr.all_weights = {"Generator": generator.weights, "Discriminator": discriminator.weights,
                      "Adjuster": [adjuster.weights[w] for w in range(16, 20)]}
    return tr


class _Iter:
    def __init__(self, items):
        self.items = list(items)

    def get_next(self):
        return self.items.pop(0)


class _Injected:
    """Replaces the random draws of `_train_step` for the duration of one call: the first tf.random.normal call is
    the generator noise (eager_trainer.py:125), the second the additive image noise (:131) -> zeros; the four
    tf.image.random_* calls (:127-130) -> identity.  new_image is then real_image_1, as in the oracle's default."""

    def __init__(self, tf, noise):
        self.tf, self.noise, self.calls = tf, noise, 0

    def __enter__(self):
        tf = self.tf
        self.saved = (tf.random.normal, tf.image.random_flip_left_right, tf.image.random_brightness,
                      tf.image.random_contrast, tf.image.random_hue)

        def normal(shape, *a, **k):
            self.calls += 1
            return tf.constant(self.noise) if self.calls == 1 else tf.zeros(shape)
        ident = lambda x, *a, **k: x
        tf.random.normal = normal
        tf.image.random_flip_left_right = tf.image.random_brightness = ident
        tf.image.random_contrast = tf.image.random_hue = ident
        return self

    def __exit__(self, *exc):
        tf = self.tf
        (tf.random.normal, tf.image.random_flip_left_right, tf.image.random_brightness, tf.image.random_contrast,
         tf.image.random_hue) = self.saved


def _record_gradients(tr):
    """Wrap the three optimisers' apply_gradients so that the gradients `_train_step` hands them are kept."""
    rec = {}
    for key, opt in (("G", tr.generator_optimizer), ("D", tr.discriminator_optimizer), ("A", tr.adjuster_optimizer)):
        orig = opt.apply_gradients

        def wrapped(grads_and_vars, *a, _orig=orig, _key=key, **k):
            gv = list(grads_and_vars)
            rec[_key] = [(g.numpy(), v.name) for g, v in gv]
            return _orig(gv, *a, **k)
        opt.apply_gradients = wrapped
    return rec


def train_goldens(ref_dir, out_dir):
    import tensorflow as tf
    import torch
    tf.compat.v1.enable_eager_execution()                  # main.py:9
    assert tf.__version__.startswith("1.15"), "requirements.txt pins tensorflow-gpu==1.15.4 (found %s)" % tf.__version__
    from oracle import littlegan_oracle as O
    B, seed = 4, 0
    step_file = {}
    for batch_no in (11, 15):
        args = _small_args(use_partition=True)
        W = O.init_weights(args, seed)
        gen, disc, adj = _build_reference(ref_dir, args, tf)
        _assign(gen, disc, adj, W)
        tr = _make_trainer(ref_dir, args, gen, disc, adj, tf)
        rec = _record_gradients(tr)
        i1, c1, i2, c2, noise = (t.numpy() for t in O.synthetic_batch(args, B, seed=5))
        p = "b%d_" % batch_no
        # per-layer activations of the forward passes at the initial weights (model.py call() outputs)
        for i, m in enumerate(disc.encoder(tf.constant(i1))):
            step_file[p + "enc%d_real1" % (i + 1)] = m.numpy()
        step_file[p + "fake0"] = gen([tf.constant(noise), tf.constant(c2)]).numpy()
        pr, c = disc(tf.constant(i1))
        step_file[p + "pr_real1"], step_file[p + "c_real1"] = pr.numpy(), c.numpy()
        step_file[p + "adj0"] = adj([tf.constant(i1), tf.constant((c2 + 1) * 0.5)]).numpy()
        with _Injected(tf, noise):
            res = tr._train_step(batch_no, _Iter([(tf.constant(i1), tf.constant(c1)),
                                                  (tf.constant(i2), tf.constant(c2))]))
        assert res[0] is True
        step_file[p + "fake_image"], step_file[p + "adj_image"] = res[1].numpy(), res[2].numpy()
        step_file[p + "losses"] = np.array([float(res[3]), float(res[4]), float(res[5])], np.float64)
        for key in "DGA":
            for j, (g, name) in enumerate(rec[key]):
                step_file[p + "grad_%s_%d" % (key, j)] = g
            step_file[p + "grad_%s_names" % key] = np.array([n for _, n in rec[key]])
        for key, ws in (("D", disc.weights), ("G", gen.weights), ("A", [adj.weights[i] for i in range(16, 20)])):
            for j, v in enumerate(ws):
                step_file[p + "new_%s_%d" % (key, j)] = v.numpy()
        if batch_no == 11:
            for key in "DGA":
                for j, w in enumerate(W[key]):
                    step_file["w_%s_%d" % (key, j)] = w.numpy()
            step_file.update(i1=i1, c1=c1, i2=i2, c2=c2, noise=noise)
    step_file["meta"] = np.array(json.dumps({"tf": tf.__version__, "B": B, "weight_seed": seed, "data_seed": 5,
                                             "args": "tests.util.small_args(use_partition=True)"}))
    np.savez_compressed(os.path.join(out_dir, "tf_step.npz"), **step_file)

    # 100-step loss trajectory
    args = _small_args(use_partition=True)
    gen, disc, adj = _build_reference(ref_dir, args, tf)
    _assign(gen, disc, adj, O.init_weights(args, seed))
    tr = _make_trainer(ref_dir, args, gen, disc, adj, tf)
    traj = {"gen": [], "disc": [], "adj": [], "tf": tf.__version__, "B": B, "weight_seed": seed, "data_seed": 1000}
    for b in range(1, 101):
        i1, c1, i2, c2, noise = (t.numpy() for t in O.synthetic_batch(args, B, seed=1000 + b))
        with _Injected(tf, noise):
            res = tr._train_step(b, _Iter([(tf.constant(i1), tf.constant(c1)), (tf.constant(i2), tf.constant(c2))]))
        traj["gen"].append(float(res[3])); traj["disc"].append(float(res[4]))
        traj["adj"].append(None if res[5] is None else float(res[5]))
    with open(os.path.join(out_dir, "tf_trajectory.json"), "w") as f:
        json.dump(traj, f)
    print("wrote tf_step.npz and tf_trajectory.json to", out_dir)


def inception_golden(ref_dir, pb, out_dir):
    """pool_3:0 of the real Inception-2015 graph through the reference's own fid.py session code."""
    import tensorflow as tf
    sys.path.insert(0, ref_dir)
    tf.compat.v1.disable_eager_execution()
    import fid as RF                                       # the reference's fid.py (needs scipy.misc.imread only
    RF.create_inception_graph(pb)                          # at import of its file helpers; see SURVEY 8 c)
    imgs = np.random.RandomState(0).randint(0, 256, (4, 128, 128, 3)).astype(np.float32)
    with tf.compat.v1.Session() as sess:
        sess.run(tf.compat.v1.global_variables_initializer())
        act = RF.get_activations(imgs, sess, batch_size=2)
    np.savez_compressed(os.path.join(out_dir, "tf_pool3.npz"), images=imgs.astype(np.uint8), pool3=act)
    print("wrote tf_pool3.npz to", out_dir)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    ap.add_argument("--inception", default=None, help="classify_image_graph_def.pb of inception-2015-12-05.tgz")
    a = ap.parse_args()
    try:
        import tensorflow  # noqa: F401
    except ImportError:
        sys.exit("TensorFlow 1.15 is required (requirements.txt:2); it is not installable in the build container.")
    if a.inception:
        inception_golden(a.reference, a.inception, a.out)
    else:
        train_goldens(a.reference, a.out)
