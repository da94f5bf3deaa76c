"""Shared helpers for the parity tests (oracle = checker only)."""
import torch

from oracle import littlegan_oracle as O


def rel_err(got, ref):
    got = got.detach().double().cpu()
    ref = ref.detach().double().cpu()
    return float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


def tol(dtype):
    # north_star: 1e-4 relative in fp32 mode, 2e-2 in bf16 mode
    return 1e-4 if dtype == torch.float32 else 2e-2


def small_args(**over):
    """Reduced architecture (32x32 images) that keeps every structural feature of the model."""
    base = dict(init_dim=2, conv_filter=[48, 32, 16, 8, 8], cond_dim=5, noise_dim=11, batch_size=4,
                use_partition=False, train_adj=True)
    base.update(over)
    return O.make_args(**base)


def product_args(oargs, **over):
    """The product-side Arg carrying the same hyper-parameters as an oracle args namespace."""
    from littlegan_b200.config import Arg
    keys = ["batch_size", "image_channel", "noise_dim", "init_dim", "conv_filter", "kernel_size", "leaky_alpha",
            "dropout_rate", "l1_lambda", "lr", "beta_1", "beta_2", "use_gp", "use_clip", "clip_range",
            "use_partition", "partition_interval", "train_adj"]
    d = {k: getattr(oargs, k) for k in keys}
    d["attr"] = list(range(oargs.cond_dim))
    d["image_dim"] = oargs.init_dim * 16
    d["augment"] = False          # oracle comparisons: new_image = real_image_1 unless a test injects one
    d.update(over)
    return Arg.from_dict(**d)


def build_product(pargs, seed=0):
    from littlegan_b200 import model as M
    M.set_init_seed(seed)
    dec, enc = M.Decoder(pargs), M.Encoder(pargs)
    gen = M.Generator(pargs, dec)
    disc = M.Discriminator(pargs, enc)
    adj = M.Adjuster(pargs, disc, gen)
    return gen, disc, adj


def product_weights_to_oracle(gen, disc, adj):
    cp = lambda ws: [w.detach().cpu().clone() for w in ws]
    return dict(D=cp(disc.weights), G=cp(gen.weights), A=cp(adj.weights[16:20]))
