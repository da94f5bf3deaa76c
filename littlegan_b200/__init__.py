"""littlegan_b200 - B200-native implementation of the LittleGAN hot path.

Mirrors the reference's Python surface (model.py builders, eager_trainer.EagerTrainer,
config.Arg, fid.*) on top of hand-written sm_100a CUDA kernels reached through the C-ABI in
include/littlegan_b200.h.  There is no CPU path.
"""
__version__ = "0.1.0"
