"""InstanceNormalization layer object (reference `instance.py:9-144`).

The reference only ever constructs it with the defaults (`axis=None`, `epsilon=1e-3`, scalar
gamma/beta of shape (1,), `model.py:16,41,84,121`): per-sample statistics over ALL non-batch
axes, epsilon added to the standard deviation (`instance.py:107-116`).  That is the only mode
the CUDA kernels implement; other `axis` values raise.
"""
import torch

from . import kernels as K


def _device():
    return torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")


class InstanceNormalization:
    def __init__(self, axis=None, epsilon=1e-3, center=True, scale=True, **kwargs):
        if axis is not None:
            raise NotImplementedError("littlegan_b200 implements InstanceNormalization(axis=None) only "
                                      "(the only form the reference uses)")
        if not (center and scale):
            raise NotImplementedError("center=False / scale=False are not used by the reference")
        self.axis, self.epsilon = axis, float(epsilon)
        self.gamma = torch.ones(1, dtype=torch.float32, device=_device())
        self.beta = torch.zeros(1, dtype=torch.float32, device=_device())

    @property
    def weights(self):
        return [self.gamma, self.beta]

    def __call__(self, inputs, training=None):
        """Stand-alone forward (statistics pass + apply).  inputs: [N, ...] CUDA tensor."""
        x = inputs.contiguous()
        stats = torch.zeros(x.shape[0], 2, dtype=torch.float64, device=x.device)
        K.rowstats(x, stats, 1.0)
        out = torch.empty_like(x)
        return K.instnorm_act_fwd(x, stats, self.gamma, self.beta, None, out, self.epsilon, 1.0, 1.0)
