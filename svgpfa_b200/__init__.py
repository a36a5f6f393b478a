"""svgpfa_b200: B200-native (sm_100a) implementation of svGPFA's variational lower bound
(point-process expected log-likelihood with exponential link minus the KL divergence over
inducing points) and its gradients, behind the model protocol svGPFA's own ``SVEM_PyTorch``
drives.  The arithmetic lives in ``libsvgpfa_b200.so`` (hand-written CUDA, C ABI in
``include/svgpfa_b200.h``); there is no CPU fallback.
"""
from .kernels import ExponentialQuadraticKernel, PeriodicKernel, build_kernels  # noqa: F401


def __getattr__(name):
    if name in ("B200SVLowerBound", "buildModelB200"):
        from . import model
        return getattr(model, name)
    raise AttributeError(name)
