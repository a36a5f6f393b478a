"""Kernel descriptors with the reference's class and method names
(``/root/reference/src/svGPFA/stats/kernels.py:7-108``) so that scripts written against
``svGPFA.stats.kernels`` run unchanged.  They only carry hyper-parameters and metadata: the
covariance arithmetic of the lower bound happens in the CUDA library.  ``buildKernelMatrix``
is provided for post-fit utilities on small inputs and uses plain tensor ops.
"""
from __future__ import annotations

import math

import torch


class Kernel:
    ktype = None
    n_params = 0

    def getParams(self):
        return self._params

    def setParams(self, params):
        self._params = params

    def buildKernelMatrixDiag(self, X):
        # kappa(t, t) = scale^2, no regulariser (kernels.py:48-51, 87-90)
        return self._scale ** 2 * torch.ones(X.shape, dtype=X.dtype, device=X.device)


class ExponentialQuadraticKernel(Kernel):
    ktype = "expquad"
    n_params = 1

    def __init__(self, scale=1.0, lengthscaleScale=1.0, dtype=torch.double):
        self._scale = torch.tensor(scale, dtype=dtype)
        self._lengthscaleScale = lengthscaleScale

    def buildKernelMatrix(self, X1, X2=None):
        X2 = X1 if X2 is None else X2
        ell = self._params[0] / self._lengthscaleScale
        delta = X1 - X2.transpose(1, 2) if X1.ndim == 3 else X1.reshape(-1, 1) - X2.reshape(1, -1)
        return self._scale.to(delta.device) ** 2 * torch.exp(-0.5 * delta ** 2 / ell ** 2)

    def getScaledParams(self):
        return torch.tensor([self._params[0] / self._lengthscaleScale])

    def getNamedParams(self):
        return {"scale": self._scale, "lengthscale": self._params[0]}


class PeriodicKernel(Kernel):
    ktype = "periodic"
    n_params = 2

    def __init__(self, scale=1.0, lengthscaleScale=1.0, periodScale=1.0, dtype=torch.double):
        self._scale = torch.tensor(scale, dtype=dtype)
        self._lengthscaleScale = lengthscaleScale
        self._periodScale = periodScale

    def buildKernelMatrix(self, X1, X2=None):
        X2 = X1 if X2 is None else X2
        ell = self._params[0] / self._lengthscaleScale
        period = self._params[1] / self._periodScale
        delta = X1 - X2.transpose(1, 2) if X1.ndim == 3 else X1.reshape(-1, 1) - X2.reshape(1, -1)
        return self._scale.to(delta.device) ** 2 * torch.exp(
            -2.0 * torch.sin(math.pi * delta / period) ** 2 / ell ** 2)

    def getScaledParams(self):
        return torch.tensor([self._params[0] / self._lengthscaleScale,
                             self._params[1] / self._periodScale])

    def getNamedParams(self):
        return {"scale": self._scale, "lengthscale": self._params[0], "period": self._params[1]}


def kernel_spec(kernel):
    """(ktype code, scale^2, 1/lengthscaleScale, 1/periodScale) of a kernel object; accepts
    this module's classes and, by duck typing, the reference's own kernel objects."""
    name = type(kernel).__name__
    scale = float(getattr(kernel, "_scale", 1.0))
    ils = 1.0 / float(getattr(kernel, "_lengthscaleScale", 1.0))
    if "Periodic" in name:
        return 1, scale * scale, ils, 1.0 / float(getattr(kernel, "_periodScale", 1.0))
    if "ExponentialQuadratic" in name:
        return 0, scale * scale, ils, 1.0
    raise ValueError(f"unsupported kernel class {name}")


def build_kernels(kernel_types, kernel_params=None):
    """Counterpart of ``svGPFA.utils.miscUtils.buildKernels`` (utils/miscUtils.py:38-50)."""
    out = []
    for k, kt in enumerate(kernel_types):
        kt = kt.lower()
        if kt in ("expquad", "exponentialquadratic", "rbfkernel"):
            kern = ExponentialQuadraticKernel()
        elif kt in ("periodic", "periodickernel"):
            kern = PeriodicKernel()
        else:
            raise ValueError(f"Invalid kernel type {kt}")
        if kernel_params is not None:
            kern.setParams(kernel_params[k])
        out.append(kern)
    return out
