"""Torch (CPU) implementation of the vector primitives ``svgpfa_b200.lbfgs.LBFGS`` is written against.
TEST INFRASTRUCTURE ONLY: lets the optimiser's decision logic run in the CPU suite (and under gloo); the product's
primitives are the CUDA kernels of ``csrc/lbfgs.cu`` (``svgpfa_b200.lbfgs.CudaVectorOps``)."""
import torch


class TorchVectorOps:
    def __init__(self, device="cpu"):
        self.device = torch.device(device)

    def empty(self, n):
        return torch.empty(n, dtype=torch.float64, device=self.device)

    def multidot(self, vecs, probes):
        return torch.stack(vecs) @ torch.stack(probes).T

    def combine(self, d, vecs, coefs, g):
        d.zero_()
        for c, v in zip(coefs, vecs):
            d.add_(v, alpha=float(c))
        return torch.stack([g.dot(d), d.abs().max() if d.numel() else d.sum()])

    def stats(self, a, b=None):
        z = torch.zeros((), dtype=torch.float64, device=self.device)
        return torch.stack([a.dot(b) if b is not None else z, a.abs().max(), a.abs().sum(),
                            b.abs().max() if b is not None else z])

    def update(self, s, y, d, t, g, g_prev):
        torch.mul(d, float(t), out=s)
        torch.sub(g, g_prev, out=y)
        g_prev.copy_(g)

    def step(self, x, x0, d, t):
        torch.add(x0, d, alpha=float(t), out=x)
