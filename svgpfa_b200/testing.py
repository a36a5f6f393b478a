"""Glue between the seeded case dicts of ``svgpfa_b200.synthetic`` and the model object
(used by tests, ``bench.py`` and ``__graft_entry__.smoke``)."""
from __future__ import annotations

import numpy as np
import torch

from .kernels import build_kernels
from .synthetic import nested_spikes


def initial_params_from_case(case, device=None):
    t = lambda a: (a.detach().to(dtype=torch.float64) if isinstance(a, torch.Tensor)
                   else torch.tensor(np.asarray(a), dtype=torch.float64, device=device))
    return {
        "posterior_on_latents": {
            "posterior_on_ind_points": {"mean": [t(a) for a in case["m"]],
                                        "cholVecs": [t(a) for a in case["chol_vecs"]]},
            "kernels_matrices_store": {"kernels_params0": [t(a) for a in case["kernel_params"]],
                                       "inducing_points_locs0": [t(a) for a in case["Z"]]}},
        "embedding": {"C0": t(case["C"]), "d0": t(case["d"])}}


def model_from_case(case, device=None, process_group=None, nested=False, check_errors=True, shard_mode="auto",
                    spike_chunks=0, spike_method="auto"):
    """A fully specified ``B200SVLowerBound`` for ``case``.  ``nested=True`` feeds the spikes
    through ``setMeasurements`` (nested python lists, the reference's format) instead of the
    flat fast path."""
    from .model import B200SVLowerBound
    model = B200SVLowerBound(kernels=build_kernels(case["kernel_types"]), device=device,
                             process_group=process_group, check_errors=check_errors, shard_mode=shard_mode)
    model._spike_chunks = spike_chunks
    model.spike_method = spike_method
    model.setInitialParams(initial_params_from_case(case))
    if nested:
        model.setMeasurements(nested_spikes(case))
    else:
        model.setMeasurementsFlat(case["spike_times"], case["spike_counts"])
    tt = lambda a: a if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a))
    model.setELLCalculationParams({"leg_quad_points": tt(case["leg_quad_points"]),
                                   "leg_quad_weights": tt(case["leg_quad_weights"])})
    model.setPriorCovRegParam(case["reg"])
    model.buildKernelsMatrices()
    return model


def set_requires_grad(model, posterior=True, embedding=True, kernels=True, indlocs=True):
    for p in model.getSVPosteriorOnIndPointsParams():
        p.requires_grad_(posterior)
    for p in model.getSVEmbeddingParams():
        p.requires_grad_(embedding)
    for p in model.getKernelsParams():
        p.requires_grad_(kernels)
    for p in model.getIndPointsLocs():
        p.requires_grad_(indlocs)


def grads_as_dict(model):
    """Gradients of the last backward in the key layout of the golden fixtures."""
    K = len(model.getKernelsParams())
    post = model.getSVPosteriorOnIndPointsParams()
    out = {}
    g = lambda p: None if p.grad is None else p.grad.detach().cpu().numpy()
    C, d = model.getSVEmbeddingParams()
    out["grad_C"], out["grad_d"] = g(C), g(d)
    for k in range(K):
        out[f"grad_m_{k}"] = g(post[k])
        out[f"grad_chol_vecs_{k}"] = g(post[K + k])
        out[f"grad_kernel_params_{k}"] = g(model.getKernelsParams()[k])
        out[f"grad_Z_{k}"] = g(model.getIndPointsLocs()[k])
    return out
