"""Runs the UNMODIFIED reference (imported from /root/reference/src) on a case dict.

Only usable in the build container: /root/reference does not exist on the GPU box, so
nothing under tests/ that runs with ``-m gpu`` imports this module.  It is used by
``make_golden.py`` to produce the committed fixtures and by the optional
``test_oracle_vs_reference.py`` (skipped when /root/reference is absent).

``gcnu_common`` is a third-party, un-vendored, un-pinned dependency of the reference
(setup.cfg:21); none of the ELBO arithmetic lives in it.  The in-memory stand-in below
supplies only the names the reference imports at module load
(utils/miscUtils.py:13-14).
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

REF_SRC = "/root/reference/src"


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_SRC, "svGPFA"))


def _install_gcnu_shim():
    if "gcnu_common" in sys.modules:
        return

    def leggaussVarLimits(n, a, b):
        x, w = np.polynomial.legendre.leggauss(n)
        return (torch.from_numpy(0.5 * (b - a) * x + 0.5 * (b + a)),
                torch.from_numpy(0.5 * (b - a) * w))

    names = ["gcnu_common", "gcnu_common.numerical_methods", "gcnu_common.numerical_methods.utils",
             "gcnu_common.stats", "gcnu_common.stats.gaussianProcesses",
             "gcnu_common.stats.gaussianProcesses.eval", "gcnu_common.stats.pointProcesses",
             "gcnu_common.stats.pointProcesses.sampling", "gcnu_common.utils",
             "gcnu_common.utils.config_dict", "gcnu_common.utils.argparse"]
    mods = {n: types.ModuleType(n) for n in names}
    for n, mod in mods.items():
        if "." in n:
            parent, child = n.rsplit(".", 1)
            setattr(mods[parent], child, mod)
    mods["gcnu_common.numerical_methods.utils"].leggaussVarLimits = leggaussVarLimits
    sys.modules.update(mods)


def import_reference():
    if not reference_available():
        raise RuntimeError("reference tree not present")
    _install_gcnu_shim()
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    import svGPFA.stats.kernels  # noqa: F401
    import svGPFA.stats.svGPFAModelFactory  # noqa: F401
    import svGPFA
    return svGPFA


def build_reference_model(case, requires_grad=True):
    """Reference model via its own factory (stats/svGPFAModelFactory.py:40-148)."""
    svGPFA = import_reference()
    import svGPFA.stats.kernels as rk
    import svGPFA.stats.svGPFAModelFactory as rf
    from svgpfa_b200.synthetic import nested_spikes

    K = len(case["kernel_types"])
    kernels = []
    for k in range(K):
        if case["kernel_types"][k] == "periodic":
            kernels.append(rk.PeriodicKernel(scale=1.0))
        else:
            kernels.append(rk.ExponentialQuadraticKernel(scale=1.0))
    t = lambda a: torch.tensor(np.asarray(a), dtype=torch.double)
    params = dict(
        m=[t(a) for a in case["m"]], chol_vecs=[t(a) for a in case["chol_vecs"]],
        C=t(case["C"]), d=t(case["d"]),
        kernel_params=[t(a) for a in case["kernel_params"]], Z=[t(a) for a in case["Z"]],
    )
    initial_params = {
        "posterior_on_latents": {
            "posterior_on_ind_points": {"mean": params["m"], "cholVecs": params["chol_vecs"]},
            "kernels_matrices_store": {"kernels_params0": params["kernel_params"],
                                       "inducing_points_locs0": params["Z"]}},
        "embedding": {"C0": params["C"], "d0": params["d"]}}
    spikes = [[torch.from_numpy(np.ascontiguousarray(s)) for s in trial]
              for trial in nested_spikes(case)]
    model = rf.SVGPFAModelFactory.buildModelPyTorch(kernels=kernels)
    if requires_grad:
        for group in ("m", "chol_vecs", "kernel_params", "Z"):
            for p in params[group]:
                p.requires_grad_(True)
        params["C"].requires_grad_(True)
        params["d"].requires_grad_(True)
    model.setParamsAndData(
        measurements=spikes, initial_params=initial_params,
        eLLCalculationParams={"leg_quad_points": t(case["leg_quad_points"]),
                              "leg_quad_weights": t(case["leg_quad_weights"])},
        priorCovRegParam=case["reg"])
    return model, params


def reference_outputs(case, with_stats=True):
    """ELBO, ELL, KL, every gradient of the ELBO, the spike stacking, and (optionally)
    the latent / embedding statistics the reference's unit tests pin."""
    model, params = build_reference_model(case, requires_grad=True)
    model.buildKernelsMatrices()
    ell = model._eLL.evalSumAcrossTrialsAndNeurons()
    kl = model._klDiv.evalSumAcrossLatentsAndTrials()
    elbo = model.eval()
    elbo.backward()
    K = len(case["kernel_types"])
    out = {"elbo": elbo.item(), "ell": ell.item(), "kl": kl.item(),
           "grad_C": params["C"].grad.numpy(), "grad_d": params["d"].grad.numpy()}
    for k in range(K):
        out[f"grad_m_{k}"] = params["m"][k].grad.numpy()
        out[f"grad_chol_vecs_{k}"] = params["chol_vecs"][k].grad.numpy()
        out[f"grad_kernel_params_{k}"] = params["kernel_params"][k].grad.numpy()
        out[f"grad_Z_{k}"] = params["Z"][k].grad.numpy()
    assoc = model._eLL._svEmbeddingAssocTimes
    out["stacked_neuron_index"] = np.concatenate(
        [np.asarray(ix.numpy(), dtype=np.int64).reshape(-1) for ix in assoc._neuronForSpikeIndex])
    out["stacked_trial_offsets"] = np.concatenate(
        [[0], np.cumsum([len(ix) for ix in assoc._neuronForSpikeIndex])]).astype(np.int64)
    if with_stats:
        with torch.no_grad():
            stats = model.computeSVPosteriorOnLatentsStats()
            out["quad_latent_mean"] = stats["allTimes"][0].numpy()
            out["quad_latent_var"] = stats["allTimes"][1].numpy()
            out["spike_latent_mean"] = np.concatenate([a.numpy() for a in stats["assocTimes"][0]], 0)
            out["spike_latent_var"] = np.concatenate([a.numpy() for a in stats["assocTimes"][1]], 0)
            ell_cached = model.evalELLSumAcrossTrialsAndNeurons(svPosteriorOnLatentsStats=stats)
            out["ell_cached"] = ell_cached.item()
            eq_mean, eq_var = model._eLL._svEmbeddingAllTimes.computeMeansAndVars()
            out["quad_embedding_mean"] = eq_mean.numpy()
            out["quad_embedding_var"] = eq_var.numpy()
            es_mean, _ = assoc.computeMeansAndVars()
            out["spike_embedding_mean"] = np.concatenate([a.numpy().reshape(-1) for a in es_mean])
    return out
