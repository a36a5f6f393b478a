"""Fixture for the rank-1-plus-diagonal parameterisation of the variational covariance (SURVEY.md §8f-4,
stats/svPosteriorOnIndPoints.py:86-119: S_kr = q q^T + diag(d^2)), produced by the UNMODIFIED reference built with
``indPointsCovRep=indPointsCovRank1PlusDiag`` (stats/svGPFAModelFactory.py:32,65-67).

    python tests/golden/make_rank1.py        ->  tests/golden/tiny_rank1.npz

Inputs: the "tiny" mixed-kernel problem with random (q, d); ``chol_vecs`` of the saved case are the Cholesky vectors of
the same covariances, so the standard model on this case has the same bound.  Outputs: bound, ELL, KL and the gradient
with respect to every parameter group, q and d included.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, HERE)

from svgpfa_b200 import synthetic  # noqa: E402
import ref_harness  # noqa: E402


def main():
    ref_harness.import_reference()
    import svGPFA.stats.kernels as rk
    import svGPFA.stats.svGPFAModelFactory as rf
    case = synthetic.make_case("tiny", seed=3)
    K = len(case["kernel_types"])
    rng = np.random.default_rng(11)
    q = [0.3 * rng.standard_normal(np.asarray(case["m"][k]).shape) for k in range(K)]
    d = [0.2 + 0.3 * rng.random(np.asarray(case["m"][k]).shape) for k in range(K)]
    # Cholesky vectors of the same covariances (row-major tril order, miscUtils.py:135-139)
    chol = []
    for k in range(K):
        R, M, _ = q[k].shape
        S = q[k] @ q[k].transpose(0, 2, 1) + np.stack([np.diag(d[k][r, :, 0] ** 2) for r in range(R)])
        L = np.linalg.cholesky(S)
        ti = np.tril_indices(M)
        chol.append(L[:, ti[0], ti[1]][:, :, None])
    case["chol_vecs"] = chol
    t = lambda a: torch.tensor(np.asarray(a), dtype=torch.double)
    params = dict(m=[t(a) for a in case["m"]], q=[t(a) for a in q], dg=[t(a) for a in d], C=t(case["C"]), d=t(case["d"]),
                  kernel_params=[t(a) for a in case["kernel_params"]], Z=[t(a) for a in case["Z"]])
    for group in ("m", "q", "dg", "kernel_params", "Z"):
        for p in params[group]:
            p.requires_grad_(True)
    params["C"].requires_grad_(True)
    params["d"].requires_grad_(True)
    kernels = [rk.PeriodicKernel(scale=1.0) if kt == "periodic" else rk.ExponentialQuadraticKernel(scale=1.0)
               for kt in case["kernel_types"]]
    model = rf.SVGPFAModelFactory.buildModelPyTorch(kernels=kernels, indPointsCovRep=rf.indPointsCovRank1PlusDiag)
    initial_params = {
        "posterior_on_latents": {
            "posterior_on_ind_points": {"mean": params["m"], "qSVec0": params["q"], "qSDiag0": params["dg"]},
            "kernels_matrices_store": {"kernels_params0": params["kernel_params"], "inducing_points_locs0": params["Z"]}},
        "embedding": {"C0": params["C"], "d0": params["d"]}}
    spikes = [[torch.from_numpy(np.ascontiguousarray(s)) for s in trial] for trial in synthetic.nested_spikes(case)]
    model.setParamsAndData(measurements=spikes, initial_params=initial_params,
                           eLLCalculationParams={"leg_quad_points": t(case["leg_quad_points"]),
                                                 "leg_quad_weights": t(case["leg_quad_weights"])},
                           priorCovRegParam=case["reg"])
    model.buildKernelsMatrices()
    ell = model._eLL.evalSumAcrossTrialsAndNeurons()
    kl = model._klDiv.evalSumAcrossLatentsAndTrials()
    elbo = model.eval()
    elbo.backward()
    got = model.getSVPosteriorOnIndPointsParams()
    assert len(got) == 3 * K and got[K] is params["q"][0] and got[2 * K] is params["dg"][0]      # mean, qSVec, qSDiag
    out = {"elbo": elbo.item(), "ell": ell.item(), "kl": kl.item(),
           "grad_C": params["C"].grad.numpy(), "grad_d": params["d"].grad.numpy()}
    for k in range(K):
        out[f"in_q_svec_{k}"] = q[k]
        out[f"in_q_sdiag_{k}"] = d[k]
        out[f"grad_m_{k}"] = params["m"][k].grad.numpy()
        out[f"grad_q_svec_{k}"] = params["q"][k].grad.numpy()
        out[f"grad_q_sdiag_{k}"] = params["dg"][k].grad.numpy()
        out[f"grad_kernel_params_{k}"] = params["kernel_params"][k].grad.numpy()
        out[f"grad_Z_{k}"] = params["Z"][k].grad.numpy()
    path = os.path.join(HERE, "tiny_rank1.npz")
    synthetic.save_case(path, case, extra=out)
    print(path, out["elbo"], out["ell"], out["kl"])


if __name__ == "__main__":
    main()
