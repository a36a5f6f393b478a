"""Fixture for the pseudo-inverse Kzz store (stats/kernelsMatricesStore.py:146-159:
IndPointsLocsKMS_PInv, Kzz^-1 applied as torch.linalg.pinv(Kzz, rcond=1e-15) @ x), produced by the UNMODIFIED reference
built with ``kernelMatrixInvMethod=kernelMatrixInvPInv`` (stats/svGPFAModelFactory.py:27,73-75) on the inputs of
``tiny_mixed.npz``.

    python tests/golden/make_pinv.py        ->  tests/golden/tiny_mixed_pinv.npz
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
    case, _ = synthetic.load_case(os.path.join(HERE, "tiny_mixed.npz"))
    K = len(case["kernel_types"])
    t = lambda a: torch.tensor(np.asarray(a), dtype=torch.double)
    params = dict(m=[t(a) for a in case["m"]], chol_vecs=[t(a) for a in case["chol_vecs"]], C=t(case["C"]), d=t(case["d"]),
                  kernel_params=[t(a) for a in case["kernel_params"]], Z=[t(a) for a in case["Z"]])
    for group in ("m", "chol_vecs", "kernel_params", "Z"):
        for p in params[group]:
            p.requires_grad_(True)
    params["C"].requires_grad_(True)
    params["d"].requires_grad_(True)
    kernels = [rk.PeriodicKernel(scale=1.0) if kt == "periodic" else rk.ExponentialQuadraticKernel(scale=1.0)
               for kt in case["kernel_types"]]
    model = rf.SVGPFAModelFactory.buildModelPyTorch(kernels=kernels, kernelMatrixInvMethod=rf.kernelMatrixInvPInv)
    initial_params = {
        "posterior_on_latents": {
            "posterior_on_ind_points": {"mean": params["m"], "cholVecs": params["chol_vecs"]},
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
    conds = [float(torch.linalg.cond(Kk).max()) for Kk in model._klDiv.get_indPointsLocsKMS().getKzz()]
    out = {"elbo": elbo.item(), "ell": ell.item(), "kl": kl.item(), "kzz_cond_max": np.array(max(conds)),
           "grad_C": params["C"].grad.numpy(), "grad_d": params["d"].grad.numpy()}
    for k in range(K):
        out[f"grad_m_{k}"] = params["m"][k].grad.numpy()
        out[f"grad_chol_vecs_{k}"] = params["chol_vecs"][k].grad.numpy()
        out[f"grad_kernel_params_{k}"] = params["kernel_params"][k].grad.numpy()
        out[f"grad_Z_{k}"] = params["Z"][k].grad.numpy()
    path = os.path.join(HERE, "tiny_mixed_pinv.npz")
    synthetic.save_case(path, case, extra=out)
    print(path, out["elbo"], "cond(Kzz) <=", max(conds))


if __name__ == "__main__":
    main()
