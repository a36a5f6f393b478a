"""SURVEY.md 8f-4, one of the reference's model variants: the variational covariance parameterised as
S = q q^T + diag(d^2) (SVPosteriorOnIndPointsRank1PlusDiag, stats/svPosteriorOnIndPoints.py:86-119; selected by
``buildModelPyTorch(indPointsCovRep=indPointsCovRank1PlusDiag)``, stats/svGPFAModelFactory.py:32,65-67) on the CUDA
path, against a fixture produced by the UNMODIFIED reference (tests/golden/make_rank1.py)."""
import os
import pickle

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_err
from svgpfa_b200 import synthetic

pytestmark = pytest.mark.gpu
ELBO_TOL, GRAD_TOL = 1e-10, 1e-8


def _model(case, ref):
    import svgpfa_b200
    from svgpfa_b200.model import indPointsCovRank1PlusDiag
    from svgpfa_b200.testing import initial_params_from_case
    K = len(case["kernel_types"])
    model = svgpfa_b200.buildModelB200(kernels=svgpfa_b200.build_kernels(case["kernel_types"]),
                                       indPointsCovRep=indPointsCovRank1PlusDiag)
    ip = initial_params_from_case(case)
    post = ip["posterior_on_latents"]["posterior_on_ind_points"]
    del post["cholVecs"]
    post["qSVec0"] = [torch.tensor(ref[f"in_q_svec_{k}"], dtype=torch.float64) for k in range(K)]
    post["qSDiag0"] = [torch.tensor(ref[f"in_q_sdiag_{k}"], dtype=torch.float64) for k in range(K)]
    measurements = [[torch.from_numpy(np.ascontiguousarray(s)) for s in trial] for trial in synthetic.nested_spikes(case)]
    model.setParamsAndData(measurements=measurements, initial_params=ip,
                           eLLCalculationParams={"leg_quad_points": torch.from_numpy(case["leg_quad_points"]),
                                                 "leg_quad_weights": torch.from_numpy(case["leg_quad_weights"])},
                           priorCovRegParam=case["reg"])
    return model


def test_bound_and_every_gradient_match_the_reference():
    case, ref = synthetic.load_case(os.path.join(GOLDEN, "tiny_rank1.npz"))
    K = len(case["kernel_types"])
    model = _model(case, ref)
    post = model.getSVPosteriorOnIndPointsParams()
    assert len(post) == 3 * K                                        # mean, qSVec, qSDiag (svPosteriorOnIndPoints.py:96-101)
    for p in post + model.getSVEmbeddingParams() + model.getKernelsParams() + model.getIndPointsLocs():
        p.requires_grad_(True)
    v = model.eval()
    v.backward()
    assert abs(v.item() - float(ref["elbo"])) <= ELBO_TOL * abs(float(ref["elbo"]))
    g = lambda p: p.grad.detach().cpu().numpy()
    C, d = model.getSVEmbeddingParams()
    assert rel_err(g(C), ref["grad_C"]) <= GRAD_TOL and rel_err(g(d), ref["grad_d"]) <= GRAD_TOL
    for k in range(K):
        assert rel_err(g(post[k]), ref[f"grad_m_{k}"]) <= GRAD_TOL
        assert rel_err(g(post[K + k]), ref[f"grad_q_svec_{k}"]) <= GRAD_TOL
        assert rel_err(g(post[2 * K + k]), ref[f"grad_q_sdiag_{k}"]) <= GRAD_TOL
        assert rel_err(g(model.getKernelsParams()[k]), ref[f"grad_kernel_params_{k}"]) <= GRAD_TOL
        assert rel_err(g(model.getIndPointsLocs()[k]), ref[f"grad_Z_{k}"]) <= GRAD_TOL
    # the same covariances through the Cholesky-vector parameterisation: the same bound
    from svgpfa_b200.testing import model_from_case
    assert abs(model_from_case(case).eval().item() - v.item()) <= ELBO_TOL * abs(v.item())


def test_estep_on_q_and_d_improves_the_bound_and_survives_pickling():
    from svgpfa_b200 import ecm
    case, ref = synthetic.load_case(os.path.join(GOLDEN, "tiny_rank1.npz"))
    model = _model(case, ref)
    b0 = model.eval().item()
    kw = dict(max_iter=8, lr=1.0, tolerance_grad=1e-7, tolerance_change=1e-9, line_search_fn="strong_wolfe")
    b1, niter, nfeval = ecm.run_step(model, "estep", kw, optimizer="b200")
    assert b1 > b0 and niter >= 1
    stats = model.computeSVPosteriorOnLatentsStats()                 # derived Cholesky vectors follow (q, d)
    twin = pickle.loads(pickle.dumps(model))
    assert abs(twin.eval().item() - b1) <= ELBO_TOL * abs(b1)
    b2, _, _ = ecm.run_step(twin, "estep", kw)                       # the reloaded leaves still drive the kernels
    assert b2 >= b1 - 1e-9 * abs(b1)
    assert stats["allTimes"][0].shape[0] == len(case["spike_counts"])


def test_pinv_kzz_store_variant_matches_the_reference():
    """buildModelB200(kernelMatrixInvMethod=kernelMatrixInvPInv) against the unmodified reference built with
    IndPointsLocsKMS_PInv (tests/golden/make_pinv.py): full-rank Kzz, same kernels, BASELINE.json's tolerances."""
    import svgpfa_b200
    from svgpfa_b200.model import kernelMatrixInvPInv
    from svgpfa_b200.testing import initial_params_from_case, set_requires_grad, grads_as_dict
    case, ref = synthetic.load_case(os.path.join(GOLDEN, "tiny_mixed_pinv.npz"))
    model = svgpfa_b200.buildModelB200(kernels=svgpfa_b200.build_kernels(case["kernel_types"]),
                                       kernelMatrixInvMethod=kernelMatrixInvPInv)
    measurements = [[torch.from_numpy(np.ascontiguousarray(s)) for s in trial] for trial in synthetic.nested_spikes(case)]
    model.setParamsAndData(measurements=measurements, initial_params=initial_params_from_case(case),
                           eLLCalculationParams={"leg_quad_points": torch.from_numpy(case["leg_quad_points"]),
                                                 "leg_quad_weights": torch.from_numpy(case["leg_quad_weights"])},
                           priorCovRegParam=case["reg"])
    set_requires_grad(model)
    v = model.eval()
    v.backward()
    assert abs(v.item() - float(ref["elbo"])) <= ELBO_TOL * abs(float(ref["elbo"]))
    for key, g in grads_as_dict(model).items():
        assert rel_err(g, ref[key]) <= GRAD_TOL, key
    with pytest.raises(ValueError):
        svgpfa_b200.buildModelB200(kernels=svgpfa_b200.build_kernels(case["kernel_types"]), kernelMatrixInvMethod=3)
