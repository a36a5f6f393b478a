"""Hand-derived adjoints (oracle/analytic_np.py, the decomposition the CUDA kernels use,
explicit L^-1) against the reference's own autograd outputs in the golden fixtures.
CPU only.  Tolerances are BASELINE.json's product tolerances tightened 10x:
1e-11 on the bound, 1e-9 on every gradient tensor."""
import os

import pytest

from conftest import GOLDEN, golden_names, rel_err
from oracle import analytic_np
from svgpfa_b200 import synthetic


@pytest.mark.parametrize("name", golden_names())
def test_analytic_matches_reference(name):
    case, ref = synthetic.load_case(os.path.join(GOLDEN, name + ".npz"))
    out = analytic_np.elbo_and_grads(case)
    for key in ("elbo", "ell", "kl"):
        assert abs(out[key] - float(ref[key])) <= 1e-11 * abs(float(ref[key])), key
    K = len(case["kernel_types"])
    keys = ["grad_C", "grad_d"]
    for k in range(K):
        keys += [f"grad_m_{k}", f"grad_chol_vecs_{k}", f"grad_kernel_params_{k}", f"grad_Z_{k}"]
    worst = max((rel_err(out[key], ref[key]), key) for key in keys)
    assert worst[0] <= 1e-9, worst
    if "quad_latent_mean" in ref:
        for key in ("quad_latent_mean", "quad_latent_var", "spike_latent_mean"):
            assert rel_err(out[key], ref[key]) <= 1e-10, key
