"""The UNMODIFIED reference driver ``svGPFA.stats.svEM.SVEM_PyTorch`` (stats/svEM.py:76-294), imported from
/root/reference/src, drives a model object through the duck-typed protocol of SURVEY.md §8b, and
tests/ecm_driver.py -- the restatement the -m gpu tests use on the GPU box, where /root/reference does not
exist -- issues exactly the same calls: identical step logs (bound, niter, nfeval) on twin models.

Build container only (skipped when /root/reference is absent).  The model here is the oracle-backed protocol
object; the CUDA model implements the same protocol and is replayed against the reference's own log in
tests/test_gpu_parity.py::test_config1_svem_replay.
"""
import io
import os
import re
import sys

import pytest
import torch

from conftest import GOLDEN
from svgpfa_b200 import synthetic

sys.path.insert(0, os.path.join(GOLDEN))
import ref_harness  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_harness.reference_available(), reason="needs /root/reference")

KW = dict(max_iter=8, lr=1.0, tolerance_grad=1e-7, tolerance_change=1e-9, line_search_fn="strong_wolfe")
STEPS = ("estep", "mstep_embedding", "mstep_kernels", "mstep_indpointslocs")


def _optim_params(em_max_iter):
    p = {"em_max_iter": em_max_iter, "optim_method": "ecm", "verbose": True}
    for s in STEPS:
        p[f"{s}_estimate"] = True
        p[f"{s}_optim_params"] = dict(KW)
    return p


def test_unmodified_svem_and_ecm_driver_issue_the_same_calls():
    import ecm_driver
    ref_harness.import_reference()
    import svGPFA.stats.svEM as ref_svem
    case, _ = synthetic.load_case(os.path.join(GOLDEN, "tiny_mixed.npz"))
    torch.set_num_threads(1)
    out = io.StringIO()
    hist_ref, _, term, _ = ref_svem.SVEM_PyTorch().maximize(
        model=ecm_driver.OracleModel(case), optim_params=_optim_params(2), method="ecm", out=out)
    assert "Maximum number of iterations" in term.message
    pat = re.compile(r"Iteration (\d+), (\w+) end: ([-\d.eE+naif]+), niter: (\d+), nfeval: (\d+)")
    rows = [pat.match(line).groups() for line in out.getvalue().splitlines() if pat.match(line)]
    hist, log = ecm_driver.maximize(ecm_driver.OracleModel(case), em_max_iter=2, lbfgs_kwargs=KW)
    assert len(rows) == len(log) == 8
    for (it, name, bound, niter, nfeval), row in zip(log, rows):
        assert (it, name, niter, nfeval) == (int(row[0]), row[1], int(row[3]), int(row[4]))
        assert f"{bound:f}" == row[2]                      # same deterministic CPU arithmetic: same printed bound
    assert hist == hist_ref                                # bit-identical lower-bound history


def test_unmodified_svem_runs_on_the_b200_optimiser_through_the_patch():
    """``svgpfa_b200.lbfgs.patched_torch_lbfgs`` swaps the device-resident optimiser in underneath the UNMODIFIED
    ``SVEM_PyTorch`` (which builds ``torch.optim.LBFGS`` by name, svEM.py:221,229,243,262): same step log as the
    unpatched run -- niter, nfeval of all 8 steps, printed bounds to 1e-9.  (CPU: the torch implementation of the vector
    primitives is bound in; on a GPU the default is the CUDA library.)"""
    sys.path.insert(0, os.path.dirname(__file__))
    import ecm_driver
    from vector_ops_torch import TorchVectorOps
    from svgpfa_b200.lbfgs import LBFGS, patched_torch_lbfgs
    ref_harness.import_reference()
    import svGPFA.stats.svEM as ref_svem
    case, _ = synthetic.load_case(os.path.join(GOLDEN, "tiny_mixed.npz"))
    torch.set_num_threads(1)
    pat = re.compile(r"Iteration (\d+), (\w+) end: ([-\d.eE+naif]+), niter: (\d+), nfeval: (\d+)")

    def run():
        out = io.StringIO()
        hist, _, term, _ = ref_svem.SVEM_PyTorch().maximize(
            model=ecm_driver.OracleModel(case), optim_params=_optim_params(2), method="ecm", out=out)
        assert "Maximum number of iterations" in term.message
        return hist, [pat.match(line).groups() for line in out.getvalue().splitlines() if pat.match(line)]
    original = torch.optim.LBFGS
    hist_a, rows_a = run()
    with patched_torch_lbfgs(ops=TorchVectorOps()):
        assert issubclass(torch.optim.LBFGS, LBFGS)
        hist_b, rows_b = run()
    assert torch.optim.LBFGS is original
    assert len(rows_a) == len(rows_b) == 8
    for a, b in zip(rows_a, rows_b):
        assert (a[0], a[1], a[3], a[4]) == (b[0], b[1], b[3], b[4])
        assert float(b[2]) == pytest.approx(float(a[2]), rel=1e-9)
    assert hist_b == pytest.approx(hist_a, rel=1e-9)
