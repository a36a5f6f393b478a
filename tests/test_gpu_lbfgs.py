"""The device-resident L-BFGS on the GPU: its CUDA vector primitives (csrc/lbfgs.cu, through the C ABI)
against torch, and ``svgpfa_b200.lbfgs.LBFGS`` in place of ``torch.optim.LBFGS`` under the ECM driver -- on the
reference's own example (config #1) it must reproduce the step log of the UNMODIFIED reference's SVEM_PyTorch."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(__file__))
from conftest import GOLDEN  # noqa: E402
from svgpfa_b200 import ecm, synthetic  # noqa: E402

pytestmark = pytest.mark.gpu
F64 = torch.float64


def _ops():
    from svgpfa_b200.lbfgs import CudaVectorOps
    return CudaVectorOps(torch.device("cuda:0"))


@pytest.mark.parametrize("n", [1, 2, 7, 1000, 262145, 3_000_001])
@pytest.mark.parametrize("nv,npb", [(1, 1), (3, 3), (8, 2), (21, 3), (64, 3), (70, 3)])
def test_multidot_and_combine(n, nv, npb):
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(n + nv)
    vecs = [torch.randn(n, dtype=F64, device="cuda", generator=g) for _ in range(nv)]
    probes = [vecs[0]] + [torch.randn(n, dtype=F64, device="cuda", generator=g) for _ in range(npb - 1)]
    out = ops.multidot(vecs, probes).cpu()
    ref = (torch.stack(vecs) @ torch.stack(probes).T).cpu()
    scale = torch.stack(vecs).abs().cpu() @ torch.stack(probes).abs().T.cpu()
    assert torch.all((out - ref).abs() <= 1e-14 * scale + 1e-300)
    assert torch.equal(out, ops.multidot(vecs, probes).cpu())          # block-ordered partial sums: reproducible
    coefs = np.linspace(-1.0, 1.0, nv).tolist()
    d = ops.empty(n)
    gd = ops.combine(d, vecs, coefs, probes[-1]).cpu()
    dref = sum(c * v for c, v in zip(coefs, vecs))
    bound = 1e-14 * sum(abs(c) * v.abs() for c, v in zip(coefs, vecs)) + 1e-300
    assert torch.all((d - dref).abs() <= bound)
    assert abs(gd[0] - probes[-1].dot(d).cpu()) <= 1e-13 * (probes[-1].abs() * d.abs()).sum().cpu() + 1e-300
    assert gd[1] == d.abs().max().cpu()


@pytest.mark.parametrize("n,offset", [(1, 0), (5, 1), (4096, 0), (100003, 1), (100003, 2)])
def test_stats_update_step(n, offset):
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(n)
    a, b, d = (torch.randn(n, dtype=F64, device="cuda", generator=g) for _ in range(3))
    st = ops.stats(a, b).cpu()
    assert abs(st[0] - a.dot(b).cpu()) <= 1e-13 * (a.abs() * b.abs()).sum().cpu()
    assert st[1] == a.abs().max().cpu() and st[3] == b.abs().max().cpu()
    assert abs(st[2] - a.abs().sum().cpu()) <= 1e-13 * a.abs().sum().cpu()
    st1 = ops.stats(a).cpu()
    assert st1[0] == 0 and st1[3] == 0 and st1[1] == st[1]
    s, y, gp = ops.empty(n), ops.empty(n), b.clone()
    ops.update(s, y, d, 0.375, a, gp)
    assert torch.equal(s, 0.375 * d) and torch.equal(y, a - b) and torch.equal(gp, a)
    # the trial point is written into a slice of a larger (packed) buffer: any 8-byte alignment
    buf = torch.zeros(n + 4, dtype=F64, device="cuda")
    ops.step(buf[offset:offset + n], a, d, -0.25)
    assert torch.allclose(buf[offset:offset + n], a - 0.25 * d, rtol=1e-15, atol=0)
    assert buf[:offset].abs().sum() == 0 and buf[offset + n:].abs().sum() == 0


def test_misaligned_vectors_are_refused():
    ops = _ops()
    x = torch.zeros(9, dtype=F64, device="cuda")
    with pytest.raises(RuntimeError, match="16-byte"):
        ops.multidot([x[1:]], [x[1:]])


KW = dict(max_iter=8, lr=1.0, tolerance_grad=1e-7, tolerance_change=1e-9, line_search_fn="strong_wolfe")


def _optim_params(em_max_iter, kw):
    p = {"em_max_iter": em_max_iter}
    for s in ecm.STEP_ORDER["ecm"]:
        p[f"{s}_estimate"] = True
        p[f"{s}_optim_params"] = dict(kw)
    return p


@pytest.mark.parametrize("line_search_fn", ["strong_wolfe", None])
def test_ecm_with_the_b200_optimiser_follows_torch_lbfgs_on_the_gpu_model(line_search_fn):
    """Same model, same kernels, the two optimisers: same niter / nfeval in each of the 8 steps, bounds to 1e-8."""
    from svgpfa_b200.testing import model_from_case
    case, _ = synthetic.load_case(os.path.join(GOLDEN, "tiny_mixed.npz"))
    kw = dict(KW, line_search_fn=line_search_fn, lr=1.0 if line_search_fn else 1e-3)
    logs = {}
    for opt in ("torch", "b200"):
        model = model_from_case(case, spike_method="direct")
        hist, _, msg, log = ecm.maximize(model, _optim_params(2, kw), out=None, optimizer=opt)
        assert "Maximum number of iterations" in msg
        logs[opt] = (hist, log)
    (ha, la), (hb, lb) = logs["torch"], logs["b200"]
    assert [r[:2] + r[3:] for r in la] == [r[:2] + r[3:] for r in lb]
    assert [r[2] for r in lb] == pytest.approx([r[2] for r in la], rel=1e-8)
    assert hb == pytest.approx(ha, rel=1e-8)


def test_config1_svem_replay_with_the_b200_optimiser():
    """BASELINE.json config #1 (the reference's smoke test on its shipped data) through the product's ECM driver with
    the device-resident L-BFGS: the step log of the unmodified reference's SVEM_PyTorch (bounds to 1e-7, niter / nfeval
    with the allowance of tests/test_config1_example.py::check_step_log), and far fewer host reads than closure calls
    would cost torch.optim.LBFGS."""
    from test_config1_example import LBFGS_545, check_step_log
    from svgpfa_b200 import B200SVLowerBound, build_kernels, lbfgs
    from svgpfa_b200.testing import initial_params_from_case
    case, ref = synthetic.load_case(os.path.join(GOLDEN, "config1_example.npz"))
    measurements = [[torch.from_numpy(np.ascontiguousarray(s)) for s in trial] for trial in synthetic.nested_spikes(case)]
    # (up to three attempts: the kernels' FP64 atomics make repeated runs differ in the last bits and this example has
    #  three steps that stop on borderline tolerance tests -- see tests/test_gpu_parity.py::test_config1_svem_replay)
    last = None
    for attempt in range(3):
        model = B200SVLowerBound(kernels=build_kernels(case["kernel_types"]))
        model.spike_method = "direct"
        model.setParamsAndData(
            measurements=measurements, initial_params=initial_params_from_case(case),
            eLLCalculationParams={"leg_quad_points": torch.from_numpy(case["leg_quad_points"]),
                                  "leg_quad_weights": torch.from_numpy(case["leg_quad_weights"])},
            priorCovRegParam=case["reg"])
        made = []

        def factory(params, **kw):
            made.append(lbfgs.LBFGS(params, **kw))
            return made[-1]
        hist, _, msg, log = ecm.maximize(model, _optim_params(2, LBFGS_545), out=None, optimizer=factory)
        assert "Maximum number of iterations" in msg
        try:
            check_step_log(log, ref["svem_step_log"], exact=False)
            assert hist[1:] == pytest.approx(ref["svem_lower_bound_hist"][1:].tolist(), rel=1e-7)
            last = None
            break
        except AssertionError as e:
            last = e
    if last is not None:
        raise last
    for opt, row in zip(made, log):
        nfeval, niter = row[4], row[3]
        assert opt.host_reads <= nfeval + 2 * niter + 1          # one read per closure call, two per iteration
    # the E-step leaves (K means, K cholVecs) are two runs of adjacent views of the packed buffers: two segments
    assert len(made[0]._segments) == 2 and len(made[0]._params) == 2 * len(case["kernel_types"])


def test_patched_torch_lbfgs_context():
    """The unmodified reference builds ``torch.optim.LBFGS`` by name (stats/svEM.py:221,229,243,262): the context manager
    swaps the device-resident optimiser in for the duration of a ``maximize`` call and restores torch's afterwards."""
    from svgpfa_b200 import lbfgs
    original = torch.optim.LBFGS
    with lbfgs.patched_torch_lbfgs():
        assert torch.optim.LBFGS is lbfgs.LBFGS
        x = torch.ones(4, dtype=F64, device="cuda", requires_grad=True)
        opt = torch.optim.LBFGS([x], line_search_fn="strong_wolfe")

        def closure():
            opt.zero_grad()
            loss = ((x - 3.0) ** 2).sum()
            loss.backward()
            return loss
        opt.step(closure)
        assert torch.allclose(x.detach(), torch.full_like(x, 3.0), atol=1e-6)
    assert torch.optim.LBFGS is original
