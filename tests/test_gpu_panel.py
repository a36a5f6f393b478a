"""The panel path of the spike-time term (include/svgpfa_b200.h: SVGPFA_SPIKE_PANEL; svgpfa_b200/csrc/panel.cu) against
the reference fixtures and against the direct kernel: same tolerances as every other parity test (1e-10 on the bound,
1e-8 on every gradient tensor)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_err
from svgpfa_b200 import synthetic

pytestmark = pytest.mark.gpu

ELBO_TOL, GRAD_TOL = 1e-10, 1e-8


def _grad_keys(K):
    keys = ["grad_C", "grad_d"]
    for k in range(K):
        keys += [f"grad_m_{k}", f"grad_chol_vecs_{k}", f"grad_kernel_params_{k}", f"grad_Z_{k}"]
    return keys


def _eval(case, method, nested=False):
    from svgpfa_b200 import _cabi
    from svgpfa_b200.testing import grads_as_dict, model_from_case, set_requires_grad
    model = model_from_case(case, nested=nested, spike_method=method)
    set_requires_grad(model)
    v = model.eval()
    v.backward()
    out = grads_as_dict(model)
    out["elbo"] = v.item()
    assert model._dims.spike_method == (_cabi.SPIKE_PANEL if method == "panel" else _cabi.SPIKE_DIRECT)
    return model, out


@pytest.mark.parametrize("name", ["tiny_mixed", "tiny_f32", "tiny_empty", "tiny_reg1e-5", "config2_r8", "config3_r4",
                                  "config4_r3", "config1_example"])
def test_panel_path_matches_reference(name):
    """Fixtures of the unmodified reference; trials of 1 s, 8-12 panels for the shortest length scale (0.1) /
    period (0.55) of the synthetic configurations, 4 for the example data (length scales 2 and 1)."""
    case, ref = synthetic.load_case(os.path.join(GOLDEN, name + ".npz"))
    model, out = _eval(case, "panel", nested=name.startswith("tiny"))
    assert abs(out["elbo"] - float(ref["elbo"])) <= ELBO_TOL * abs(float(ref["elbo"])), (out["elbo"], float(ref["elbo"]))
    worst = max((rel_err(out[key], ref[key]), key) for key in _grad_keys(len(case["kernel_types"])))
    assert worst[0] <= GRAD_TOL, worst


def test_panel_and_direct_kernels_agree_on_stage_outputs():
    """Config-#5 shape on a few trials: every buffer the spike stage produces, panel against direct -- far tighter
    than the parity tolerance (the interpolation is accurate to ~3e-14 of each sum)."""
    cfg = dict(synthetic.CONFIGS["config5"], R=6)
    case = synthetic.make_case(cfg, seed=2)
    md, od = _eval(case, "direct")
    mp, op = _eval(case, "panel")
    assert mp._dims.pm_B == 8
    torch.cuda.synchronize()
    a, b = md._ws["abar_spk"].cpu().numpy(), mp._ws["abar_spk"].cpu().numpy()
    assert rel_err(b, a) <= 1e-12
    assert abs(op["elbo"] - od["elbo"]) <= 1e-12 * abs(od["elbo"])
    for key in _grad_keys(cfg["K"]):
        assert rel_err(op[key], od[key]) <= 1e-10, key


def test_automatic_choice_and_fallback():
    """The model picks the panel path when it pays (many spikes per trial, resolvable length scales) and the direct
    kernel otherwise: few spikes, a length scale that would need more than 32 panels, or long trials."""
    from svgpfa_b200 import _cabi
    from svgpfa_b200.testing import model_from_case
    big = synthetic.make_case(dict(synthetic.CONFIGS["config5"], R=4), seed=1)
    m = model_from_case(big)
    m.eval()
    assert m._dims.spike_method == _cabi.SPIKE_PANEL and m._dims.pm_B == 8
    # shrinking a length scale raises the panel count, then forces the direct kernel
    with torch.no_grad():
        m.getKernelsParams()[0].fill_(0.04)
    m.buildKernelsMatrices()
    v_panel = m.eval().item()
    assert m._dims.spike_method == _cabi.SPIKE_PANEL and m._dims.pm_B == 24
    ref = model_from_case(big, spike_method="direct")
    with torch.no_grad():
        ref.getKernelsParams()[0].fill_(0.04)
    assert abs(v_panel - ref.eval().item()) <= 1e-11 * abs(v_panel)
    with torch.no_grad():
        m.getKernelsParams()[0].fill_(0.01)
    m.buildKernelsMatrices()
    m.eval()
    assert m._dims.spike_method == _cabi.SPIKE_DIRECT
    # the MATLAB fixture: 20 s trials, periods 1.5 / 1.2 -> hundreds of panels -> direct
    case, _ = synthetic.load_case(os.path.join(GOLDEN, "matlab_r5.npz"))
    mm = model_from_case(case)
    mm.eval()
    assert mm._dims.spike_method == _cabi.SPIKE_DIRECT
    with pytest.raises(RuntimeError, match="panels"):
        model_from_case(case, spike_method="panel").eval()
    small = model_from_case(synthetic.make_case("tiny", seed=1))
    small.eval()
    assert small._dims.spike_method == _cabi.SPIKE_DIRECT


def test_panel_moments_are_a_partition_of_unity():
    """sum_i tau[r][n][i] = number of spikes of (r, n) (the Lagrange cardinal functions sum to one), and the first
    moment reproduces sum_s t_s: the static data the panel path is built on."""
    from svgpfa_b200.testing import model_from_case
    cfg = dict(synthetic.CONFIGS["config4"], R=5, N=40)
    case = synthetic.make_case(cfg, seed=3)
    model = model_from_case(case, spike_method="panel")
    model.eval()
    B = model._dims.pm_B
    NB = 16 * B
    tau = model._ws["pm_tau"].view(cfg["R"], cfg["N"], NB).cpu().numpy()
    counts = case["spike_counts"]
    assert np.abs(tau.sum(-1) - counts).max() <= 1e-10
    i = np.arange(NB)
    nodes = model._dims.pm_lo + model._dims.pm_w * (i // 16 + 0.5 + 0.5 * np.cos(np.pi * ((i % 16) + 0.5) / 16))
    seg_t = np.array([[t.sum() for t in trial] for trial in synthetic.nested_spikes(case)])
    assert np.abs(tau @ nodes - seg_t).max() <= 1e-9 * max(1.0, np.abs(seg_t).max())


@pytest.mark.parametrize("flags", [1, 2, 4, 8, 5])
def test_panel_path_gradient_subsets(flags):
    """svEM's steps ask for one parameter group at a time (E-step, embedding, kernels, inducing points): each subset
    through the panel path equals the direct kernel."""
    from svgpfa_b200.testing import grads_as_dict, model_from_case, set_requires_grad
    cfg = dict(synthetic.CONFIGS["config4"], R=4, N=60)
    case = synthetic.make_case(cfg, seed=4)
    outs = []
    for method in ("direct", "panel"):
        model = model_from_case(case, spike_method=method)
        set_requires_grad(model, posterior=bool(flags & 1), embedding=bool(flags & 2), kernels=bool(flags & 4),
                          indlocs=bool(flags & 8))
        for rep in range(2):                       # the second evaluation exercises the cached spike statistic
            for p in model._leaves():
                p.grad = None
            v = model.eval()
            (-v).backward()
        out = grads_as_dict(model)
        out["elbo"] = v.item()
        outs.append(out)
    d, p = outs
    assert abs(p["elbo"] - d["elbo"]) <= 1e-12 * abs(d["elbo"])
    for key in d:
        if key == "elbo":
            continue
        assert (d[key] is None) == (p[key] is None), key
        if d[key] is not None:
            assert rel_err(p[key], d[key]) <= 1e-9, key


def test_host_buffer_entry_rebuilds_panels():
    """svgpfa_elbo_grad_host with new spikes (copy_static) rebuilds the panel moments block by block."""
    from svgpfa_b200.testing import model_from_case
    cfg = dict(synthetic.CONFIGS["config5"], R=8, N=120)
    case = synthetic.make_case(cfg, seed=6)
    ref = model_from_case(case, spike_method="direct")
    want = ref.eval().item()
    model = model_from_case(case, spike_method="panel")
    io = model.makeHostIO(pin=True)
    model.eval()
    model._ws["pm_tau"].fill_(float("nan"))            # a missed rebuild cannot go unnoticed
    for nb in (1, 3):
        elbo, _, _ = model.evalAndGradHost(io, copy_static=True, n_blocks=nb)
        assert abs(elbo - want) <= 1e-11 * abs(want)


def test_cached_statistics_through_the_panel_moments():
    """Embedding M-step (svEM.py:225-232) with the panel path active: the per-neuron sums of the spike-time means come
    from the panel moments (no (S, K) array is built), the cached ELL and its (C, d) gradients equal the direct path,
    and the spike-time means are still available on demand."""
    from svgpfa_b200.testing import model_from_case, set_requires_grad
    cfg = dict(synthetic.CONFIGS["config4"], R=5, N=50)
    case = synthetic.make_case(cfg, seed=8)
    res = {}
    for method in ("direct", "panel"):
        model = model_from_case(case, spike_method=method)
        set_requires_grad(model, posterior=False, embedding=True, kernels=False, indlocs=False)
        stats = model.computeSVPosteriorOnLatentsStats()
        assert ("_b200_gsum" in stats) == (method == "panel")
        vals = []
        for rep in range(3):                        # later closures reuse the per-neuron sums
            for p in model.getSVEmbeddingParams():
                p.grad = None
            v = model.evalELLSumAcrossTrialsAndNeurons(svPosteriorOnLatentsStats=stats)
            (-v).backward()
            vals.append(v.item())
            with torch.no_grad():
                model.getSVEmbeddingParams()[0].mul_(1.01)       # what the optimiser does between closures
        C, d = model.getSVEmbeddingParams()
        res[method] = (vals, C.grad.cpu().numpy(), d.grad.cpu().numpy(),
                       torch.cat(list(stats["assocTimes"][0]), 0).cpu().numpy())
    for a, b in zip(res["direct"][0], res["panel"][0]):
        assert abs(a - b) <= 1e-12 * abs(a)
    assert rel_err(res["panel"][1], res["direct"][1]) <= 1e-11
    assert rel_err(res["panel"][2], res["direct"][2]) <= 1e-11
    assert rel_err(res["panel"][3], res["direct"][3]) <= 1e-13
    assert res["direct"][0][0] != res["direct"][0][1]
