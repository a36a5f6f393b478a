"""Stage-level checks of individual C-ABI entry points on the GPU against the numpy restatement
(oracle/analytic_np.py) and libdevice."""
import ctypes
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_err
from svgpfa_b200 import synthetic

pytestmark = pytest.mark.gpu


def test_exp_neg_accuracy():
    """The library's exp for non-positive arguments vs numpy/libdevice exp down to exp(-706.9); below that
    (true value < 1e-306) it returns a value in [0, 1e-306]."""
    from svgpfa_b200 import _cabi
    lib = _cabi.probes()
    dev = torch.device("cuda")
    gen = torch.Generator(device="cpu").manual_seed(0)
    x = torch.cat([-torch.rand(200000, generator=gen, dtype=torch.float64) * 40.0,
                   -torch.rand(100000, generator=gen, dtype=torch.float64) * 700.0,
                   -torch.rand(100000, generator=gen, dtype=torch.float64) * 1e-3,
                   torch.tensor([0.0, -1e-300, -708.0, -709.5, -745.0, -800.0, -1e6, -1e300],
                                dtype=torch.float64)]).to(dev)
    yf, yr = torch.empty_like(x), torch.empty_like(x)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _cabi.check_probe(lib.svgpfa_exp_neg_eval(x.data_ptr(), yf.data_ptr(), yr.data_ptr(), x.numel(), stream))
    torch.cuda.synchronize()
    x, yf, yr = x.cpu().numpy(), yf.cpu().numpy(), yr.cpu().numpy()
    truth = np.exp(x)
    normal = x > -706.9
    # single-constant range reduction: relative error grows like 8e-17 |x| (the rounding of the ARGUMENT
    # nh*q, which every implementation including the reference shares, is already 1.1e-16 |x|)
    err = np.abs(yf[normal] - truth[normal]) / truth[normal]
    bound = 4e-16 + 1.0e-16 * np.abs(x[normal])
    assert np.all(err <= bound), float((err / bound).max())
    assert np.abs(yf[normal] - truth[normal]).max() <= 2.3e-16          # absolute error: below half an ulp of 1
    assert np.all(yf[~normal] <= 1e-306) and np.all(yf[~normal] >= 0.0)
    assert yf[x == 0.0][0] == 1.0
    # libdevice itself, for scale
    err_ref = np.abs(yr[normal] - truth[normal]) / truth[normal]
    assert err_ref.max() <= 4e-16


@pytest.mark.parametrize("variant,bound", [(0, 4.5e-16), (3, 2.2e-14)])
def test_exp2m_accuracy(variant, bound):
    """The spike kernels' pre-scaled exponential 2^(-w2/256) vs numpy (long double) over its whole range;
    arguments above the limit are clamped (result ~6e-308 instead of 0 or a subnormal).  Variant 0 is what the kernels
    use; variant 3 (degree-3 economised polynomial, maximum relative error 1.8e-14, rounded integer returned through
    I2F) is the measured-but-not-adopted faster form (svgpfa_b200/csrc/spike.cu)."""
    from svgpfa_b200 import _cabi
    lib = _cabi.probes()
    dev = torch.device("cuda")
    gen = torch.Generator(device="cpu").manual_seed(1)
    w2 = torch.cat([torch.rand(200000, generator=gen, dtype=torch.float64) * 400.0,
                    torch.rand(200000, generator=gen, dtype=torch.float64) * 2.0e4,
                    torch.rand(100000, generator=gen, dtype=torch.float64) * 2.61e5,
                    torch.rand(50000, generator=gen, dtype=torch.float64),
                    torch.tensor([0.0, 0.5, 1.0, 127.5, 128.0, 255.999, 256.0, 2.609e5, 2.61e5, 3e6, 1e12, 1e300],
                                 dtype=torch.float64)]).to(dev)
    y = torch.empty_like(w2)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _cabi.check_probe(lib.svgpfa_exp2m_eval(w2.data_ptr(), y.data_ptr(), w2.numel(), variant, stream))
    torch.cuda.synchronize()
    w2, y = w2.cpu().numpy(), y.cpu().numpy()
    inside = w2 <= 2.609e5
    truth = np.exp2(-(w2[inside].astype(np.longdouble)) / np.longdouble(256)).astype(np.float64)
    err = np.abs(y[inside] - truth) / truth
    assert err.max() <= bound, float(err.max())
    assert abs(y[w2 == 0.0][0] - 1.0) <= bound
    assert np.all(y[~inside] > 0.0) and np.all(y[~inside] < 1e-306)


@pytest.mark.parametrize("name", ["tiny_mixed", "matlab_r5"])
def test_stage_buffers_match_numpy_restatement(name):
    """Li, X, c, alpha, KL_rk, mu/var at quadrature points, abar (spike) against oracle/analytic_np.py."""
    from oracle import analytic_np
    from svgpfa_b200.testing import model_from_case, set_requires_grad
    case, ref = synthetic.load_case(os.path.join(GOLDEN, name + ".npz"))
    out = analytic_np.elbo_and_grads(case)
    fw = out["stages"]["fw"]
    model = model_from_case(case)
    set_requires_grad(model)
    v = model.eval()
    torch.cuda.synchronize()
    ws = {k: t.cpu().numpy() for k, t in model._ws.items()}
    R, K = model._R, model._K
    for r in range(R):
        for k in range(K):
            M = model._M[k]
            mo = r * model._MM + model._mmoff[k]
            vo = r * model._KM + model._moff[k]
            f = fw[r, k]
            assert rel_err(ws["L"][mo:mo + M * M].reshape(M, M), f["L"]) <= 1e-10
            assert rel_err(ws["Li"][mo:mo + M * M].reshape(M, M), f["Li"]) <= 1e-9
            assert rel_err(ws["X"][mo:mo + M * M].reshape(M, M), f["X"]) <= 1e-9
            assert rel_err(ws["c"][vo:vo + M], f["c"]) <= 1e-9
            assert rel_err(ws["alpha"][vo:vo + M], f["alpha"]) <= 1e-8
            assert abs(ws["kl_rk"][r * K + k] - f["kl"]) <= 1e-9 * abs(f["kl"])
            abar_s = out["stages"]["bw"][r, k][1]
            assert rel_err(ws["abar_spk"][vo:vo + M], abar_s) <= 1e-10 or np.linalg.norm(abar_s) == 0
    Q = model._Q
    assert rel_err(ws["mu_q"].reshape(R, Q, K), out["quad_latent_mean"]) <= 1e-10
    assert rel_err(ws["var_q"].reshape(R, Q, K), out["quad_latent_var"]) <= 1e-9
    assert abs(v.item() - out["elbo"]) <= 1e-10 * abs(out["elbo"])


def test_predict_latents_matches_quadrature_stats():
    """predictLatents at the quadrature nodes reproduces the cached quadrature statistics
    (svPosteriorOnLatents.py:57-77 vs :79-86)."""
    from svgpfa_b200.testing import model_from_case
    case, ref = synthetic.load_case(os.path.join(GOLDEN, "tiny_mixed.npz"))
    model = model_from_case(case)
    mu, var = model.predictLatents(torch.as_tensor(case["leg_quad_points"]))
    assert rel_err(mu.cpu().numpy(), ref["quad_latent_mean"]) <= 1e-10
    assert rel_err(var.cpu().numpy(), ref["quad_latent_var"]) <= 1e-9
    e_mu, e_var = model.predictEmbedding(torch.as_tensor(case["leg_quad_points"]))
    assert rel_err(e_mu.cpu().numpy(), ref["quad_embedding_mean"]) <= 1e-10
    assert rel_err(e_var.cpu().numpy(), ref["quad_embedding_var"]) <= 1e-9


@pytest.mark.parametrize("name,n_blocks", [("config2_r8", 1), ("config2_r8", 3), ("config2_r8", 8), ("tiny_mixed", 2),
                                           ("matlab_r5", 5)])
def test_host_buffer_entry_matches_reference(name, n_blocks):
    """svgpfa_elbo_grad_host: everything in order on one stream (n_blocks = 1) and pipelined over blocks of trials
    (copy-in / kernels / copy-out on three streams) against the reference fixtures; the device buffers and the host
    outputs are poisoned first so that a missed copy cannot go unnoticed."""
    from svgpfa_b200.testing import model_from_case
    case, ref = synthetic.load_case(os.path.join(GOLDEN, name + ".npz"))
    model = model_from_case(case, nested=(name != "config2_r8"))
    io = model.makeHostIO(pin=True)
    for t in (model._Zbuf, model._mbuf, model._cvbuf, model._thbuf, model._C, model._d, model._tq, model._wq,
              model._spike_t):
        t.detach().fill_(float("nan"))
    model._seg_off.fill_(-1)
    for key in ("shared", "gZ", "gm", "gcholvec"):
        io[key].fill_(float("nan"))
    elbo, h2d, d2h = model.evalAndGradHost(io, copy_static=True, n_blocks=n_blocks)
    assert abs(elbo - float(ref["elbo"])) <= 1e-10 * abs(float(ref["elbo"]))
    R, K, N = model._R, model._K, model._N
    for k in range(K):
        M, P = model._M[k], model._P[k]
        gm = io["gm"][R * model._moff[k]:R * (model._moff[k] + M)].numpy().reshape(R, M, 1)
        assert rel_err(gm, ref[f"grad_m_{k}"]) <= 1e-8
        gz = io["gZ"][R * model._moff[k]:R * (model._moff[k] + M)].numpy().reshape(R, M, 1)
        assert rel_err(gz, ref[f"grad_Z_{k}"]) <= 1e-8
        gc = io["gcholvec"][R * model._poff[k]:R * (model._poff[k] + P)].numpy().reshape(R, P, 1)
        assert rel_err(gc, ref[f"grad_chol_vecs_{k}"]) <= 1e-8
    h = 8
    assert rel_err(io["shared"][h:h + N * K].numpy().reshape(N, K), ref["grad_C"]) <= 1e-8
    assert rel_err(io["shared"][h + N * K:h + N * K + N].numpy().reshape(-1), np.asarray(ref["grad_d"]).reshape(-1)) <= 1e-8
    th = io["shared"][h + N * K + N:].numpy()
    ref_th = np.concatenate([np.asarray(ref[f"grad_kernel_params_{k}"]).reshape(-1) for k in range(K)])
    assert rel_err(th, ref_th) <= 1e-8
    assert h2d > 0 and d2h > 0
    # a second call with parameters only (static inputs stay resident) gives the same bound
    elbo2, _, _ = model.evalAndGradHost(io, copy_static=False, n_blocks=n_blocks)
    assert abs(elbo2 - elbo) <= 1e-12 * abs(elbo)
