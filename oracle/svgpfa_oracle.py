"""CPU oracle for the svGPFA lower-bound hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

A float64 restatement, in plain PyTorch-CPU tensor ops, of the algorithm the reference
runs for ``SVLowerBound.eval()`` (= PointProcessELLExpLink - KLDivergence) and, through
``torch.autograd`` exactly as the reference does (it has no hand-written derivatives,
SURVEY.md §3.5), of its gradients.  Every function cites the reference lines it follows.
It keeps the reference's evaluation ORDER where that decides rounding (Cholesky per
trial, ``cholesky_solve`` for every Kzz solve, ``slogdet`` for both log-determinants,
variance at spike times computed and then dropped by the exp-link) so that it is also a
fair stand-in for the reference's CPU cost.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import this module.  The product package ``svgpfa_b200`` never
does; it fails loudly when its CUDA library is missing.

Parity pin: checked against fixtures produced by the unmodified reference
(``tests/golden/*.npz`` written by ``tests/golden/make_golden.py``; includes the MATLAB
golden problem of the reference's own unit tests) in ``tests/test_oracle_golden.py``.
"""
from __future__ import annotations

import math

import numpy as np
import torch

F64 = torch.float64


# --------------------------------------------------------------------------------------
# kernels  (reference: src/svGPFA/stats/kernels.py)
# --------------------------------------------------------------------------------------
def kernel_matrix(ktype: str, params: torch.Tensor, x1: torch.Tensor, x2: torch.Tensor,
                  scale: float = 1.0) -> torch.Tensor:
    """Covariance between two sets of times.

    ``x1``: (..., A, 1), ``x2``: (..., B, 1) batched, or 1-D/2-D unbatched.
    expquad : scale^2 exp(-0.5 (x-x')^2 / l^2)                 kernels.py:33-46
    periodic: scale^2 exp(-2 sin^2(pi (x-x') / p) / l^2)       kernels.py:73-85
    """
    if x1.ndim == 3:
        delta = x1 - x2.transpose(1, 2)
    else:
        delta = x1.reshape(-1, 1) - x2.reshape(1, -1)
    if ktype == "expquad":
        lengthscale = params[0]
        return scale ** 2 * torch.exp(-0.5 * delta ** 2 / lengthscale ** 2)
    if ktype == "periodic":
        lengthscale, period = params[0], params[1]
        rr = math.pi * delta / period
        return scale ** 2 * torch.exp(-2.0 * torch.sin(rr) ** 2 / lengthscale ** 2)
    raise ValueError(f"unknown kernel type {ktype!r}")


def kernel_diag(x: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
    """kappa(t, t) = scale^2 for both kernels, no regulariser (kernels.py:48-51, 87-90)."""
    return scale ** 2 * torch.ones(x.shape, dtype=x.dtype)


# --------------------------------------------------------------------------------------
# spike stacking  (reference: expectedLogLikelihood.py:157-173)
# --------------------------------------------------------------------------------------
def stack_spike_times(measurements):
    """Per trial: all spike times in neuron-major order (within-neuron order kept, NOT
    time sorted) and the neuron index of every stacked spike (int64)."""
    times, neuron_index = [], []
    for trial in measurements:
        t_list = [torch.as_tensor(np.asarray(s)).reshape(-1) for s in trial]
        n_list = [torch.full((len(s),), n, dtype=torch.int64) for n, s in enumerate(t_list)]
        times.append(torch.cat(t_list) if t_list else torch.zeros(0, dtype=F64))
        neuron_index.append(torch.cat(n_list) if n_list else torch.zeros(0, dtype=torch.int64))
    return times, neuron_index


# --------------------------------------------------------------------------------------
# variational covariance  (reference: utils/miscUtils.py:135-155)
# --------------------------------------------------------------------------------------
def chol_from_vec(vec: torch.Tensor, M: int) -> torch.Tensor:
    """Row-major lower-triangular scatter (torch.tril_indices order), miscUtils.py:135-139."""
    Ls = torch.zeros((M, M), dtype=F64)
    ti = torch.tril_indices(M, M)
    Ls[ti[0], ti[1]] = vec
    return Ls


def build_covs(chol_vecs):
    """S_kr = Ls Ls^T per latent and trial (python loop K x R), miscUtils.py:141-155."""
    covs = []
    for cv in chol_vecs:
        R, P = cv.shape[0], cv.shape[1]
        M = int((-1 + math.sqrt(1 + 8 * P)) / 2)
        rows = []
        for r in range(R):
            Ls = chol_from_vec(cv[r, :, 0], M)
            rows.append(Ls @ Ls.T)
        covs.append(torch.stack(rows))
    return covs


# --------------------------------------------------------------------------------------
# kernel-matrix stores  (reference: kernelsMatricesStore.py:107-138, 186-195, 208-221)
# --------------------------------------------------------------------------------------
def build_kzz(kernel_types, kernel_params, Z, reg):
    """Kzz_k = kappa_k(Z_k, Z_k) + reg I and its per-trial Cholesky factor
    (kernelsMatricesStore.py:107-117; utils/miscUtils.py:209-216)."""
    Kzz, Lzz = [], []
    for k, z in enumerate(Z):
        Kk = kernel_matrix(kernel_types[k], kernel_params[k], z, z) + reg * torch.eye(z.shape[1], dtype=F64)
        Lk = torch.stack([torch.linalg.cholesky(Kk[r]) for r in range(Kk.shape[0])])
        Kzz.append(Kk)
        Lzz.append(Lk)
    return Kzz, Lzz


def build_ktz_quad(kernel_types, kernel_params, Z, t_quad):
    """Ktz_k (R, Q, M_k) and KttDiag (R, Q, K) at quadrature times (kernelsMatricesStore.py:186-195)."""
    Ktz = [kernel_matrix(kernel_types[k], kernel_params[k], t_quad, Z[k]) for k in range(len(Z))]
    ktt = torch.stack([kernel_diag(t_quad).squeeze(-1) for _ in Z], dim=2)
    return Ktz, ktt


def build_ktz_spikes(kernel_types, kernel_params, Z, spike_times):
    """Ktz[k][r] (S_r, M_k), KttDiag[k][r] (S_r,) at spike times, python loop K x R
    (kernelsMatricesStore.py:208-221)."""
    K, R = len(Z), len(spike_times)
    Ktz = [[kernel_matrix(kernel_types[k], kernel_params[k], spike_times[r], Z[k][r]) for r in range(R)]
           for k in range(K)]
    ktt = [[kernel_diag(spike_times[r]) for r in range(R)] for k in range(K)]
    return Ktz, ktt


# --------------------------------------------------------------------------------------
# posterior on latents  (reference: svPosteriorOnLatents.py:185-216, 265-300)
# --------------------------------------------------------------------------------------
def latents_at_quad(Kzz, Lzz, Ktz, ktt, m, covs):
    """mu (R,Q,K), var (R,Q,K): A = Kzz^-1 m, mu = Ktz A, B = Kzz^-1 Kzt,
    var = ktt + sum_m B o ((S - Kzz) B)   (svPosteriorOnLatents.py:185-216)."""
    mu, var = [], []
    for k in range(len(Kzz)):
        A = torch.cholesky_solve(m[k], Lzz[k])
        mu.append((Ktz[k] @ A).squeeze(-1))
        B = torch.cholesky_solve(Ktz[k].transpose(1, 2), Lzz[k])
        var.append(ktt[:, :, k] + (B * ((covs[k] - Kzz[k]) @ B)).sum(dim=1))
    return torch.stack(mu, dim=2), torch.stack(var, dim=2)


def latents_at_spikes(Kzz, Lzz, Ktz, ktt, m, covs, with_var=True):
    """Per trial (S_r, K) means and variances, python loop R x K
    (svPosteriorOnLatents.py:265-300)."""
    K, R = len(Kzz), len(Ktz[0])
    A = [torch.cholesky_solve(m[k], Lzz[k]) for k in range(K)]
    mu, var = [], []
    for r in range(R):
        mu_r, var_r = [], []
        for k in range(K):
            mu_r.append((Ktz[k][r] @ A[k][r]).reshape(-1))
            if with_var:
                B = torch.cholesky_solve(Ktz[k][r].T, Lzz[k][r])
                var_r.append(ktt[k][r].reshape(-1) + (B * ((covs[k][r] - Kzz[k][r]) @ B)).sum(dim=0))
        mu.append(torch.stack(mu_r, dim=1) if mu_r else torch.zeros(0, K, dtype=F64))
        var.append(torch.stack(var_r, dim=1) if with_var else None)
    return mu, var


# --------------------------------------------------------------------------------------
# embedding  (reference: svEmbedding.py:80-84, 137-144)
# --------------------------------------------------------------------------------------
def embed_quad(mu, var, C, d):
    """(R,Q,N) mean = mu C^T + d, var = var (C^T)^2  (svEmbedding.py:80-84)."""
    return mu @ C.T + d.reshape(1, 1, -1), var @ (C.T ** 2)


def embed_spikes(mu, var, C, d, neuron_index):
    """Per trial (S_r,) mean = sum_k mu[s,k] C[n_s,k] + d[n_s] (and the C^2 analogue)
    (svEmbedding.py:137-144)."""
    dd = d.reshape(-1)
    e_mu, e_var = [], []
    for r in range(len(mu)):
        idx = neuron_index[r]
        e_mu.append((mu[r] * C[idx, :]).sum(dim=1) + dd[idx])
        e_var.append((var[r] * C[idx, :] ** 2).sum(dim=1) if var[r] is not None else None)
    return e_mu, e_var


# --------------------------------------------------------------------------------------
# expected log-likelihood and KL  (reference: expectedLogLikelihood.py:107-135,199-213;
#                                  klDivergence.py:18-44)
# --------------------------------------------------------------------------------------
def ell_exp_link(eq_mu, eq_var, es_mu, w):
    """-sum_r w_r^T exp(mean + var/2) 1 + sum_spikes mean  (the spike-time variance is
    ignored by the exponential link, expectedLogLikelihood.py:205-213)."""
    e_link = torch.exp(eq_mu + 0.5 * eq_var)
    term1 = (w.transpose(1, 2) @ e_link).sum()
    term2 = torch.cat(es_mu).sum() if len(es_mu) else torch.zeros((), dtype=F64)
    return -term1 + term2


def kl_divergence(Kzz, Lzz, m, covs):
    """sum_{k,r} 0.5 [tr(Kzz^-1 (S + m m^T)) + log|Kzz| - log|S| - M], slogdet for both
    log-determinants, python loop K x R  (klDivergence.py:18-44)."""
    total = torch.zeros((), dtype=F64)
    for k in range(len(Kzz)):
        ess = covs[k] + m[k] @ m[k].transpose(1, 2)
        for r in range(m[k].shape[0]):
            _, ld_k = Kzz[k][r].slogdet()
            _, ld_s = covs[k][r].slogdet()
            tr = torch.trace(torch.cholesky_solve(ess[r], Lzz[k][r]))
            total = total + 0.5 * (tr + ld_k - ld_s - ess.shape[1])
    return total


# --------------------------------------------------------------------------------------
# the whole path
# --------------------------------------------------------------------------------------
def build_covs_rank1(q_svec, q_sdiag):
    """S_kr = q q^T + diag(d^2): the rank-1-plus-diagonal parameterisation
    (svPosteriorOnIndPoints.py:86-119, SVPosteriorOnIndPointsRank1PlusDiag.buildCov)."""
    covs = []
    for q, d in zip(q_svec, q_sdiag):
        covs.append(q @ q.transpose(1, 2) + torch.diag_embed(d[:, :, 0] ** 2))
    return covs


def to_tensors(case, requires_grad=False):
    t = lambda a: torch.tensor(np.asarray(a), dtype=F64)
    p = dict(m=[t(a) for a in case["m"]], chol_vecs=[t(a) for a in case["chol_vecs"]],
             C=t(case["C"]), d=t(case["d"]),
             kernel_params=[t(a) for a in case["kernel_params"]], Z=[t(a) for a in case["Z"]])
    if requires_grad:
        for group in ("m", "chol_vecs", "kernel_params", "Z"):
            for a in p[group]:
                a.requires_grad_(True)
        p["C"].requires_grad_(True)
        p["d"].requires_grad_(True)
    return p


def case_spikes(case):
    """Stacked spike times (dtype preserved: float32 inputs are promoted to float64 only
    inside the kernel difference, kernels.py:42-44) and neuron indices per trial."""
    counts = np.asarray(case["spike_counts"])
    R, N = counts.shape
    per_trial = counts.sum(axis=1)
    off = np.concatenate([[0], np.cumsum(per_trial)])
    st = np.asarray(case["spike_times"])
    times = [torch.from_numpy(np.ascontiguousarray(st[off[r]:off[r + 1]])) for r in range(R)]
    idx = [torch.from_numpy(np.repeat(np.arange(N, dtype=np.int64), counts[r])) for r in range(R)]
    return times, idx


def elbo_terms(case, p, spike_var=True):
    """(ELL, KL) as 0-dim tensors attached to the autograd graph of ``p``."""
    kt, reg = case["kernel_types"], case["reg"]
    tq = torch.tensor(np.asarray(case["leg_quad_points"]), dtype=F64)
    w = torch.tensor(np.asarray(case["leg_quad_weights"]), dtype=F64)
    times, idx = case_spikes(case)
    Kzz, Lzz = build_kzz(kt, p["kernel_params"], p["Z"], reg)
    Ktz_q, ktt_q = build_ktz_quad(kt, p["kernel_params"], p["Z"], tq)
    Ktz_s, ktt_s = build_ktz_spikes(kt, p["kernel_params"], p["Z"], times)
    covs = build_covs_rank1(p["q_svec"], p["q_sdiag"]) if "q_svec" in p else build_covs(p["chol_vecs"])
    mu_q, var_q = latents_at_quad(Kzz, Lzz, Ktz_q, ktt_q, p["m"], covs)
    mu_s, var_s = latents_at_spikes(Kzz, Lzz, Ktz_s, ktt_s, p["m"], covs, with_var=spike_var)
    eq_mu, eq_var = embed_quad(mu_q, var_q, p["C"], p["d"])
    es_mu, _ = embed_spikes(mu_s, var_s, p["C"], p["d"], idx)
    ell = ell_exp_link(eq_mu, eq_var, es_mu, w)
    kl = kl_divergence(Kzz, Lzz, p["m"], covs)
    return ell, kl, dict(mu_q=mu_q, var_q=var_q, mu_s=mu_s, var_s=var_s, eq_mu=eq_mu,
                         eq_var=eq_var, es_mu=es_mu)


def elbo_and_grads(case, spike_var=True, with_stats=False):
    """One unit of work of the benchmark: build matrices, evaluate the
    lower bound, back-propagate to every parameter group.  Returns a dict shaped like the
    golden fixtures' ``out_*`` entries."""
    p = to_tensors(case, requires_grad=True)
    ell, kl, stats = elbo_terms(case, p, spike_var=spike_var)
    elbo = ell - kl
    elbo.backward()
    out = {"elbo": elbo.item(), "ell": ell.item(), "kl": kl.item(),
           "grad_C": p["C"].grad.numpy(), "grad_d": p["d"].grad.numpy()}
    for k in range(len(case["kernel_types"])):
        out[f"grad_m_{k}"] = p["m"][k].grad.numpy()
        out[f"grad_chol_vecs_{k}"] = p["chol_vecs"][k].grad.numpy()
        out[f"grad_kernel_params_{k}"] = p["kernel_params"][k].grad.numpy()
        out[f"grad_Z_{k}"] = p["Z"][k].grad.numpy()
    if with_stats:
        out["quad_latent_mean"] = stats["mu_q"].detach().numpy()
        out["quad_latent_var"] = stats["var_q"].detach().numpy()
        out["spike_latent_mean"] = torch.cat(stats["mu_s"]).detach().numpy()
        if spike_var:
            out["spike_latent_var"] = torch.cat(stats["var_s"]).detach().numpy()
        out["quad_embedding_mean"] = stats["eq_mu"].detach().numpy()
        out["quad_embedding_var"] = stats["eq_var"].detach().numpy()
        out["spike_embedding_mean"] = torch.cat(stats["es_mu"]).detach().numpy()
    return out


def ell_from_cached_stats(case, mu_q, var_q, mu_s, C, d):
    """Embedding M-step objective: ELL from cached latent statistics
    (svLowerBound.py:72-75 -> expectedLogLikelihood.py:107-135 with svPosteriorOnLatentsStats)."""
    w = torch.tensor(np.asarray(case["leg_quad_weights"]), dtype=F64)
    _, idx = case_spikes(case)
    eq_mu, eq_var = embed_quad(mu_q, var_q, C, d)
    es_mu, _ = embed_spikes(mu_s, [None] * len(mu_s), C, d, idx)
    return ell_exp_link(eq_mu, eq_var, es_mu, w)


def elbo_and_grads_rank1(case, q_svec, q_sdiag):
    """The same unit of work with the variational covariance given as (q, d) instead of Cholesky vectors
    (stats/svGPFAModelFactory.py: indPointsCovRep = indPointsCovRank1PlusDiag)."""
    p = to_tensors(case, requires_grad=True)
    p.pop("chol_vecs")
    p["q_svec"] = [torch.tensor(np.asarray(a), dtype=F64, requires_grad=True) for a in q_svec]
    p["q_sdiag"] = [torch.tensor(np.asarray(a), dtype=F64, requires_grad=True) for a in q_sdiag]
    ell, kl, _ = elbo_terms(case, p, spike_var=False)
    elbo = ell - kl
    elbo.backward()
    out = {"elbo": elbo.item(), "ell": ell.item(), "kl": kl.item(),
           "grad_C": p["C"].grad.numpy(), "grad_d": p["d"].grad.numpy()}
    for k in range(len(case["kernel_types"])):
        out[f"grad_m_{k}"] = p["m"][k].grad.numpy()
        out[f"grad_q_svec_{k}"] = p["q_svec"][k].grad.numpy()
        out[f"grad_q_sdiag_{k}"] = p["q_sdiag"][k].grad.numpy()
        out[f"grad_kernel_params_{k}"] = p["kernel_params"][k].grad.numpy()
        out[f"grad_Z_{k}"] = p["Z"][k].grad.numpy()
    return out
