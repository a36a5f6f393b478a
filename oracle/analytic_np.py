"""Stage-by-stage numpy restatement of the lower bound WITH hand-derived adjoints --
TEST INFRASTRUCTURE, NOT PRODUCT.

The reference has no hand-written derivatives (autograd only); the CUDA
library has nothing else.  This module states, in slow per-(trial, latent) numpy loops,
exactly the decomposition the CUDA kernels use (triangular
formulation with an explicit L^-1), so that every intermediate buffer of the C-ABI
(Li, X, c, alpha, mu/var at quadrature points, mubar/varbar, Xbar, Lbar, alphabar, ...)
can be compared stage by stage.  It is itself pinned against the autograd oracle
(``svgpfa_oracle.py``) and the reference's golden fixtures in
``tests/test_analytic_vs_oracle.py``.

Stages (one CUDA kernel each, names as in include/svgpfa_b200.h):
  kzz_chol          Kzz = kappa(Z,Z)+reg I, L = chol(Kzz), Li = L^-1      kernelsMatricesStore.py:107-138
  indpoints_fwd     Ls, X = Li Ls, c = Li m, alpha = Li^T c, KL_rk         svPosteriorOnIndPoints.py:47-49, klDivergence.py:31-44
  quad_latent_fwd   mu, var at quadrature points                          svPosteriorOnLatents.py:185-216
  quad_embed        exp-link integral + embedding adjoints                svEmbedding.py:80-84, expectedLogLikelihood.py:107-135
  quad_latent_bwd   adjoints of the quadrature posterior                  (autograd in the reference)
  spike_fwd_bwd     spike-time log-intensity term and all its adjoints    svPosteriorOnLatents.py:265-300, svEmbedding.py:137-144
  indpoints_bwd     adjoints through the solves and the Cholesky          (autograd in the reference)
"""
from __future__ import annotations

import numpy as np


# ---- kernel function and its partial derivatives (kernels.py:33-46, 73-85) -----------------
def kappa(ktype, theta, delta):
    """value, d/d(delta), d/d(theta_i) of kappa at delta = x - z (scale = 1)."""
    if ktype == "expquad":
        ell = theta[0]
        k = np.exp(-0.5 * delta ** 2 / ell ** 2)
        dk_ddelta = -k * delta / ell ** 2
        dk_dtheta = [k * delta ** 2 / ell ** 3]
    else:
        ell, p = theta[0], theta[1]
        s = np.sin(np.pi * delta / p)
        s2 = np.sin(2.0 * np.pi * delta / p)
        k = np.exp(-2.0 * s ** 2 / ell ** 2)
        dk_ddelta = -k * (2.0 * np.pi / (p * ell ** 2)) * s2
        dk_dtheta = [k * 4.0 * s ** 2 / ell ** 3,
                     k * (2.0 / ell ** 2) * s2 * (np.pi * delta / p ** 2)]
    return k, dk_ddelta, dk_dtheta


def unpack_chol(vec, M):
    Ls = np.zeros((M, M))
    Ls[np.tril_indices(M)] = vec
    return Ls


def kzz_chol(ktype, theta, z, reg):
    delta = z[:, None] - z[None, :]
    Kzz = kappa(ktype, theta, delta)[0] + reg * np.eye(len(z))
    L = np.linalg.cholesky(Kzz)
    Li = np.linalg.solve(L, np.eye(len(z)))
    return L, np.tril(Li)


def indpoints_fwd(L, Li, m, cholvec):
    M = len(m)
    Ls = unpack_chol(cholvec, M)
    X = Li @ Ls
    c = Li @ m
    alpha = Li.T @ c
    kl = 0.5 * ((X ** 2).sum() + (c ** 2).sum() + 2.0 * np.log(np.diag(L)).sum()
                - 2.0 * np.log(np.abs(np.diag(Ls))).sum() - M)
    return Ls, X, c, alpha, kl


def quad_latent_fwd(ktype, theta, z, tq, Li, X, alpha):
    Kq = kappa(ktype, theta, tq[:, None] - z[None, :])[0]     # (Q, M)
    V = Kq @ Li.T                                             # v_q = Li k_q
    U = V @ X                                                 # u_q = X^T v_q
    mu = Kq @ alpha
    var = 1.0 - (V ** 2).sum(1) + (U ** 2).sum(1)
    return mu, var


def quad_embed(mu, var, w, C, d):
    """One trial: mu, var (Q, K); w (Q,).  Returns term1, dC, dd, mubar, varbar."""
    H = mu @ C.T + d[None, :]
    Sg = var @ (C.T ** 2)
    E = np.exp(H + 0.5 * Sg)
    term1 = float((w[:, None] * E).sum())
    G = -w[:, None] * E                                        # (Q, N)
    dd = G.sum(0)
    dC = G.T @ mu + C * (G.T @ var)
    mubar = G @ C
    varbar = 0.5 * G @ (C ** 2)
    return term1, dC, dd, mubar, varbar


def quad_latent_bwd(ktype, theta, z, tq, Li, X, alpha, mubar, varbar):
    delta = tq[:, None] - z[None, :]
    Kq, dK_dd, dK_dth = kappa(ktype, theta, delta)
    V = Kq @ Li.T
    U = V @ X
    Ubar = 2.0 * varbar[:, None] * U
    Vbar = -2.0 * varbar[:, None] * V + Ubar @ X.T
    Xbar = np.tril(V.T @ Ubar)
    Kbar_v = Vbar @ Li                                         # Li^T vbar_q, as rows
    Lbar = -np.tril(Kbar_v.T @ V)
    alphabar = Kq.T @ mubar
    Kbar = Kbar_v + mubar[:, None] * alpha[None, :]
    dz = -(Kbar * dK_dd).sum(0)                                # d delta / d z = -1
    dtheta = np.array([(Kbar * g).sum() for g in dK_dth])
    return Xbar, Lbar, alphabar, dz, dtheta


def spike_fwd_bwd(ktype, theta, z, alpha, times, neuron_index, Ck, N):
    """One (trial, latent): Ck = C[:, k].  Returns alphabar (M), dCk (N), dz, dtheta.
    The value of the term is alpha . alphabar (+ the d part, added by the caller)."""
    delta = times.astype(np.float64)[:, None] - z[None, :]
    Ks, dK_dd, dK_dth = kappa(ktype, theta, delta)
    cs = Ck[neuron_index]
    alphabar = (cs[:, None] * Ks).sum(0)
    mu_s = Ks @ alpha
    dCk = np.bincount(neuron_index, weights=mu_s, minlength=N)
    Kbar = cs[:, None] * alpha[None, :]
    dz = -(Kbar * dK_dd).sum(0)
    dtheta = np.array([(Kbar * g).sum() for g in dK_dth])
    return alphabar, dCk, dz, dtheta, mu_s


def indpoints_bwd(ktype, theta, z, L, Li, Ls, X, c, alpha, alphabar, Xbar, Lbar,
                  need_kernel_grads=True):
    """Adjoints through alpha = Li^T c, c = Li m, X = Li Ls, the KL term and L = chol(Kzz).
    Returns mbar, cholvec-bar, dz, dtheta (ELBO gradients)."""
    M = len(z)
    Lbar = Lbar.copy()
    # alpha = L^-T c
    y = Li @ alphabar
    cbar = y - c                                               # KL: -c
    Lbar -= np.outer(alpha, y)
    # KL pieces
    Xb = Xbar - X
    Lbar[np.diag_indices(M)] -= 1.0 / np.diag(L)
    # X = L^-1 Ls
    T = Li.T @ np.tril(Xb)
    Lsbar = np.tril(T)
    Lsbar[np.diag_indices(M)] += 1.0 / np.diag(Ls)
    Lbar -= T @ X.T
    # c = L^-1 m
    mbar = Li.T @ cbar
    Lbar -= np.outer(mbar, c)
    cvbar = Lsbar[np.tril_indices(M)]
    if not need_kernel_grads:
        return mbar, cvbar, np.zeros(M), np.zeros(len(theta))
    # L = chol(Kzz)
    Lbar = np.tril(Lbar)
    P = np.tril(L.T @ Lbar)
    P[np.diag_indices(M)] *= 0.5
    Kbar = 0.5 * Li.T @ (P + P.T) @ Li
    delta = z[:, None] - z[None, :]
    _, dK_dd, dK_dth = kappa(ktype, theta, delta)
    # Kzz_ij depends on z_i (delta_ij = z_i - z_j, +) and z_j (-)
    Gm = Kbar * dK_dd
    dz = Gm.sum(1) - Gm.sum(0)
    dtheta = np.array([(Kbar * g).sum() for g in dK_dth])
    return mbar, cvbar, dz, dtheta


def elbo_and_grads(case):
    """Whole path, returning the same dict layout as the golden fixtures, plus the
    per-stage intermediates under ``stages``."""
    kt = case["kernel_types"]
    K = len(kt)
    counts = np.asarray(case["spike_counts"])
    R, N = counts.shape
    C = np.asarray(case["C"], dtype=np.float64)
    d = np.asarray(case["d"], dtype=np.float64).reshape(-1)
    tq = np.asarray(case["leg_quad_points"])[:, :, 0]
    wq = np.asarray(case["leg_quad_weights"])[:, :, 0]
    Q = tq.shape[1]
    per_trial = counts.sum(1)
    off = np.concatenate([[0], np.cumsum(per_trial)])
    st = np.asarray(case["spike_times"])
    theta = [np.asarray(t, dtype=np.float64) for t in case["kernel_params"]]

    fw = {}
    mu = np.zeros((R, Q, K))
    var = np.zeros((R, Q, K))
    kl = 0.0
    for r in range(R):
        for k in range(K):
            z = case["Z"][k][r, :, 0]
            L, Li = kzz_chol(kt[k], theta[k], z, case["reg"])
            Ls, X, c, alpha, kl_rk = indpoints_fwd(L, Li, case["m"][k][r, :, 0], case["chol_vecs"][k][r, :, 0])
            kl += kl_rk
            mu[r, :, k], var[r, :, k] = quad_latent_fwd(kt[k], theta[k], z, tq[r], Li, X, alpha)
            fw[r, k] = dict(z=z, L=L, Li=Li, Ls=Ls, X=X, c=c, alpha=alpha, kl=kl_rk)
    term1 = 0.0
    dC = np.zeros_like(C)
    dd = np.zeros(N)
    mubar = np.zeros((R, Q, K))
    varbar = np.zeros((R, Q, K))
    for r in range(R):
        t1, dC_r, dd_r, mubar[r], varbar[r] = quad_embed(mu[r], var[r], wq[r], C, d)
        term1 += t1
        dC += dC_r
        dd += dd_r
    term2 = 0.0
    out = {}
    g_m = [np.zeros_like(np.asarray(a, dtype=np.float64)) for a in case["m"]]
    g_cv = [np.zeros_like(np.asarray(a, dtype=np.float64)) for a in case["chol_vecs"]]
    g_Z = [np.zeros_like(np.asarray(a, dtype=np.float64)) for a in case["Z"]]
    g_th = [np.zeros_like(t) for t in theta]
    mu_s = np.zeros((int(off[-1]), K))
    alphabar_all = {}
    for r in range(R):
        times = st[off[r]:off[r + 1]]
        nidx = np.repeat(np.arange(N), counts[r])
        term2 += d[nidx].sum()
        dd += np.bincount(nidx, minlength=N)
        for k in range(K):
            f = fw[r, k]
            Xbar, Lbar, abar_q, dz_q, dth_q = quad_latent_bwd(
                kt[k], theta[k], f["z"], tq[r], f["Li"], f["X"], f["alpha"], mubar[r, :, k], varbar[r, :, k])
            abar_s, dCk, dz_s, dth_s, mu_s[off[r]:off[r + 1], k] = spike_fwd_bwd(
                kt[k], theta[k], f["z"], f["alpha"], times, nidx, C[:, k], N)
            term2 += f["alpha"] @ abar_s
            dC[:, k] += dCk
            mbar, cvbar, dz_c, dth_c = indpoints_bwd(
                kt[k], theta[k], f["z"], f["L"], f["Li"], f["Ls"], f["X"], f["c"], f["alpha"],
                abar_q + abar_s, Xbar, Lbar)
            g_m[k][r, :, 0] = mbar
            g_cv[k][r, :, 0] = cvbar
            g_Z[k][r, :, 0] = dz_q + dz_s + dz_c
            g_th[k] += dth_q + dth_s + dth_c
            alphabar_all[r, k] = (abar_q, abar_s, Xbar, Lbar)
    ell = -term1 + term2
    out.update(elbo=ell - kl, ell=ell, kl=kl, grad_C=dC, grad_d=dd.reshape(np.asarray(case["d"]).shape))
    for k in range(K):
        out[f"grad_m_{k}"] = g_m[k]
        out[f"grad_chol_vecs_{k}"] = g_cv[k]
        out[f"grad_Z_{k}"] = g_Z[k]
        out[f"grad_kernel_params_{k}"] = g_th[k]
    out["quad_latent_mean"] = mu
    out["quad_latent_var"] = var
    out["spike_latent_mean"] = mu_s
    out["stages"] = dict(fw=fw, mubar=mubar, varbar=varbar, bw=alphabar_all, term1=term1, term2=term2)
    return out
