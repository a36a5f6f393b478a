"""A minimal ECM driver with the call sequence of the reference's ``SVEM_PyTorch`` (TEST INFRASTRUCTURE).

/root/reference is not available on the GPU box, so the parity tests cannot import ``svGPFA.stats.svEM``
there.  This file restates only the *protocol usage* of ``stats/svEM.py`` -- which model methods are called, in
which order, with which ``requires_grad`` toggling and LBFGS closure -- so that the same driver can be run on
the CUDA model and on an oracle-backed CPU model:
    maximize            svEM.py:76-216   (ECM: estep, mstep_embedding, mstep_kernels, mstep_indpointslocs)
    _eStep              svEM.py:218-223
    _mStepEmbedding     svEM.py:225-232
    _mStepKernels       svEM.py:234-254  (parameters restored on exception)
    _mStepIndPointsLocs svEM.py:256-264
    _setupAndMaximizeStep / _maximizeStep  svEM.py:266-294
"""
import copy

import numpy as np
import torch

from oracle import svgpfa_oracle as orc


def _maximize_step(x, eval_func, lbfgs_kwargs):
    optimizer = torch.optim.LBFGS(x, **lbfgs_kwargs)
    for p in x:
        p.requires_grad = True

    def closure():
        optimizer.zero_grad()
        cur = -eval_func()
        cur.backward(retain_graph=True)
        return cur
    optimizer.step(closure)
    lower_bound = eval_func()
    state = optimizer.state[optimizer._params[0]]
    for p in x:
        p.requires_grad = False
    return {"lowerBound": lower_bound, "nfeval": state["func_evals"], "niter": state["n_iter"]}


def e_step(model, kw):
    return _maximize_step(model.getSVPosteriorOnIndPointsParams(), model.eval, kw)


def m_step_embedding(model, kw):
    x = model.getSVEmbeddingParams()
    stats = model.computeSVPosteriorOnLatentsStats()
    return _maximize_step(x, lambda: model.evalELLSumAcrossTrialsAndNeurons(svPosteriorOnLatentsStats=stats), kw)


def _build_and_eval(model):
    model.buildKernelsMatrices()
    return model.eval()


def m_step_kernels(model, kw):
    x = model.getKernelsParams()
    prev = [copy.deepcopy(p) for p in x]
    try:
        return _maximize_step(x, lambda: _build_and_eval(model), kw)
    except Exception:
        for p, q in zip(x, prev):
            p.detach()[:] = q[:]
        raise


def m_step_indpointslocs(model, kw):
    return _maximize_step(model.getIndPointsLocs(), lambda: _build_and_eval(model), kw)


STEPS = (("estep", e_step), ("mstep_embedding", m_step_embedding), ("mstep_kernels", m_step_kernels),
         ("mstep_indpointslocs", m_step_indpointslocs))


def maximize(model, em_max_iter, lbfgs_kwargs):
    hist = [model.eval().item()]
    log = []
    for it in range(1, em_max_iter + 1):
        for name, fn in STEPS:
            res = fn(model, lbfgs_kwargs)
            log.append((it, name, res["lowerBound"].item(), res["niter"], res["nfeval"]))
        hist.append(log[-1][2])
    return hist, log


class OracleModel:
    """The model protocol of SURVEY.md §8b on top of the CPU oracle (autograd), for driver-level parity tests."""

    def __init__(self, case):
        self.case = case
        self.p = orc.to_tensors(case, requires_grad=False)

    def getSVPosteriorOnIndPointsParams(self):
        return list(self.p["m"]) + list(self.p["chol_vecs"])

    def getSVEmbeddingParams(self):
        return [self.p["C"], self.p["d"]]

    def getKernelsParams(self):
        return list(self.p["kernel_params"])

    def getIndPointsLocs(self):
        return list(self.p["Z"])

    def buildKernelsMatrices(self):
        pass                                  # the oracle rebuilds everything on every evaluation

    def eval(self):
        ell, kl, _ = orc.elbo_terms(self.case, self.p, spike_var=False)
        return ell - kl

    def computeSVPosteriorOnLatentsStats(self):
        with torch.no_grad():
            _, _, st = orc.elbo_terms(self.case, self.p, spike_var=False)
        return {"allTimes": (st["mu_q"], st["var_q"]), "assocTimes": (st["mu_s"], [None] * len(st["mu_s"]))}

    def evalELLSumAcrossTrialsAndNeurons(self, svPosteriorOnLatentsStats):
        mu_q, var_q = svPosteriorOnLatentsStats["allTimes"]
        return orc.ell_from_cached_stats(self.case, mu_q, var_q, svPosteriorOnLatentsStats["assocTimes"][0],
                                         self.p["C"], self.p["d"])
