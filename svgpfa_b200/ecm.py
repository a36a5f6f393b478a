"""Shard-aware ECM driver for trial-sharded models (SURVEY.md §8e option (i), §8f-3).

On ONE GPU the reference's own ``svGPFA.stats.svEM.SVEM_PyTorch`` drives ``B200SVLowerBound`` unchanged; this
module is not needed there.  The reference has no multi-process mode at all, so with one rank per GPU (torchrun,
trials sharded, ``process_group`` set on the model) something has to decide where the ranks meet.  This driver
keeps the reference's step order, optimiser (``torch.optim.LBFGS``, one fresh instance per step), closure and
``requires_grad`` toggling (stats/svEM.py:112-114, 218-294) and adds exactly that:

  * steps over SHARED parameters (``mstep_embedding``: C, d; ``mstep_kernels``: theta) run in lock-step: every
    closure evaluation all-reduces the packed ``[elbo | dC | dd | dtheta]`` buffer inside the model, all ranks
    receive bit-identical sums and take identical optimiser decisions -- the trajectory of the 1-GPU run;
  * steps over PER-TRIAL parameters (``estep``: m, cholVecs; ``mstep_indpointslocs``: Z) exploit that the bound is
    a sum of independent per-trial terms given the shared parameters: every rank maximises the bound of its own
    trials with NO collective inside the closure (ranks may take different numbers of closure calls), and the step
    ends with ONE all-reduce of ``[final local bound, failure flag]``.  Same objective as the single-process joint
    optimisation, different (block-wise) L-BFGS trajectory -- trajectory parity with the reference is claimed for
    1-GPU runs and for the shared-parameter steps only;
  * a failure on one rank (non-positive-definite Kzz while its Z move) is agreed on at that all-reduce, so all
    ranks leave the step together instead of one of them abandoning the others in a collective.

Two further choices:

  * ``optimizer="b200"`` replaces ``torch.optim.LBFGS`` by ``svgpfa_b200.lbfgs.LBFGS`` (same decisions, state and
    vector work on the device, coefficient-space two-loop recursion);
  * ``sharded_steps="joint"`` (needs that optimiser) runs the PER-TRIAL steps as ONE optimisation over the
    concatenation of all ranks' vectors: every reduction the optimiser decides from is global (one small all-reduce per
    iteration and per closure call), every closure evaluation is all-reduced by the model, so all ranks stay in
    lock-step and follow the single-process trajectory -- SURVEY.md §8e option (ii) without gathering gradients.

``optim_params`` is the reference's hierarchical dictionary (``utils/initUtils.py:13-70``):
``em_max_iter``, ``{step}_estimate``, ``{step}_optim_params`` (keyword arguments of ``torch.optim.LBFGS``).
Works on any object with the model protocol of SURVEY.md §8b; ``process_group=None`` makes it a plain
single-process ECM loop.
"""
from __future__ import annotations

import sys
import time

import torch

STEP_ORDER = {
    "ecm": ("estep", "mstep_embedding", "mstep_kernels", "mstep_indpointslocs"),
    # McLachlan & Krishnan, ch. 5: an E-step before every conditional M-step (stats/svEM.py:118-121)
    "mecm": ("estep", "mstep_embedding", "estep", "mstep_kernels", "estep", "mstep_indpointslocs"),
}
SHARDED_STEPS = ("estep", "mstep_indpointslocs")          # their parameters live on the rank that owns the trials


class StepFailed(RuntimeError):
    """Raised on EVERY rank when a per-trial step failed on at least one of them."""


def _parameters_and_objective(model, step):
    if step == "estep":
        return model.getSVPosteriorOnIndPointsParams(), model.eval
    if step == "mstep_embedding":
        stats = model.computeSVPosteriorOnLatentsStats()          # cached once per step (svEM.py:227)
        return (model.getSVEmbeddingParams(),
                lambda: model.evalELLSumAcrossTrialsAndNeurons(svPosteriorOnLatentsStats=stats))

    def rebuild_and_eval():
        model.buildKernelsMatrices()
        return model.eval()
    if step == "mstep_kernels":
        return model.getKernelsParams(), rebuild_and_eval
    if step == "mstep_indpointslocs":
        return model.getIndPointsLocs(), rebuild_and_eval
    raise ValueError(f"unknown step {step!r}")


def _optimizer_factory(optimizer):
    if callable(optimizer):
        return optimizer
    if optimizer == "torch":
        return torch.optim.LBFGS
    if optimizer == "b200":
        from .lbfgs import LBFGS
        return LBFGS
    raise ValueError(f"optimizer must be 'torch', 'b200' or a factory, not {optimizer!r}")


def _lbfgs_step(params, objective, lbfgs_kwargs, factory=torch.optim.LBFGS):
    """One ``optimizer.step(closure)`` on -objective, then one more forward for the value that is logged."""
    optimizer = factory(params, **lbfgs_kwargs)
    for p in params:
        p.requires_grad = True
    try:
        def closure():
            optimizer.zero_grad()
            loss = -objective()
            loss.backward(retain_graph=True)
            return loss
        optimizer.step(closure)
        bound = objective()
        state = optimizer.state[optimizer._params[0]]
        return bound, int(state["n_iter"]), int(state["func_evals"])
    finally:
        for p in params:
            p.requires_grad = False


def _sum_over_ranks(values, group, device, op=None):
    import torch.distributed as dist
    t = torch.tensor(values, dtype=torch.float64, device=device)
    dist.all_reduce(t, op=op or dist.ReduceOp.SUM, group=group)
    return t.tolist()


def run_step(model, step, lbfgs_kwargs, process_group=None, optimizer="torch", sharded_steps="blockwise"):
    """Runs one conditional maximisation.  Returns ``(bound, niter, nfeval)``; with a process group the bound of a
    block-wise per-trial step is the sum over ranks and niter / nfeval are the maxima over ranks."""
    if sharded_steps not in ("blockwise", "joint"):
        raise ValueError("sharded_steps must be 'blockwise' or 'joint'")
    factory = _optimizer_factory(optimizer)
    params, objective = _parameters_and_objective(model, step)
    if process_group is not None and step in SHARDED_STEPS and sharded_steps == "joint":
        if optimizer == "torch":
            raise ValueError("sharded_steps='joint' needs the shard-aware optimiser (optimizer='b200')")
        # one optimisation over all ranks' vectors: global reductions inside the optimiser, every closure evaluation
        # all-reduced by the model (status included: a failed Cholesky on one rank raises on every rank)
        previous = getattr(model, "shard_mode", None)
        model.shard_mode = "reduce"
        try:
            bound, niter, nfeval = _lbfgs_step(params, objective, lbfgs_kwargs,
                                               lambda p, **kw: factory(p, process_group=process_group, **kw))
        finally:
            model.shard_mode = previous if previous is not None else "auto"
        return float(bound.item()), niter, nfeval
    if process_group is None or step not in SHARDED_STEPS:
        saved = [p.detach().clone() for p in params] if step == "mstep_kernels" else None
        try:
            bound, niter, nfeval = _lbfgs_step(params, objective, lbfgs_kwargs, factory)
        except Exception:
            if saved is not None:                       # a failed kernels step leaves the old hyper-parameters
                for p, q in zip(params, saved):         # (svEM.py:236,249-253)
                    p.detach().copy_(q)
            raise
        return float(bound.item()), niter, nfeval
    import torch.distributed as dist
    device = params[0].device
    failed, err, local = 0.0, None, (0.0, 0, 0)
    try:
        bound, niter, nfeval = _lbfgs_step(params, objective, lbfgs_kwargs, factory)
        if hasattr(model, "checkErrors"):
            model.checkErrors()
        local = (float(bound.item()), niter, nfeval)
    except Exception as e:                              # noqa: BLE001 -- agreed on below, then re-raised everywhere
        failed, err = 1.0, e
    total, n_failed = _sum_over_ranks([local[0], failed], process_group, device)
    if n_failed > 0:
        raise StepFailed(f"{step} failed on {int(n_failed)} rank(s)" + (f": {err}" if err is not None else ""))
    niter, nfeval = _sum_over_ranks([local[1], local[2]], process_group, device, op=dist.ReduceOp.MAX)
    return total, int(niter), int(nfeval)


def maximize(model, optim_params, method="ecm", process_group=None, out=sys.stdout, verbose=True, optimizer="torch",
             sharded_steps="blockwise"):
    """ECM / mECM maximisation of the lower bound.  Returns ``(lower_bound_hist, elapsed_time_hist,
    termination_message, step_log)`` with ``step_log`` rows ``(iteration, step, bound, niter, nfeval)`` -- the
    quantities of the reference's log lines (svEM.py:164-166).  ``process_group`` defaults to the model's."""
    if process_group is None:
        process_group = getattr(model, "_pg", None)
    steps = STEP_ORDER.get(method.lower())
    if steps is None:
        raise ValueError(f"Invalid method={method}. Supported values are ECM and mECM")
    hist = [float(model.eval().item())]                 # lock-step forward: reduced over the shards by the model
    elapsed, t0, log = [0.0], time.time(), []
    for it in range(1, int(optim_params["em_max_iter"]) + 1):
        bound = None
        for step in steps:
            if not optim_params.get(f"{step}_estimate", True):
                continue
            try:
                bound, niter, nfeval = run_step(model, step, optim_params[f"{step}_optim_params"], process_group,
                                                optimizer, sharded_steps)
            except Exception as e:                      # every rank gets here together (see run_step)
                return hist, elapsed, f"Error occured while processing {step} in iteration {it}: {e}", log
            log.append((it, step, bound, niter, nfeval))
            if verbose and out is not None:
                out.write(f"Iteration {it:02d}, {step} end: {bound:f}, niter: {niter:d}, nfeval: {nfeval:d}\n")
        elapsed.append(time.time() - t0)
        if bound is not None:
            hist.append(bound)
    return hist, elapsed, f"Maximum number of iterations ({int(optim_params['em_max_iter'])}) reached", log
