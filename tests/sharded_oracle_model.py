"""A trial-sharded model with the protocol of SURVEY.md §8b on top of the CPU oracle (TEST INFRASTRUCTURE).

Stands in for ``B200SVLowerBound(process_group=...)`` in the world-size-2 gloo tests, where there is no GPU: the
numbers come from the oracle's autograd, but everything the multi-rank host logic decides is the PRODUCT's code --
``sharding.evaluation_is_reduced`` (when an evaluation may contain a collective), ``sharding.pack_shared`` /
``all_reduce_shared`` (the one exchange step) -- and the structure is the CUDA model's: value and gradients are
computed eagerly in ``forward`` (one fused pass), the packed buffer is all-reduced, ``backward`` hands the stored
gradients out.
"""
import numpy as np
import torch

import ecm_driver
from oracle import svgpfa_oracle as orc
from svgpfa_b200 import sharding


class _ShardedBoundFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, stats, *leaves):
        K = model.K
        need = ctx.needs_input_grad[2:]
        flags = 0
        flags |= sharding.GRAD_POSTERIOR if any(need[0:2 * K]) else 0
        flags |= sharding.GRAD_EMBEDDING if any(need[2 * K:2 * K + 2]) else 0
        flags |= sharding.GRAD_KERNEL if any(need[2 * K + 2:3 * K + 2]) else 0
        flags |= sharding.GRAD_INDLOCS if any(need[3 * K + 2:4 * K + 2]) else 0
        twins = [p.detach().clone().requires_grad_(n) for p, n in zip(leaves, need)]
        p = dict(m=twins[0:K], chol_vecs=twins[K:2 * K], C=twins[2 * K], d=twins[2 * K + 1],
                 kernel_params=twins[2 * K + 2:3 * K + 2], Z=twins[3 * K + 2:4 * K + 2])
        with torch.enable_grad():
            if stats is None:
                ell, kl, _ = orc.elbo_terms(model.case, p, spike_var=False)
                val = ell - kl
            else:
                flags &= sharding.GRAD_EMBEDDING
                val = orc.ell_from_cached_stats(model.case, stats["allTimes"][0], stats["allTimes"][1],
                                                stats["assocTimes"][0], p["C"], p["d"])
            wanted = [t for t in twins if t.requires_grad]
            grads = torch.autograd.grad(val, wanted, allow_unused=True) if wanted else []
        it = iter(grads)
        full = [(next(it) if t.requires_grad else None) for t in twins]
        full = [torch.zeros_like(t) if (g is None and t.requires_grad) else g for g, t in zip(full, twins)]
        N = model.case["C"].shape[0]
        zeros = lambda shape: np.zeros(shape)
        gnp = lambda g, shape: g.numpy() if g is not None else zeros(shape)
        dth = np.concatenate([gnp(full[2 * K + 2 + k], twins[2 * K + 2 + k].shape).reshape(-1) for k in range(K)])
        shared = torch.from_numpy(sharding.pack_shared(val.item(), 0.0, 0.0, gnp(full[2 * K], (N, K)),
                                                       gnp(full[2 * K + 1], (N,)), dth))
        model.n_evals += 1
        if model.pg is not None and (model.shard_mode == "reduce" or
                                     (model.shard_mode == "auto" and sharding.evaluation_is_reduced(flags))):
            sharding.all_reduce_shared(shared, model.pg)
            model.n_reduced += 1
        lay = sharding.shared_layout(N, K, dth.size)
        if full[2 * K] is not None:
            full[2 * K] = shared[lay["C"][0]:lay["C"][1]].view(N, K).clone()
        if full[2 * K + 1] is not None:
            full[2 * K + 1] = shared[lay["d"][0]:lay["d"][1]].view(twins[2 * K + 1].shape).clone()
        off = lay["theta"][0]
        for k in range(K):
            n = twins[2 * K + 2 + k].numel()
            if full[2 * K + 2 + k] is not None:
                full[2 * K + 2 + k] = shared[off:off + n].clone()
            off += n
        ctx.grads = full
        return shared[0].clone()

    @staticmethod
    def backward(ctx, grad_out):
        return (None, None, *[None if g is None else g * grad_out for g in ctx.grads])


class ShardedOracleModel(ecm_driver.OracleModel):
    """Trials [r0, r1) of ``case`` on this rank; ``pg=None``: a single-process model with the same structure."""

    def __init__(self, case, pg=None):
        super().__init__(case)
        self.pg, self._pg = pg, pg
        self.K = len(case["kernel_types"])
        self.n_evals = self.n_reduced = 0
        self.shard_mode = "auto"                 # B200SVLowerBound.shard_mode: "reduce" = every evaluation is a collective

    def _leaves(self):
        p = self.p
        return list(p["m"]) + list(p["chol_vecs"]) + [p["C"], p["d"]] + list(p["kernel_params"]) + list(p["Z"])

    def eval(self):
        return _ShardedBoundFn.apply(self, None, *self._leaves())

    def evalELLSumAcrossTrialsAndNeurons(self, svPosteriorOnLatentsStats):
        return _ShardedBoundFn.apply(self, svPosteriorOnLatentsStats, *self._leaves())
