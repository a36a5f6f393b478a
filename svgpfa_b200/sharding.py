"""Trial sharding across GPUs.

Given the shared parameters (C, d, theta) the lower bound is a sum of independent per-trial
terms, so every rank owns a contiguous block of trials -- its per-trial state (m, chol-vecs, Z,
spikes, quadrature nodes) and their gradients never leave the GPU -- and ONE all-reduce(sum) per
evaluation over the packed float64 buffer ``[elbo, ell, kl, term1, term2, 0, 0, 0 | dC | dd |
dtheta]`` (the ``shared`` buffer of ``include/svgpfa_b200.h``) is the only exchange step.
"""
from __future__ import annotations

import numpy as np

SHARED_HDR = 8


def trial_blocks(spikes_per_trial, world_size, per_spike_cost=1.0, per_trial_cost=0.0):
    """Contiguous trial blocks ``[(r0, r1), ...]`` balanced by the estimated cost
    ``per_spike_cost * S_r + per_trial_cost`` (ragged spike counts: BASELINE.json config #4)."""
    s = np.asarray(spikes_per_trial, dtype=np.float64)
    R = s.size
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    cost = per_spike_cost * s + per_trial_cost
    if R == 0 or cost.sum() <= 0:
        cuts = [(R * i) // world_size for i in range(world_size + 1)]
    else:
        cum = np.concatenate([[0.0], np.cumsum(cost)])
        targets = cum[-1] * np.arange(1, world_size) / world_size
        inner = np.searchsorted(cum, targets, side="left")
        # choose the closer of the two neighbouring cut points
        inner = np.array([i if abs(cum[i] - t) <= abs(cum[max(i - 1, 0)] - t) else i - 1
                          for i, t in zip(inner, targets)], dtype=np.int64)
        cuts = np.concatenate([[0], np.maximum.accumulate(np.clip(inner, 0, R)), [R]]).tolist()
    return [(int(cuts[i]), int(cuts[i + 1])) for i in range(world_size)]


GRAD_POSTERIOR, GRAD_EMBEDDING, GRAD_KERNEL, GRAD_INDLOCS = 1, 2, 4, 8      # include/svgpfa_b200.h


def evaluation_is_reduced(grad_flags):
    """Whether an evaluation of a trial-sharded model all-reduces its packed buffer (``shard_mode="auto"``).

    Given the shared parameters (C, d, theta) the bound is a sum of independent per-trial terms.  An optimiser over
    SHARED leaves needs the global value and gradient at every closure call; all ranks receive bit-identical sums,
    take identical decisions and stay in lock-step, so the all-reduce inside the evaluation is safe.  An optimiser
    over PER-TRIAL leaves only (E-step: m, cholVecs; inducing-point M-step: Z) sees the other ranks' terms as
    constants: each rank maximises its own block, with its own number of closure calls, so there must be NO
    collective inside the evaluation (ranks would call it different numbers of times); the caller sums the final
    bounds once per step (``svgpfa_b200.ecm``).  Evaluations that differentiate nothing, or both kinds of leaves
    (one lock-step evaluation per rank, e.g. the benchmark's unit of work), are reduced."""
    sharded = bool(grad_flags & (GRAD_POSTERIOR | GRAD_INDLOCS))
    shared = bool(grad_flags & (GRAD_EMBEDDING | GRAD_KERNEL))
    return shared or not sharded


def trial_costs(spikes_per_trial, N, K, M, Q):
    """Estimated cost of every trial, in nanoseconds on one B200, from the measured stage times of config #5
    (profiles/README.md): the spike kernel is linear in spikes x (latent, inducing point) pairs, the quadrature,
    embedding and inducing-point kernels cost the same for every trial."""
    s = np.asarray(spikes_per_trial, dtype=np.float64)
    per_trial = 0.9e-3 * K * Q * M * M + 0.6e-3 * Q * N * K + 1.8e-3 * K * M ** 3
    return 1.03e-3 * s * K * M + per_trial


def shared_layout(N, K, TH):
    """Offsets of the packed all-reduced buffer."""
    o_c = SHARED_HDR
    o_d = o_c + N * K
    o_t = o_d + N
    return dict(elbo=0, ell=1, kl=2, term1=3, term2=4, C=(o_c, o_d), d=(o_d, o_t), theta=(o_t, o_t + TH),
                length=o_t + TH)


def pack_shared(elbo, ell, kl, dC, dd, dtheta):
    """Host-side counterpart of what svgpfa_finalize writes (used by tests and tools)."""
    N, K = dC.shape
    lay = shared_layout(N, K, dtheta.size)
    buf = np.zeros(lay["length"])
    buf[0], buf[1], buf[2] = elbo, ell, kl
    buf[lay["C"][0]:lay["C"][1]] = dC.reshape(-1)
    buf[lay["d"][0]:lay["d"][1]] = np.asarray(dd).reshape(-1)
    buf[lay["theta"][0]:lay["theta"][1]] = dtheta
    return buf


def all_reduce_shared(shared, group):
    """The exchange step: in-place sum of the packed buffer over the trial shards."""
    import torch.distributed as dist
    dist.all_reduce(shared, op=dist.ReduceOp.SUM, group=group)
    return shared
