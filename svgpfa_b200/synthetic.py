"""Seeded synthetic svGPFA problem instances.

One generator feeds the CPU checker, the golden-vector script, the parity tests and
``bench.py`` so that every arm sees the same numbers.  A *case* is a plain dict
of numpy arrays / python lists with the field names the reference's
``initial_params`` dictionary uses (``/root/reference/src/svGPFA/utils/initUtils.py:468-481``):

    kernel_types   list[K] of "expquad" | "periodic"
    kernel_params  list[K] of float64 arrays, (1,) = [lengthscale] or (2,) = [lengthscale, period]
    Z              list[K] of (R, M_k, 1)   inducing-point locations
    m              list[K] of (R, M_k, 1)   variational means
    chol_vecs      list[K] of (R, P_k, 1)   row-major lower-triangular entries of Ls
    C (N, K), d (N, 1)                      linear embedding
    leg_quad_points / leg_quad_weights      (R, Q, 1)
    spike_times    (S,) float64 or float32, trial-major then neuron-major (the order
                   ``PointProcessELL.__stackSpikeTimes`` produces,
                   ``/root/reference/src/svGPFA/stats/expectedLogLikelihood.py:157-173``)
    spike_counts   (R, N) int64             spikes of neuron n in trial r
    reg            float                    prior-covariance regulariser
"""
from __future__ import annotations

import numpy as np

# name -> (R, N, K, M, Q, mixed kernels, heavy ragged)   BASELINE.json "configs"
CONFIGS = {
    "tiny":    dict(R=4,     N=7,   K=3,  M=5,  Q=16,  mixed=True,  ragged=False),
    "config2": dict(R=200,   N=100, K=3,  M=9,  Q=200, mixed=False, ragged=False),
    "config3": dict(R=2000,  N=200, K=10, M=20, Q=200, mixed=True,  ragged=False),
    "config4": dict(R=5000,  N=300, K=10, M=32, Q=200, mixed=False, ragged=True),
    "config5": dict(R=20000, N=500, K=20, M=32, Q=200, mixed=False, ragged=False),
    # north_star "M up to 64": config #5's per-trial shape with 64 inducing points (the M > 32 kernels)
    "m64":     dict(R=4000,  N=500, K=20, M=64, Q=200, mixed=False, ragged=False),
}


def leg_quad(Q: int, a: float, b: float):
    """Gauss-Legendre nodes/weights on [a, b]; stands in for the un-vendored
    ``gcnu_common.numerical_methods.utils.leggaussVarLimits`` called at
    ``/root/reference/src/svGPFA/utils/miscUtils.py:234``."""
    x, w = np.polynomial.legendre.leggauss(Q)
    return 0.5 * (b - a) * x + 0.5 * (b + a), 0.5 * (b - a) * w


def tril_size(M: int) -> int:
    return M * (M + 1) // 2


def make_params(R, N, K, M, Q, *, mixed=False, T=1.0, seed=0, reg=1e-3, M_list=None,
                d_2d=True):
    """Parameters only (no spikes).  ``M_list`` allows heterogeneous M_k."""
    rng = np.random.default_rng(seed)
    M_list = list(M_list) if M_list is not None else [M] * K
    kernel_types, kernel_params, Z, m, chol_vecs = [], [], [], [], []
    for k in range(K):
        Mk = M_list[k]
        if mixed and (k % 2 == 1):
            kernel_types.append("periodic")
            kernel_params.append(np.array([1.0 + 0.1 * k, 0.5 + 0.05 * k]))
        else:
            kernel_types.append("expquad")
            kernel_params.append(np.array([0.1 + 0.05 * k]))
        base = np.linspace(0.0, T, Mk)
        Z.append((base[None, :] + rng.uniform(-0.1, 0.1, size=(R, Mk)) * T / Mk)[:, :, None])
        m.append(rng.standard_normal((R, Mk, 1)))
        Pk = tril_size(Mk)
        rows, cols = np.tril_indices(Mk)
        diag_mask = (rows == cols).astype(np.float64)
        chol_vecs.append((0.1 * diag_mask[None, :] + 0.01 * rng.standard_normal((R, Pk)))[:, :, None])
    rate = rng.uniform(5.0, 35.0, size=N)
    C = 0.3 * rng.standard_normal((N, K))
    d = np.log(rate) + 0.1 * rng.standard_normal(N)
    d = d[:, None] if d_2d else d
    x, w = leg_quad(Q, 0.0, T)
    return dict(
        kernel_types=kernel_types, kernel_params=kernel_params, Z=Z, m=m, chol_vecs=chol_vecs,
        C=C, d=d,
        leg_quad_points=np.repeat(x[None, :, None], R, axis=0),
        leg_quad_weights=np.repeat(w[None, :, None], R, axis=0),
        reg=float(reg), T=float(T), rate=rate,
    )


def make_spikes(R, N, *, T=1.0, seed=0, ragged=False, rate=None, dtype=np.float64):
    """Poisson counts per (trial, neuron) and sorted-uniform spike times, flattened in
    trial-major / neuron-major order."""
    rng = np.random.default_rng(seed + 7919)
    if ragged:
        rate_n = np.exp(rng.uniform(np.log(1.0), np.log(200.0), size=N))
        g_r = rng.lognormal(0.0, 0.5, size=R)
    else:
        rate_n = rate if rate is not None else rng.uniform(5.0, 35.0, size=N)
        g_r = np.ones(R)
    counts = rng.poisson(g_r[:, None] * rate_n[None, :] * T).astype(np.int64)
    S = int(counts.sum())
    times = rng.uniform(0.0, T, size=S)
    # sort within each (trial, neuron) segment: add the segment id (times < T) and sort once
    seg = np.repeat(np.arange(R * N, dtype=np.float64), counts.reshape(-1))
    order = np.argsort(seg * (2.0 * T) + times, kind="stable")
    times = times[order]
    return times.astype(dtype), counts


def make_case(name_or_cfg="tiny", *, seed=0, reg=1e-3, R=None, spike_dtype=np.float64,
              M_list=None, d_2d=True):
    cfg = dict(CONFIGS[name_or_cfg]) if isinstance(name_or_cfg, str) else dict(name_or_cfg)
    if R is not None:
        cfg["R"] = R
    case = make_params(cfg["R"], cfg["N"], cfg["K"], cfg["M"], cfg["Q"], mixed=cfg["mixed"],
                       seed=seed, reg=reg, M_list=M_list, d_2d=d_2d)
    times, counts = make_spikes(cfg["R"], cfg["N"], T=case["T"], seed=seed, ragged=cfg["ragged"],
                                rate=case["rate"], dtype=spike_dtype)
    case["spike_times"] = times
    case["spike_counts"] = counts
    return case


def nested_spikes(case):
    """``measurements[r][n]`` as the reference's ``setMeasurements`` expects it
    (``/root/reference/src/svGPFA/stats/svLowerBound.py:16-24``)."""
    counts = case["spike_counts"]
    R, N = counts.shape
    pieces = np.split(case["spike_times"], np.cumsum(counts.reshape(-1))[:-1])
    return [[pieces[r * N + n] for n in range(N)] for r in range(R)]


def slice_trials(case, r0, r1):
    """The sub-problem made of trials [r0, r1) (trial sharding)."""
    counts = case["spike_counts"]
    per_trial = counts.sum(axis=1)
    off = np.concatenate([[0], np.cumsum(per_trial)])
    out = dict(case)
    for key in ("Z", "m", "chol_vecs"):
        out[key] = [a[r0:r1] for a in case[key]]
    for key in ("leg_quad_points", "leg_quad_weights"):
        out[key] = case[key][r0:r1]
    out["spike_counts"] = counts[r0:r1]
    out["spike_times"] = case["spike_times"][off[r0]:off[r1]]
    return out


_LIST_KEYS = ("kernel_params", "Z", "m", "chol_vecs")


def save_case(path, case, extra=None):
    flat = {}
    K = len(case["kernel_types"])
    flat["kernel_types"] = np.array(case["kernel_types"])
    for key in _LIST_KEYS:
        for k in range(K):
            flat[f"{key}_{k}"] = np.asarray(case[key][k])
    for key in ("C", "d", "leg_quad_points", "leg_quad_weights", "spike_times", "spike_counts"):
        flat[key] = np.asarray(case[key])
    flat["reg"] = np.array(case["reg"])
    for k, v in (extra or {}).items():
        flat["out_" + k] = np.asarray(v)
    np.savez_compressed(path, **flat)


def load_case(path):
    z = np.load(path, allow_pickle=False)
    kt = [str(s) for s in z["kernel_types"]]
    K = len(kt)
    case = dict(kernel_types=kt)
    for key in _LIST_KEYS:
        case[key] = [z[f"{key}_{k}"] for k in range(K)]
    for key in ("C", "d", "leg_quad_points", "leg_quad_weights", "spike_times", "spike_counts"):
        case[key] = z[key]
    case["reg"] = float(z["reg"])
    extra = {k[4:]: z[k] for k in z.files if k.startswith("out_")}
    return case, extra


GEN_BLOCK = 32        # trials per random-stream block of the device generator


def _block_generator(device, seed, block, stream):
    """One torch.Generator per (seed, block of GEN_BLOCK trials, stream): what a trial holds depends on
    (seed, trial index) alone, never on how the trials are cut into shards."""
    import torch
    gen = torch.Generator(device=device)
    gen.manual_seed(((seed * 1000003 + block) * 4 + stream) * 2 + 1)
    return gen


def _shared_params(cfg, seed):
    """(rate, C, d) -- the quantities every trial shares; functions of ``seed`` only."""
    N, K = cfg["N"], cfg["K"]
    shared = np.random.default_rng(seed)
    if cfg["ragged"]:
        rate = np.exp(shared.uniform(np.log(1.0), np.log(200.0), size=N))
        d = np.log(shared.uniform(5.0, 35.0, size=N)) + 0.1 * shared.standard_normal(N)
    else:
        rate = shared.uniform(5.0, 35.0, size=N)
        d = np.log(rate) + 0.1 * shared.standard_normal(N)
    C = 0.3 * shared.standard_normal((N, K))
    return rate, C, d


def _block_counts(cfg, device, seed, block, rate_t, T):
    """Spike counts (rows, N) int64 of the trials of one generator block."""
    import torch
    rows = min(GEN_BLOCK, cfg["R"] - block * GEN_BLOCK)
    gen = _block_generator(device, seed, block, 0)
    f64 = dict(dtype=torch.float64, device=device)
    if cfg["ragged"]:
        g_r = torch.exp(0.5 * torch.randn(rows, generator=gen, **f64))
    else:
        g_r = torch.ones(rows, **f64)
    return torch.poisson(g_r[:, None] * rate_t[None, :] * T, generator=gen).to(torch.int64)


def spike_counts_torch(cfg, device, *, seed=0, T=1.0):
    """Spike counts (R, N) of EVERY trial of the configuration (cheap: no spike times), so that each rank of a
    trial-sharded run can compute the same cost-balanced trial blocks before generating its own shard."""
    import torch
    cfg = dict(CONFIGS[cfg]) if isinstance(cfg, str) else dict(cfg)
    rate, _, _ = _shared_params(cfg, seed)
    rate_t = torch.tensor(rate, dtype=torch.float64, device=device)
    nblk = -(-cfg["R"] // GEN_BLOCK)
    return torch.cat([_block_counts(cfg, device, seed, b, rate_t, T) for b in range(nblk)], 0)


def make_case_torch(cfg, device, *, seed=0, reg=1e-3, r0=0, r1=None, T=1.0):
    """Device-side generator for the large benchmark workloads (same distributions as ``make_case``;
    different random stream).  Generates trials [r0, r1) of the configuration so that every rank of a
    trial-sharded run builds only its own shard.  Random streams are drawn per block of GEN_BLOCK trials
    from generators seeded by (seed, block index), so trial r holds the same numbers whatever the shard
    boundaries are: an N-rank run evaluates exactly the data of the 1-rank run.  The shared parameters
    (C, d, theta, per-neuron rates) depend on ``seed`` alone.  Returns a case dict of torch tensors on
    ``device`` (spike_times float64 (S,), spike_counts int64 (R_local, N))."""
    import torch
    cfg = dict(CONFIGS[cfg]) if isinstance(cfg, str) else dict(cfg)
    R, N, K, M, Q = cfg["R"], cfg["N"], cfg["K"], cfg["M"], cfg["Q"]
    r1 = R if r1 is None else r1
    Rl = r1 - r0
    rate, C, d = _shared_params(cfg, seed)
    f64 = dict(dtype=torch.float64, device=device)
    P = tril_size(M)
    rows_i, cols_i = np.tril_indices(M)
    diag = torch.tensor((rows_i == cols_i).astype(np.float64), **f64)
    base = torch.linspace(0.0, T, M, **f64)
    rate_t = torch.tensor(rate, **f64)
    Zb, mb, cb, cnt_b, times_b = [], [], [], [], []
    for b in range(r0 // GEN_BLOCK, -(-r1 // GEN_BLOCK) if Rl > 0 else r0 // GEN_BLOCK):
        t0 = b * GEN_BLOCK
        rows = min(GEN_BLOCK, R - t0)
        lo, hi = max(r0, t0) - t0, min(r1, t0 + rows) - t0          # rows of the block inside [r0, r1)
        gen = _block_generator(device, seed, b, 1)
        jit = (torch.rand(rows, K, M, generator=gen, **f64) * 0.2 - 0.1) * T / M
        Zb.append((base[None, None, :] + jit)[lo:hi])
        mb.append(torch.randn(rows, K, M, generator=gen, **f64)[lo:hi])
        cb.append((0.1 * diag[None, None, :] + 0.01 * torch.randn(rows, K, P, generator=gen, **f64))[lo:hi])
        counts = _block_counts(cfg, device, seed, b, rate_t, T)
        S = int(counts.sum().item())
        seg = torch.repeat_interleave(torch.arange(rows * N, device=device), counts.reshape(-1), output_size=S)
        gen_t = _block_generator(device, seed, b, 2)
        key = seg.to(torch.float64) + torch.rand(S, generator=gen_t, **f64).clamp_(max=1.0 - 1e-9)
        key, _ = torch.sort(key)
        off = torch.cat([torch.zeros(1, dtype=torch.int64, device=device), counts.sum(1).cumsum(0)]).tolist()
        times_b.append(((key - torch.floor(key)) * T)[off[lo]:off[hi]])
        cnt_b.append(counts[lo:hi])
    cat = lambda xs, shape: torch.cat(xs, 0) if xs else torch.zeros(shape, **f64)
    Zall, mall, call = cat(Zb, (0, K, M)), cat(mb, (0, K, M)), cat(cb, (0, K, P))
    kernel_types, kernel_params = [], []
    for k in range(K):
        if cfg["mixed"] and (k % 2 == 1):
            kernel_types.append("periodic")
            kernel_params.append(torch.tensor([1.0 + 0.1 * k, 0.5 + 0.05 * k], **f64))
        else:
            kernel_types.append("expquad")
            kernel_params.append(torch.tensor([0.1 + 0.05 * k], **f64))
    Z = [Zall[:, k, :, None].contiguous() for k in range(K)]
    m = [mall[:, k, :, None].contiguous() for k in range(K)]
    chol_vecs = [call[:, k, :, None].contiguous() for k in range(K)]
    del Zall, mall, call
    x, w = leg_quad(Q, 0.0, T)
    tq = torch.tensor(x, **f64)[None, :, None].repeat(Rl, 1, 1)
    wq = torch.tensor(w, **f64)[None, :, None].repeat(Rl, 1, 1)
    counts = torch.cat(cnt_b, 0) if cnt_b else torch.zeros((0, N), dtype=torch.int64, device=device)
    times = torch.cat(times_b) if times_b else torch.zeros(0, **f64)
    return dict(kernel_types=kernel_types, kernel_params=kernel_params, Z=Z, m=m, chol_vecs=chol_vecs,
                C=torch.tensor(C, **f64), d=torch.tensor(d, **f64)[:, None].contiguous(),
                leg_quad_points=tq, leg_quad_weights=wq, spike_times=times, spike_counts=counts,
                reg=float(reg), T=float(T))


def case_to_numpy(case, r0=0, r1=None):
    """Host copy of trials [r0, r1) of a (torch or numpy) case -- the bounded CPU-baseline sample."""
    import torch
    to = lambda a: a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    counts = to(case["spike_counts"])
    r1 = counts.shape[0] if r1 is None else r1
    off = np.concatenate([[0], np.cumsum(counts.sum(axis=1))])
    out = dict(kernel_types=list(case["kernel_types"]), kernel_params=[to(a) for a in case["kernel_params"]],
               C=to(case["C"]), d=to(case["d"]), reg=case["reg"])
    for key in ("Z", "m", "chol_vecs"):
        out[key] = [to(a[r0:r1]) for a in case[key]]
    for key in ("leg_quad_points", "leg_quad_weights"):
        out[key] = to(case[key][r0:r1])
    out["spike_counts"] = counts[r0:r1]
    out["spike_times"] = to(case["spike_times"][int(off[r0]):int(off[r1])])
    return out
