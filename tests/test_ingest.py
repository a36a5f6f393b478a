"""Measurement ingestion (svgpfa_b200/ingest.py): the flat CSR layout against the reference's stacking
(stats/expectedLogLikelihood.py:157-173, restated in oracle.stack_spike_times) -- bit-exact times and indices,
for lists, numpy arrays and torch tensors, float32 and float64, empty segments and empty trials."""
import os
import time

import numpy as np
import pytest
import torch

from conftest import GOLDEN, golden_names
from oracle import svgpfa_oracle as orc
from svgpfa_b200 import ingest, synthetic


def _as(kind, a):
    if kind == "list":
        return [float(x) for x in a]
    if kind == "numpy32":
        return np.asarray(a, dtype=np.float32)
    if kind == "torch32":
        return torch.as_tensor(np.asarray(a, dtype=np.float32))
    if kind == "torch64":
        return torch.as_tensor(np.asarray(a, dtype=np.float64))
    return np.asarray(a, dtype=np.float64)


@pytest.mark.parametrize("kind", ["list", "numpy64", "numpy32", "torch64", "torch32", "mixed"])
@pytest.mark.parametrize("name", ["tiny_mixed", "tiny_empty", "config2_r8"])
def test_stacking_matches_reference_order(name, kind):
    case, ref = synthetic.load_case(os.path.join(GOLDEN, name + ".npz"))
    nested = synthetic.nested_spikes(case)
    kinds = ["list", "numpy64", "numpy32", "torch64", "torch32"]
    conv = [[_as(kinds[(r + n) % 5] if kind == "mixed" else kind, s) for n, s in enumerate(tr)]
            for r, tr in enumerate(nested)]
    times, counts, idx = ingest.stack_spike_times(conv, with_neuron_index=True)
    assert times.dtype == np.float64 and counts.dtype == np.int64 and idx.dtype == np.int64
    assert np.array_equal(counts, case["spike_counts"])
    # the reference's own stacking of the same nested lists (float64 tensors)
    t_ref, i_ref = orc.stack_spike_times([[torch.as_tensor(np.asarray(s, dtype=np.float64)) for s in tr] for tr in nested])
    t_ref = np.concatenate([t.numpy() for t in t_ref]) if len(t_ref) else np.zeros(0)
    i_ref = np.concatenate([i.numpy() for i in i_ref]) if len(i_ref) else np.zeros(0, dtype=np.int64)
    assert np.array_equal(idx, i_ref)
    if "stacked_neuron_index" in ref:
        assert np.array_equal(idx, ref["stacked_neuron_index"])
    exact = kind in ("list", "numpy64", "torch64")
    if exact:
        assert np.array_equal(times, t_ref)
    else:       # float32 storage: promoted exactly, i.e. equal to the float32-rounded reference times
        want = t_ref.copy()
        if kind == "mixed":
            pos = 0
            for r, tr in enumerate(nested):
                for n, s in enumerate(tr):
                    k = kinds[(r + n) % 5]
                    if k.endswith("32"):
                        want[pos:pos + len(s)] = want[pos:pos + len(s)].astype(np.float32).astype(np.float64)
                    pos += len(s)
        else:
            want = want.astype(np.float32).astype(np.float64)
        assert np.array_equal(times, want)


def test_ragged_shapes_and_errors():
    assert ingest.stack_spike_times([])[0].size == 0
    t, c = ingest.stack_spike_times([[[], []], [[], []]])
    assert t.size == 0 and c.shape == (2, 2) and c.sum() == 0
    t, c, i = ingest.stack_spike_times([[[0.3, 0.1], []], [[], [0.2]]], with_neuron_index=True)
    assert list(t) == [0.3, 0.1, 0.2] and c.tolist() == [[2, 0], [0, 1]] and list(i) == [0, 0, 1]   # order kept, not sorted
    with pytest.raises(ValueError):
        ingest.stack_spike_times([[[0.1]], [[0.1], [0.2]]])


def test_ingestion_cost_is_linear_and_small():
    """2e5 (trial, neuron) pairs of torch tensors in well under the reference's per-pair cost (~10 us)."""
    rng = np.random.default_rng(0)
    R, N = 400, 500
    nested = [[torch.as_tensor(rng.uniform(0, 1, size=rng.integers(0, 6))) for _ in range(N)] for _ in range(R)]
    t0 = time.perf_counter()
    times, counts = ingest.stack_spike_times(nested)
    dt = time.perf_counter() - t0
    assert counts.sum() == times.size
    assert dt < 5.0, dt
