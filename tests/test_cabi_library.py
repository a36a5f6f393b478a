"""CPU checks of the C-ABI library: it builds, loads, and exports every symbol that
include/svgpfa_b200.h declares; host-only entry points behave.  No GPU compute here."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from svgpfa_b200 import _cabi


@pytest.fixture(scope="module")
def lib():
    from svgpfa_b200 import build
    build.build()
    return _cabi.lib()


def _declared_symbols(header="svgpfa_b200.h"):
    text = open(os.path.join(ROOT, "include", header)).read()
    return sorted(set(re.findall(r"\b(svgpfa_[a-z_0-9]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _cabi.SYMBOLS, f"{name} has no ctypes signature"
    assert sorted(_cabi.SYMBOLS) == declared


def test_probes_live_in_their_own_library(lib):
    """Measurement probes and test hooks (include/svgpfa_b200_probes.h) are exported by the probes library and by
    it alone: the product library carries no probe kernel."""
    probes = _cabi.probes()
    declared = _declared_symbols("svgpfa_b200_probes.h")
    assert sorted(_cabi.PROBE_SYMBOLS) == declared
    for name in declared:
        assert hasattr(probes, name)
        assert not hasattr(lib, name), f"{name} leaked into the product library"


def test_abi_constants_match_header(lib):
    text = open(os.path.join(ROOT, "include", "svgpfa_b200.h")).read()
    get = lambda n: int(re.search(rf"#define {n}\s+(\d+)", text).group(1))
    assert lib.svgpfa_abi_version() == get("SVGPFA_ABI_VERSION") == _cabi.ABI_VERSION
    assert get("SVGPFA_MAX_M") == _cabi.MAX_M
    assert get("SVGPFA_EMBED_TN") == _cabi.EMBED_TN
    assert get("SVGPFA_SHARED_HDR") == _cabi.SHARED_HDR
    assert get("SVGPFA_TERM1_SLOTS") == _cabi.TERM1_SLOTS
    # struct layouts: field counts of the ctypes mirrors equal the header's
    body = re.search(r"typedef struct svgpfa_buffers \{(.*?)\} svgpfa_buffers;", text, re.S).group(1)
    fields = re.findall(r"\*\s*([A-Za-z_0-9]+);", body)
    assert tuple(fields) == _cabi.BUFFER_FIELDS
    assert ctypes.sizeof(_cabi.LatentDesc) == 32
    assert ctypes.sizeof(_cabi.Dims) == 104
    assert get("SVGPFA_FIN_SLOTS") == _cabi.FIN_SLOTS


def test_build_segments_host(lib):
    rng = np.random.default_rng(0)
    R, N = 7, 13
    counts = rng.poisson(3.0, size=(R, N)).astype(np.int64)
    counts[2, :] = 0
    counts[:, 5] = 0
    S = int(counts.sum())
    seg = np.empty(R * N + 1, dtype=np.int64)
    idx = np.empty(S, dtype=np.int64)
    rc = lib.svgpfa_build_segments_host(R, N, counts.ctypes.data, seg.ctypes.data, idx.ctypes.data)
    assert rc == 0
    assert np.array_equal(seg, np.concatenate([[0], np.cumsum(counts.reshape(-1))]))
    # the per-spike neuron index the reference builds (expectedLogLikelihood.py:168-172)
    want = np.concatenate([np.repeat(np.arange(N), counts[r]) for r in range(R)])
    assert np.array_equal(idx, want)
    bad = counts.copy()
    bad[0, 0] = -1
    assert lib.svgpfa_build_segments_host(R, N, bad.ctypes.data, seg.ctypes.data, None) != 0
    assert b"negative" in lib.svgpfa_last_error()


def test_no_cpu_fallback():
    """The product path refuses to run without a CUDA device instead of silently using the CPU."""
    import torch
    from svgpfa_b200 import B200SVLowerBound
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from svgpfa_b200 import synthetic
    from svgpfa_b200.kernels import build_kernels
    from svgpfa_b200.testing import initial_params_from_case
    case = synthetic.make_case("tiny")
    model = B200SVLowerBound(kernels=build_kernels(case["kernel_types"]))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model.setInitialParams(initial_params_from_case(case))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        B200SVLowerBound(device="cpu")._dev()


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "svgpfa_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|import_module\(.oracle", src, re.M), \
                    f"{f} imports the oracle"
                assert "oracle" not in src, f"{f} mentions the oracle"
