"""Host-side logic of the model object that needs no kernel launch: packing, the aliasing between the leaf tensors
the optimiser mutates and the packed buffers the kernels read, pickling (svEM.py:89-92,175-181 pickle the whole
model after every step), staleness of derived state.  Runs on CPU by pointing the model's device at host memory --
only tensor bookkeeping is exercised; any evaluation would still need the CUDA library and a GPU."""
import pickle

import numpy as np
import pytest
import torch

from svgpfa_b200 import synthetic
from svgpfa_b200.kernels import build_kernels
from svgpfa_b200.model import B200SVLowerBound
from svgpfa_b200.testing import initial_params_from_case


@pytest.fixture
def host_model(monkeypatch):
    monkeypatch.setattr(B200SVLowerBound, "_dev", lambda self: torch.device("cpu"))
    case = synthetic.make_case("tiny", seed=1, M_list=[5, 4, 6])
    kernels = build_kernels(case["kernel_types"])
    model = B200SVLowerBound(kernels=kernels)
    model.setInitialParams(initial_params_from_case(case))
    return model, case


def test_getters_alias_the_packed_buffers(host_model):
    model, case = host_model
    K = len(case["kernel_types"])
    post = model.getSVPosteriorOnIndPointsParams()
    assert [tuple(p.shape) for p in post[:K]] == [(4, 5, 1), (4, 4, 1), (4, 6, 1)]
    before = model._param_versions()
    with torch.no_grad():
        model.getIndPointsLocs()[1].add_(1.0)
        model.getKernelsParams()[2].mul_(2.0)
    assert model._param_versions() != before                        # the Kzz cache key notices
    R = 4
    z1 = model._Zbuf[R * model._moff[1]:R * model._moff[2]].view(R, 4, 1)
    assert np.allclose(z1.numpy(), case["Z"][1] + 1.0)
    assert model.getKernels()[2].getParams().data_ptr() == model.getKernelsParams()[2].data_ptr()


def test_pickle_round_trip_keeps_aliasing(host_model):
    """After pickle.loads an in-place update of a getter tensor must reach the packed buffer (plain pickle does not
    preserve view/base sharing: the views are rebuilt in __setstate__), the cache keys must notice it, the kernel
    objects must alias the packed parameters again and requires_grad flags must survive."""
    model, case = host_model
    for p in model.getSVPosteriorOnIndPointsParams():
        p.requires_grad_(True)
    twin = pickle.loads(pickle.dumps(model))
    assert [p.requires_grad for p in twin.getSVPosteriorOnIndPointsParams()] == [True] * 6
    assert not any(p.requires_grad for p in twin.getIndPointsLocs())
    for a, b in zip(model._leaves(), twin._leaves()):
        assert torch.equal(a.detach(), b.detach())
    v0 = twin._param_versions()
    with torch.no_grad():
        twin.getSVPosteriorOnIndPointsParams()[0].add_(0.5)          # what LBFGS does (torch optim/lbfgs.py:306-323)
        twin.getIndPointsLocs()[2].copy_(torch.zeros(4, 6, 1, dtype=torch.float64))
        twin.getKernelsParams()[0].fill_(7.0)
    assert np.allclose(twin._mbuf[:20].numpy().reshape(4, 5, 1), case["m"][0] + 0.5)
    assert float(twin._Zbuf[4 * twin._moff[2]:].abs().sum()) == 0.0
    assert float(twin._thbuf[0]) == 7.0 and twin._param_versions() != v0
    assert twin.getKernels()[0].getParams().data_ptr() == twin._thbuf.data_ptr()
    assert type(twin.getKernels()[1]).__name__ == "PeriodicKernel"
    # the original is untouched
    assert np.allclose(model._mbuf[:20].numpy().reshape(4, 5, 1), case["m"][0])
    # a second generation pickles too
    again = pickle.loads(pickle.dumps(twin))
    assert float(again._thbuf[0]) == 7.0


def test_replacing_kernels_rebinds_parameters(host_model):
    model, case = host_model
    fresh = build_kernels(case["kernel_types"])
    model.setKernels(fresh)
    assert fresh[0].getParams().data_ptr() == model._thbuf.data_ptr()
    with pytest.raises(ValueError):
        model.setKernels(build_kernels(["periodic", "periodic", "periodic"]))


def test_regulariser_change_invalidates_derived_state(host_model):
    model, _ = host_model
    model.setPriorCovRegParam(1e-3)
    model._kzz_key = ("stale",)
    model.setPriorCovRegParam(1e-2)
    assert model._kzz_key is None and model._reg == 1e-2


def test_shard_mode_rules():
    from svgpfa_b200 import sharding
    P, E, T, Z = 1, 2, 4, 8
    assert sharding.evaluation_is_reduced(0)                         # forward only: one lock-step call per rank
    assert sharding.evaluation_is_reduced(E) and sharding.evaluation_is_reduced(T)
    assert sharding.evaluation_is_reduced(P | E | T | Z)             # the benchmark's unit of work
    assert not sharding.evaluation_is_reduced(P)                     # E-step: per-rank optimiser, no collective
    assert not sharding.evaluation_is_reduced(Z) and not sharding.evaluation_is_reduced(P | Z)
    with pytest.raises(ValueError):
        B200SVLowerBound(shard_mode="sometimes")
