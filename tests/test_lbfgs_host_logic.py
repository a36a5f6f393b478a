"""Decision logic of ``svgpfa_b200.lbfgs.LBFGS`` against ``torch.optim.LBFGS`` (the optimiser the reference builds in
stats/svEM.py:218-264) on CPU, with the torch implementation of the vector primitives injected."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(__file__))
from vector_ops_torch import TorchVectorOps  # noqa: E402

from svgpfa_b200.lbfgs import LBFGS  # noqa: E402

F64 = torch.float64


def rosenbrock(xs):
    x = torch.cat([p.reshape(-1) for p in xs])
    return (100.0 * (x[1:] - x[:-1] ** 2) ** 2 + (1.0 - x[:-1]) ** 2).sum()


def illcond(xs):
    x = torch.cat([p.reshape(-1) for p in xs])
    w = torch.logspace(0, 4, x.numel(), dtype=F64)
    return 0.5 * (w * x * x).sum() + 0.1 * torch.cos(3.0 * x).sum() + (x[:-1] * x[1:]).sum()


def make_params(n_leaves, numel, seed, packed):
    g = torch.Generator().manual_seed(seed)
    if packed:                                  # leaves = views of one buffer, like the model's getters
        buf = torch.randn(n_leaves * numel, generator=g, dtype=F64) * 0.5
        return buf, [buf[i * numel:(i + 1) * numel].view(numel, 1) for i in range(n_leaves)]
    return None, [torch.randn(numel, 1, generator=g, dtype=F64) * 0.5 for _ in range(n_leaves)]


def run(opt_cls, fn, params, n_steps, **kw):
    for p in params:
        p.requires_grad = True
    opt = opt_cls(params, **kw)
    losses = []

    def closure():
        opt.zero_grad()
        loss = fn(params)
        loss.backward()
        losses.append(loss.item())
        return loss
    for _ in range(n_steps):
        opt.step(closure)
    st = opt.state[opt._params[0]]
    return losses, st["n_iter"], st["func_evals"], [p.detach().clone() for p in params]


@pytest.mark.parametrize("fn", [rosenbrock, illcond])
@pytest.mark.parametrize("line_search_fn", ["strong_wolfe", None])
@pytest.mark.parametrize("packed", [True, False])
def test_same_trajectory_as_torch_lbfgs(fn, line_search_fn, packed):
    # (Rosenbrock amplifies rounding differences ~10x per 8 evaluations: keep its run short)
    kw = dict(lr=1.0 if line_search_fn else 1e-4, max_iter=12 if fn is rosenbrock else 40, tolerance_grad=1e-9,
              tolerance_change=1e-12, line_search_fn=line_search_fn, history_size=7)
    _, pa = make_params(3, 5, 0, packed)
    _, pb = make_params(3, 5, 0, packed)
    la, ia, ea, xa = run(torch.optim.LBFGS, fn, pa, 2, **kw)
    lb, ib, eb, xb = run(lambda p, **k: LBFGS(p, ops=TorchVectorOps(), **k), fn, pb, 2, **kw)
    assert (ia, ea) == (ib, eb)
    assert len(la) == len(lb)
    for a, b in zip(la, lb):                    # every closure call saw the same point
        assert abs(a - b) <= 1e-10 * max(1.0, abs(a))
    for a, b in zip(xa, xb):
        assert torch.allclose(a, b, rtol=1e-8, atol=1e-10)


def test_history_shift_and_rejected_pairs():
    # history_size 2 forces the shift on nearly every iteration; the concave bumps make some y.s <= 1e-10
    kw = dict(lr=1.0, max_iter=60, tolerance_grad=1e-10, tolerance_change=1e-14, line_search_fn="strong_wolfe",
              history_size=2)
    _, pa = make_params(2, 6, 3, True)
    _, pb = make_params(2, 6, 3, True)
    la, ia, ea, _ = run(torch.optim.LBFGS, illcond, pa, 1, **kw)
    lb, ib, eb, _ = run(lambda p, **k: LBFGS(p, ops=TorchVectorOps(), **k), illcond, pb, 1, **kw)
    assert (ia, ea) == (ib, eb)
    assert abs(la[-1] - lb[-1]) <= 1e-9 * max(1.0, abs(la[-1]))


def test_writes_reach_the_packed_buffer_and_bump_its_version():
    buf, params = make_params(3, 4, 1, True)
    v0 = buf._version
    before = buf.clone()
    run(lambda p, **k: LBFGS(p, ops=TorchVectorOps(), **k), rosenbrock, params, 1, max_iter=3,
        line_search_fn="strong_wolfe")
    assert not torch.equal(buf, before)         # the optimiser wrote through the segment view
    assert buf._version > v0                    # and the model's cache keys (base version counters) moved
    assert len(LBFGS(params, ops=TorchVectorOps())._segments) == 1


def test_optimal_start_returns_without_iterating():
    p = [torch.zeros(3, 1, dtype=F64)]
    _, n_iter, evals, _ = run(lambda q, **k: LBFGS(q, ops=TorchVectorOps(), **k), lambda xs: (xs[0] ** 2).sum(), p, 1)
    assert (n_iter, evals) == (0, 1)


def test_needs_cuda_without_injected_ops():
    with pytest.raises(RuntimeError, match="no CPU path"):
        LBFGS([torch.zeros(3, dtype=F64, requires_grad=True)])


@pytest.mark.parametrize("max_iter,max_eval", [(5, 6), (3, 3), (10, 12), (1, 1)])
def test_evaluation_budget_is_spent_like_torch(max_iter, max_eval):
    """``max_eval`` caps the line search (``max_ls = max_eval - current_evals``) and ends the step: same number of
    iterations and closure calls, same points, when the budget binds."""
    kw = dict(lr=1.0, max_iter=max_iter, max_eval=max_eval, tolerance_grad=1e-12, tolerance_change=1e-14,
              line_search_fn="strong_wolfe", history_size=4)
    _, pa = make_params(2, 4, 7, True)
    _, pb = make_params(2, 4, 7, True)
    la, ia, ea, xa = run(torch.optim.LBFGS, illcond, pa, 1, **kw)
    lb, ib, eb, xb = run(lambda p, **k: LBFGS(p, ops=TorchVectorOps(), **k), illcond, pb, 1, **kw)
    assert (ia, ea) == (ib, eb) and len(la) == len(lb)
    for a, b in zip(la, lb):
        assert abs(a - b) <= 1e-10 * max(1.0, abs(a))
    for a, b in zip(xa, xb):
        assert torch.allclose(a, b, rtol=1e-9, atol=1e-11)


def test_constructor_validation():
    p = [torch.zeros(3, dtype=F64, requires_grad=True)]
    with pytest.raises(RuntimeError, match="strong_wolfe"):
        LBFGS(p, ops=TorchVectorOps(), line_search_fn="armijo")
    with pytest.raises(ValueError, match="float64"):
        LBFGS([torch.zeros(3, dtype=torch.float32, requires_grad=True)], ops=TorchVectorOps())
    with pytest.raises(ValueError, match="learning rate"):
        LBFGS(p, ops=TorchVectorOps(), lr=-1.0)
