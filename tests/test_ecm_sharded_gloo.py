"""The shard-aware ECM driver (svgpfa_b200/ecm.py) on CPU: world size 1 against tests/ecm_driver.py (which is pinned
to the reference's SVEM_PyTorch), and world size 2 over gloo with each rank holding a block of trials (the oracle
standing in for the CUDA kernels; the multi-rank decisions are the product's code, see tests/sharded_oracle_model.py).
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN
from svgpfa_b200 import ecm, sharding, synthetic

KW = dict(max_iter=6, lr=1.0, tolerance_grad=1e-7, tolerance_change=1e-9, line_search_fn="strong_wolfe")


def optim_params(em_max_iter, estimate=("estep", "mstep_embedding", "mstep_kernels", "mstep_indpointslocs")):
    p = {"em_max_iter": em_max_iter}
    for s in ecm.STEP_ORDER["ecm"]:
        p[f"{s}_estimate"] = s in estimate
        p[f"{s}_optim_params"] = dict(KW)
    return p


def test_single_process_driver_equals_the_pinned_test_driver():
    import ecm_driver
    from sharded_oracle_model import ShardedOracleModel
    torch.set_num_threads(1)
    case, _ = synthetic.load_case(os.path.join(GOLDEN, "tiny_mixed.npz"))
    hist_a, log_a = ecm_driver.maximize(ecm_driver.OracleModel(case), em_max_iter=2, lbfgs_kwargs=KW)
    hist_b, _, msg, log_b = ecm.maximize(ShardedOracleModel(case), optim_params(2), out=None)
    assert "Maximum number of iterations" in msg
    assert [r[:2] + r[3:] for r in log_a] == [r[:2] + r[3:] for r in log_b]          # steps, niter, nfeval
    assert [r[2] for r in log_b] == pytest.approx([r[2] for r in log_a], rel=1e-9)
    assert hist_b == pytest.approx(hist_a, rel=1e-9)
    with pytest.raises(ValueError):
        ecm.maximize(ShardedOracleModel(case), optim_params(1), method="em", out=None)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, path, estimate, q):
    from sharded_oracle_model import ShardedOracleModel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    case, _ = synthetic.load_case(path)
    r0, r1 = sharding.trial_blocks(case["spike_counts"].sum(axis=1), world)[rank]
    model = ShardedOracleModel(synthetic.slice_trials(case, r0, r1), pg=dist.group.WORLD)
    hist, _, msg, log = ecm.maximize(model, optim_params(2, estimate), out=None)
    C, d = model.getSVEmbeddingParams()
    th = np.concatenate([p.detach().numpy().reshape(-1) for p in model.getKernelsParams()])
    q.put((rank, hist, log, msg, C.detach().numpy().copy(), th, model.n_evals, model.n_reduced))
    dist.barrier()
    dist.destroy_process_group()


def _run_world2(path, estimate):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, path, estimate, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted([q.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return results


def test_two_ranks_shared_parameter_steps_follow_the_single_process_trajectory():
    """Embedding and kernels M-steps only: every closure evaluation is all-reduced, both ranks take identical
    optimiser decisions, and the step log equals the single-process one (same niter / nfeval, bounds to 1e-9)."""
    from sharded_oracle_model import ShardedOracleModel
    path = os.path.join(GOLDEN, "tiny_mixed.npz")
    est = ("mstep_embedding", "mstep_kernels")
    res = _run_world2(path, est)
    case, _ = synthetic.load_case(path)
    torch.set_num_threads(1)
    hist1, _, _, log1 = ecm.maximize(ShardedOracleModel(case), optim_params(2, est), out=None)
    for rank, hist, log, msg, C, th, n_evals, n_reduced in res:
        assert "Maximum number of iterations" in msg
        assert [r[:2] + r[3:] for r in log] == [r[:2] + r[3:] for r in log1]
        assert [r[2] for r in log] == pytest.approx([r[2] for r in log1], rel=1e-9)
        assert hist == pytest.approx(hist1, rel=1e-9)
        assert n_evals == n_reduced                                  # lock-step: every evaluation was a collective
    assert res[0][1] == res[1][1] and np.array_equal(res[0][4], res[1][4]) and np.array_equal(res[0][5], res[1][5])


def test_two_ranks_full_ecm_terminates_and_is_monotone():
    """All four steps: the per-trial steps run rank-locally (no collective in their closures: the ranks may take
    different numbers of evaluations), every step ends on an agreed bound, the bound never decreases between steps of
    the same kind, and the ranks end with identical shared parameters."""
    path = os.path.join(GOLDEN, "tiny_mixed.npz")
    res = _run_world2(path, ecm.STEP_ORDER["ecm"])
    (_, hist0, log0, msg0, C0, th0, ne0, nr0), (_, hist1, log1, msg1, C1, th1, ne1, nr1) = res
    assert "Maximum number of iterations" in msg0 and "Maximum number of iterations" in msg1
    assert hist0 == hist1 and len(log0) == len(log1) == 8
    assert [r[:3] for r in log0] == [r[:3] for r in log1]            # same agreed bound after every step
    assert np.array_equal(C0, C1) and np.array_equal(th0, th1)
    assert nr0 == nr1 and ne0 > nr0 and ne1 > nr1                    # same number of collectives, local evaluations besides
    elbo_steps = [r[2] for r in log0 if r[1] != "mstep_embedding"]   # that step logs the ELL without the KL term
    assert hist0[0] <= elbo_steps[0] + 1e-9 * abs(hist0[0])
    assert all(b >= a - 1e-9 * abs(a) for a, b in zip(elbo_steps, elbo_steps[1:]))
    # the two-rank fit reaches (at least) the bound of the single-process joint optimisation after the same work
    import ecm_driver
    case, _ = synthetic.load_case(path)
    torch.set_num_threads(1)
    hist_single, _ = ecm_driver.maximize(ecm_driver.OracleModel(case), em_max_iter=2, lbfgs_kwargs=KW)
    assert hist0[-1] >= hist_single[-1] - 0.02 * abs(hist_single[-1])


def _failing_worker(rank, world, port, path, q):
    from sharded_oracle_model import ShardedOracleModel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    case, _ = synthetic.load_case(path)
    r0, r1 = sharding.trial_blocks(case["spike_counts"].sum(axis=1), world)[rank]

    class Failing(ShardedOracleModel):
        def eval(self):
            # a Kzz of THIS rank's trials stops being positive definite while its inducing points move
            if rank == 1 and any(z.requires_grad for z in self.getIndPointsLocs()) and self.n_evals > 40:
                raise torch.linalg.LinAlgError("linalg.cholesky: Kzz is not positive-definite")
            return super().eval()

    model = Failing(synthetic.slice_trials(case, r0, r1), pg=dist.group.WORLD)
    hist, _, msg, log = ecm.maximize(model, optim_params(2), out=None)
    q.put((rank, msg, len(log)))
    dist.barrier()
    dist.destroy_process_group()


def test_failure_on_one_rank_stops_every_rank_at_the_same_step():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    path = os.path.join(GOLDEN, "tiny_mixed.npz")
    procs = [ctx.Process(target=_failing_worker, args=(r, world, port, path, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted([q.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, msg0, n0), (_, msg1, n1) = results
    assert "mstep_indpointslocs" in msg0 and "mstep_indpointslocs" in msg1 and "failed on 1 rank" in msg0
    assert n0 == n1
