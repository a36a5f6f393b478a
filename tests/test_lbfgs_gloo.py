"""``svgpfa_b200.lbfgs.LBFGS`` inside the ECM driver on CPU (the oracle standing in for the CUDA kernels, the torch
implementation of the vector primitives injected): against ``torch.optim.LBFGS`` in one process, and over gloo at world
size 2 in JOINT mode -- the per-trial steps as one optimisation over both ranks' vectors with global reductions, which
must reproduce the single-process trajectory (SURVEY.md §8e option (ii), §8f-3)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(__file__))
from conftest import GOLDEN  # noqa: E402
from svgpfa_b200 import ecm, sharding, synthetic  # noqa: E402
from test_ecm_sharded_gloo import _free_port, optim_params  # noqa: E402


def b200_lbfgs_on_torch_ops(params, **kw):
    from vector_ops_torch import TorchVectorOps
    from svgpfa_b200.lbfgs import LBFGS
    return LBFGS(params, ops=TorchVectorOps(), **kw)


def test_ecm_with_the_b200_optimiser_follows_torch_lbfgs():
    from sharded_oracle_model import ShardedOracleModel
    torch.set_num_threads(1)
    case, _ = synthetic.load_case(os.path.join(GOLDEN, "tiny_mixed.npz"))
    hist_a, _, _, log_a = ecm.maximize(ShardedOracleModel(case), optim_params(2), out=None)
    hist_b, _, msg, log_b = ecm.maximize(ShardedOracleModel(case), optim_params(2), out=None,
                                         optimizer=b200_lbfgs_on_torch_ops)
    assert "Maximum number of iterations" in msg
    assert [r[:2] + r[3:] for r in log_a] == [r[:2] + r[3:] for r in log_b]          # steps, niter, nfeval
    assert [r[2] for r in log_b] == pytest.approx([r[2] for r in log_a], rel=1e-9)
    assert hist_b == pytest.approx(hist_a, rel=1e-9)


def test_joint_mode_is_refused_with_the_torch_optimiser():
    from sharded_oracle_model import ShardedOracleModel
    case, _ = synthetic.load_case(os.path.join(GOLDEN, "tiny_mixed.npz"))
    with pytest.raises(ValueError, match="joint"):
        ecm.run_step(ShardedOracleModel(case), "estep", {}, process_group=object(), optimizer="torch",
                     sharded_steps="joint")


def _worker(rank, world, port, path, q):
    from sharded_oracle_model import ShardedOracleModel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    case, _ = synthetic.load_case(path)
    r0, r1 = sharding.trial_blocks(case["spike_counts"].sum(axis=1), world)[rank]
    model = ShardedOracleModel(synthetic.slice_trials(case, r0, r1), pg=dist.group.WORLD)
    hist, _, msg, log = ecm.maximize(model, optim_params(2), out=None, optimizer=b200_lbfgs_on_torch_ops,
                                     sharded_steps="joint")
    m0 = model.getSVPosteriorOnIndPointsParams()[0].detach().numpy().copy()
    q.put((rank, hist, log, msg, (r0, r1), m0, model.n_evals, model.n_reduced, model.shard_mode))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_joint_mode_follows_the_single_process_trajectory():
    from sharded_oracle_model import ShardedOracleModel
    path = os.path.join(GOLDEN, "tiny_mixed.npz")
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, path, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.set_num_threads(1)
    case, _ = synthetic.load_case(path)
    single = ShardedOracleModel(case)
    hist1, _, _, log1 = ecm.maximize(single, optim_params(2), out=None, optimizer=b200_lbfgs_on_torch_ops)
    m_single = single.getSVPosteriorOnIndPointsParams()[0].detach().numpy()
    for rank, hist, log, msg, (r0, r1), m0, n_evals, n_reduced, mode in res:
        assert "Maximum number of iterations" in msg and mode == "auto"
        assert [r[:2] + r[3:] for r in log] == [r[:2] + r[3:] for r in log1]      # all 8 steps: same niter / nfeval
        assert [r[2] for r in log] == pytest.approx([r[2] for r in log1], rel=1e-9)
        assert hist == pytest.approx(hist1, rel=1e-9)
        assert n_evals == n_reduced                                  # lock-step throughout: every evaluation a collective
        np.testing.assert_allclose(m0, m_single[r0:r1], rtol=1e-6, atol=1e-8)      # each rank holds its block of the joint optimum
    assert res[0][1] == res[1][1] and res[0][2] == res[1][2]
