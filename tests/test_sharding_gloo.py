"""Multi-rank host logic on CPU: trial partitioning and the one all-reduce of the packed
[elbo | dC | dd | dtheta] buffer, world_size 2 over gloo.  Each rank evaluates its trial shard with the
oracle (test infrastructure standing in for the CUDA kernels, which need a GPU) and the reduced buffer must
equal the single-process evaluation; per-trial gradients stay local."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN, rel_err
from svgpfa_b200 import sharding, synthetic


def test_trial_blocks_cover_and_balance():
    rng = np.random.default_rng(0)
    s = rng.lognormal(8.0, 0.8, size=501)                 # heavy ragged, like config #4
    for world in (1, 2, 3, 4, 8):
        blocks = sharding.trial_blocks(s, world)
        assert blocks[0][0] == 0 and blocks[-1][1] == s.size
        assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
        loads = np.array([s[a:b].sum() for a, b in blocks])
        assert loads.max() <= (s.sum() / world) + s.max()          # within one trial of perfect balance
    assert sharding.trial_blocks(np.zeros(10), 4) == [(0, 2), (2, 5), (5, 7), (7, 10)]
    assert sharding.trial_blocks([], 2) == [(0, 0), (0, 0)]
    even = sharding.trial_blocks(np.ones(20000), 8)
    assert [b - a for a, b in even] == [2500] * 8


def test_shared_layout_matches_cabi():
    from svgpfa_b200 import _cabi
    lay = sharding.shared_layout(7, 3, 4)
    assert _cabi.SHARED_HDR == sharding.SHARED_HDR == 8
    assert lay["C"] == (8, 29) and lay["d"] == (29, 36) and lay["theta"] == (36, 40) and lay["length"] == 40


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, path, out_q):
    from oracle import svgpfa_oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    case, _ = synthetic.load_case(path)
    per_trial = case["spike_counts"].sum(axis=1)
    r0, r1 = sharding.trial_blocks(per_trial, world)[rank]
    shard = synthetic.slice_trials(case, r0, r1)
    out = orc.elbo_and_grads(shard)
    K = len(case["kernel_types"])
    dth = np.concatenate([out[f"grad_kernel_params_{k}"].reshape(-1) for k in range(K)])
    shared = torch.from_numpy(sharding.pack_shared(out["elbo"], out["ell"], out["kl"], out["grad_C"],
                                                   out["grad_d"], dth))
    sharding.all_reduce_shared(shared, dist.group.WORLD)
    local = {k: out[k] for k in out if k.startswith(("grad_m_", "grad_chol_vecs_", "grad_Z_"))}
    out_q.put((rank, (r0, r1), shared.numpy().copy(), local))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["tiny_mixed", "config2_r8"])
def test_two_rank_allreduce_equals_single_process(name):
    path = os.path.join(GOLDEN, name + ".npz")
    case, ref = synthetic.load_case(path)
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, path, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    K = len(case["kernel_types"])
    N = case["C"].shape[0]
    TH = sum(len(t) for t in case["kernel_params"])
    lay = sharding.shared_layout(N, K, TH)
    bufs = [r[2] for r in results]
    assert np.array_equal(bufs[0], bufs[1])                # every rank holds bit-identical sums
    buf = bufs[0]
    assert abs(buf[lay["elbo"]] - float(ref["elbo"])) <= 1e-12 * abs(float(ref["elbo"]))
    assert rel_err(buf[lay["C"][0]:lay["C"][1]], ref["grad_C"]) <= 1e-11
    assert rel_err(buf[lay["d"][0]:lay["d"][1]], ref["grad_d"]) <= 1e-11
    dth = np.concatenate([ref[f"grad_kernel_params_{k}"].reshape(-1) for k in range(K)])
    assert rel_err(buf[lay["theta"][0]:lay["theta"][1]], dth) <= 1e-11
    for k in range(K):
        for key in (f"grad_m_{k}", f"grad_chol_vecs_{k}", f"grad_Z_{k}"):
            stacked = np.concatenate([r[3][key] for r in results], axis=0)
            assert rel_err(stacked, ref[key]) <= 1e-11, key
