"""Generates the committed golden fixtures under tests/golden/*.npz by running the
UNMODIFIED reference (imported from /root/reference/src) in the build container.

    python tests/golden/make_golden.py

Each fixture holds the inputs of one problem instance plus the reference's own float64
outputs on it (``out_*`` keys): ELBO, ELL, KL, every gradient, the spike stacking of
``PointProcessELL.__stackSpikeTimes`` and the latent / embedding statistics that the
reference's unit tests pin.  ``matlab_r5`` is the MATLAB golden problem of
``/root/reference/src/svGPFA/stats/tests/data/Estep_Objective_PointProcess_svGPFA.mat``
(read the way ``stats/tests/test_svLowerBound.py:18-106`` reads it) together with
MATLAB's own ``Elik``/``KLd``/``obj`` values.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, HERE)

from svgpfa_b200 import synthetic  # noqa: E402
import ref_harness  # noqa: E402

DATA = "/root/reference/src/svGPFA/stats/tests/data"


def matlab_case():
    from scipy.io import loadmat
    ref_harness.import_reference()
    import svGPFA.utils.miscUtils as mu

    mat = loadmat(os.path.join(DATA, "Estep_Objective_PointProcess_svGPFA.mat"))
    K = len(mat["Z"])
    R = mat["Z"][0, 0].shape[2]
    perm = lambda a: torch.from_numpy(a).double().permute(2, 0, 1)
    q_mu = [perm(mat["q_mu"][0, k]) for k in range(K)]
    q_sqrt = [perm(mat["q_sqrt"][0, k]) for k in range(K)]
    q_diag = [perm(mat["q_diag"][0, k]) for k in range(K)]
    chol_vecs = mu.getSRQSigmaVec(qSVec=q_sqrt, qSDiag=q_diag)
    Z = [perm(mat["Z"][k, 0]) for k in range(K)]
    kernel_types, kernel_params = [], []
    for k in range(K):
        name = str(mat["kernelNames"][0, k][0])
        hp = np.asarray(mat["hprs"][k, 0], dtype=np.float64).reshape(-1)
        if name == "PeriodicKernel":
            kernel_types.append("periodic")
            kernel_params.append(hp[:2].copy())
        elif name == "rbfKernel":
            kernel_types.append("expquad")
            kernel_params.append(hp[:1].copy())
        else:
            raise ValueError(name)
    y = loadmat(os.path.join(DATA, "YNonStacked.mat"))["YNonStacked"]
    N = y[0, 0].shape[0]
    counts = np.zeros((R, N), dtype=np.int64)
    times = []
    for r in range(R):
        for n in range(N):
            s = np.asarray(y[r, 0][n, 0], dtype=np.float64).reshape(-1)
            counts[r, n] = s.size
            times.append(s)
    case = dict(
        kernel_types=kernel_types, kernel_params=kernel_params,
        Z=[z.numpy() for z in Z], m=[a.numpy() for a in q_mu],
        chol_vecs=[a.numpy() for a in chol_vecs],
        C=np.asarray(mat["C"], dtype=np.float64), d=np.asarray(mat["b"], dtype=np.float64),
        leg_quad_points=np.ascontiguousarray(np.transpose(mat["ttQuad"], (2, 0, 1))).astype(np.float64),
        leg_quad_weights=np.ascontiguousarray(np.transpose(mat["wwQuad"], (2, 0, 1))).astype(np.float64),
        spike_times=np.concatenate(times), spike_counts=counts, reg=1e-5)
    matlab = dict(matlab_Elik=float(mat["Elik"][0, 0]), matlab_KLd=float(mat["KLd"][0, 0]),
                  matlab_obj=float(mat["obj"][0, 0]))
    return case, matlab


def empty_segments_case():
    case = synthetic.make_case("tiny", seed=3)
    counts = case["spike_counts"].copy()
    pieces = synthetic.nested_spikes(case)
    counts[1, :] = 0            # a trial without any spike
    counts[:, 2] = 0            # a neuron that never fires
    counts[0, 0] = 0
    times = [pieces[r][n][:counts[r, n]] for r in range(counts.shape[0]) for n in range(counts.shape[1])]
    case["spike_counts"] = counts
    case["spike_times"] = np.concatenate(times) if times else np.zeros(0)
    return case


def main():
    torch.set_num_threads(8)
    cases = {}
    case, matlab = matlab_case()
    cases["matlab_r5"] = (case, matlab, True)
    cases["tiny_mixed"] = (synthetic.make_case("tiny", seed=1, M_list=[5, 4, 6], d_2d=False), {}, True)
    cases["tiny_f32"] = (synthetic.make_case("tiny", seed=2, spike_dtype=np.float32), {}, True)
    cases["tiny_empty"] = (empty_segments_case(), {}, True)
    cases["tiny_reg1e-5"] = (synthetic.make_case("tiny", seed=4, reg=1e-5), {}, False)
    cases["config2_r8"] = (synthetic.make_case("config2", seed=0, R=8), {}, False)
    cfg3 = dict(synthetic.CONFIGS["config3"], R=4, N=40)
    cases["config3_r4"] = (synthetic.make_case(cfg3, seed=0), {}, False)
    cfg4 = dict(synthetic.CONFIGS["config4"], R=3, N=30)
    cases["config4_r3"] = (synthetic.make_case(cfg4, seed=0), {}, False)
    for name, (case, extra, with_stats) in cases.items():
        out = ref_harness.reference_outputs(case, with_stats=with_stats)
        out.update(extra)
        path = os.path.join(HERE, name + ".npz")
        synthetic.save_case(path, case, extra=out)
        print(f"{name}: S={case['spike_times'].size} elbo={out['elbo']!r} ell={out['ell']!r} "
              f"kl={out['kl']!r} -> {os.path.getsize(path)/1e3:.0f} kB")


if __name__ == "__main__":
    main()
