"""Generates tests/golden/config1_example.npz: BASELINE.json config #1, the reference's own smoke test
(``/root/reference/examples/scripts/doEstimateSVGPFA.py:22-139`` with ``--em_max_iter=2``) run by the
UNMODIFIED reference in the build container:

    python tests/golden/make_config1.py

  data    /root/reference/examples/data/32451751_simRes.pickle   (R=15 trials, N=100 neurons, float32 spike
          tensors, 3 230 .. 27 344 spikes per trial, S = 197 662)
  params  /root/reference/examples/params/00000545_estimation_metaData.ini through the reference's own
          ``svGPFA.utils.initUtils`` (K=2 exponential-quadratic latents, M=9 equidistant inducing points,
          Q=200, reg=1e-3, LBFGS max_iter 20 / strong_wolfe for every ECM step)

The fixture holds the inputs exactly as the reference model received them, the reference's float64 outputs
at the initial point (``out_elbo`` = 277018.8745717274 -- SURVEY.md §6.2/§8c --, ELL, KL, every gradient,
the spike stacking) and the step log of the reference's ``SVEM_PyTorch.maximize`` (stats/svEM.py:76-216)
over two ECM iterations: one row ``[iteration, step index, lower bound, niter, nfeval]`` per step, steps in
the order estep, mstep_embedding, mstep_kernels, mstep_indpointslocs.

``gcnu_common`` is a third-party, un-vendored dependency of the reference (setup.cfg:21); the two helpers
the script calls (``config_dict.GetDict``, ``argparse.add_remaining_to_populated_args``) are INI/CLI
plumbing and are stood in for below; none of the lower-bound arithmetic lives there.
"""
from __future__ import annotations

import configparser
import io
import os
import pickle
import re
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, HERE)

from svgpfa_b200 import synthetic  # noqa: E402
import ref_harness  # noqa: E402

EXAMPLES = "/root/reference/examples"
STEP_NAMES = ("estep", "mstep_embedding", "mstep_kernels", "mstep_indpointslocs")


def strings_dict_from_ini(path):
    """What gcnu_common.utils.config_dict.GetDict(config).get_dict() returns: {section: {option: string}}."""
    cfg = configparser.ConfigParser()
    cfg.read(path)
    return {sec: dict(cfg[sec]) for sec in cfg.sections()}


def reference_params_and_data(em_max_iter=2):
    ref_harness.import_reference()
    import svGPFA.utils.initUtils as iu

    with open(os.path.join(EXAMPLES, "data", "32451751_simRes.pickle"), "rb") as f:
        sim = pickle.load(f)
    spikes = sim["spikes"]
    R, N, K = len(spikes), len(spikes[0]), 2
    args_info = iu.getArgsInfo()
    # doEstimateSVGPFA.py:66-69 with the script's argparse defaults plus --em_max_iter
    args = dict(sim_res_number=32451751, est_init_number=545, n_latents=K, trials_start_time=0.0,
                trials_end_time=1.0, em_max_iter=str(em_max_iter))
    dynamic = iu.getParamsDictFromArgs(n_latents=K, n_trials=R, args=args, args_info=args_info)
    cwd = os.getcwd()
    os.chdir(os.path.join(EXAMPLES, "scripts"))          # the INI names its CSV files relative to scripts/
    try:
        strings = strings_dict_from_ini("../params/00000545_estimation_metaData.ini")
        from_file = iu.getParamsDictFromStringsDict(n_latents=K, n_trials=R, strings_dict=strings,
                                                    args_info=args_info)
        params, kernels_types = iu.getParamsAndKernelsTypes(
            n_trials=R, n_neurons=N, n_latents=K, trials_start_times=[0.0] * R, trials_end_times=[1.0] * R,
            dynamic_params_spec=dynamic, config_file_params_spec=from_file)
    finally:
        os.chdir(cwd)
    return spikes, params, kernels_types


def case_from_reference_inputs(spikes, params, kernels_types):
    ip = params["initial_params"]
    pol = ip["posterior_on_latents"]
    K = len(kernels_types)
    kt = []
    for name in kernels_types:
        kt.append("periodic" if "eriodic" in name else "expquad")
    R, N = len(spikes), len(spikes[0])
    counts = np.array([[int(spikes[r][n].numel()) for n in range(N)] for r in range(R)], dtype=np.int64)
    times = torch.cat([spikes[r][n].reshape(-1) for r in range(R) for n in range(N)]).numpy()
    assert times.dtype == np.float32
    npy = lambda t: t.detach().numpy().astype(np.float64)
    return dict(
        kernel_types=kt,
        kernel_params=[npy(a) for a in pol["kernels_matrices_store"]["kernels_params0"]],
        Z=[npy(a) for a in pol["kernels_matrices_store"]["inducing_points_locs0"]],
        m=[npy(a) for a in pol["posterior_on_ind_points"]["mean"]],
        chol_vecs=[npy(a) for a in pol["posterior_on_ind_points"]["cholVecs"]],
        C=npy(ip["embedding"]["C0"]), d=npy(ip["embedding"]["d0"]),
        leg_quad_points=npy(params["ell_calculation_params"]["leg_quad_points"]),
        leg_quad_weights=npy(params["ell_calculation_params"]["leg_quad_weights"]),
        spike_times=times, spike_counts=counts,
        reg=float(params["optim_params"]["prior_cov_reg_param"]))


def run_reference_svem(spikes, params, kernels_types):
    """The script's model construction and maximisation (doEstimateSVGPFA.py:101-120), log captured."""
    import svGPFA.stats.svEM
    import svGPFA.stats.svGPFAModelFactory
    import svGPFA.utils.miscUtils
    k0 = params["initial_params"]["posterior_on_latents"]["kernels_matrices_store"]["kernels_params0"]
    kernels = svGPFA.utils.miscUtils.buildKernels(kernels_types=kernels_types, kernels_params=k0)
    model = svGPFA.stats.svGPFAModelFactory.SVGPFAModelFactory.buildModelPyTorch(kernels=kernels)
    model.setParamsAndData(measurements=spikes, initial_params=params["initial_params"],
                           eLLCalculationParams=params["ell_calculation_params"],
                           priorCovRegParam=params["optim_params"]["prior_cov_reg_param"])
    out = io.StringIO()
    svem = svGPFA.stats.svEM.SVEM_PyTorch()
    hist, _, term, _ = svem.maximize(model=model, optim_params=params["optim_params"],
                                     method=params["optim_params"]["optim_method"], out=out)
    rows = []
    pat = re.compile(r"Iteration (\d+), (\w+) end: ([-\d.eE+naif]+), niter: (\d+), nfeval: (\d+)")
    for line in out.getvalue().splitlines():
        mt = pat.match(line)
        if mt:
            rows.append([int(mt.group(1)), STEP_NAMES.index(mt.group(2)), float(mt.group(3)),
                         int(mt.group(4)), int(mt.group(5))])
    return np.array(hist), np.array(rows), out.getvalue(), model


def main():
    torch.set_num_threads(8)
    spikes, params, kernels_types = reference_params_and_data(em_max_iter=2)
    case = case_from_reference_inputs(spikes, params, kernels_types)
    # outputs at the initial point, through the same harness as every other fixture (it rebuilds the reference
    # model from the case dict -- float32 spike tensors included -- so the fixture is self-consistent)
    out = ref_harness.reference_outputs(case, with_stats=False)
    hist, rows, text, model = run_reference_svem(spikes, params, kernels_types)
    assert abs(hist[0] - out["elbo"]) <= 1e-12 * abs(out["elbo"]), (hist[0], out["elbo"])
    out["svem_lower_bound_hist"] = hist
    out["svem_step_log"] = rows
    # the fitted shared parameters after the two iterations (the bound in the log is printed with 6 decimals only)
    out["svem_final_C"] = model.getSVEmbeddingParams()[0].detach().numpy()
    out["svem_final_d"] = model.getSVEmbeddingParams()[1].detach().numpy()
    out["svem_final_kernel_params"] = np.concatenate([p.detach().numpy().reshape(-1) for p in model.getKernelsParams()])
    path = os.path.join(HERE, "config1_example.npz")
    synthetic.save_case(path, case, extra=out)
    print(text)
    print(f"config1_example: S={case['spike_times'].size} elbo={out['elbo']!r} hist={hist.tolist()!r} "
          f"-> {os.path.getsize(path)/1e3:.0f} kB")


if __name__ == "__main__":
    main()
