"""Full svEM (E-step + the three M-steps) on the CUDA model with the call sequence of the reference's SVEM_PyTorch
(tests/ecm_driver.py restates svEM.py:76-294 because /root/reference does not exist on the GPU box), torch.optim.LBFGS
with the reference's default settings shape, on BASELINE.json config #3 (R=2000, N=200, K=10, M=20, mixed kernels).
Prints one JSON line: seconds per EM iteration, closure evaluations per step, bound trajectory; optionally the same
driver on the oracle-backed CPU model for the first --cpu-trials trials (seconds per EM iteration, extrapolated
linearly in the number of trials).
    python tools/bench_svem.py [--config config3] [--em-iters 2] [--lbfgs-iters 10] [--cpu-trials 8]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="config3")
    ap.add_argument("--trials", type=int, default=None)
    ap.add_argument("--em-iters", type=int, default=2)
    ap.add_argument("--lbfgs-iters", type=int, default=10)
    ap.add_argument("--cpu-trials", type=int, default=8)
    ap.add_argument("--optimizer", default="torch", choices=("torch", "b200", "both"),
                    help="torch.optim.LBFGS, the device-resident svgpfa_b200.lbfgs.LBFGS, or one run with each")
    args = ap.parse_args()
    import torch
    import ecm_driver
    from svgpfa_b200 import synthetic
    from svgpfa_b200.testing import model_from_case
    dev = torch.device("cuda")
    cfg = dict(synthetic.CONFIGS[args.config])
    if args.trials:
        cfg["R"] = args.trials
    case = synthetic.make_case_torch(cfg, dev, seed=0)
    model = model_from_case(case, device=dev)
    kw = dict(max_iter=args.lbfgs_iters, lr=1.0, tolerance_grad=1e-7, tolerance_change=1e-9, line_search_fn="strong_wolfe")
    float(model.eval())                                   # warm-up: allocations, first launches
    # torch.optim.LBFGS pulls in torch._dynamo / sympy on its first step (~3 s of imports): not part of svEM
    x = torch.zeros(4, dtype=torch.float64, device=dev, requires_grad=True)
    opt = torch.optim.LBFGS([x], max_iter=2, line_search_fn="strong_wolfe")

    def _closure():
        opt.zero_grad()
        loss = ((x - 1.0) ** 2).sum()
        loss.backward()
        return loss
    opt.step(_closure)
    from svgpfa_b200 import ecm
    op = {"em_max_iter": args.em_iters}
    for s in ecm.STEP_ORDER["ecm"]:
        op[f"{s}_estimate"], op[f"{s}_optim_params"] = True, dict(kw)
    runs = {}
    for which in (("torch", "b200") if args.optimizer == "both" else (args.optimizer,)):
        if runs:                                          # same initial state for the second optimiser
            model = model_from_case(case, device=dev)
            float(model.eval())
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        hist, _, _, log = ecm.maximize(model, op, out=None, optimizer=which)
        torch.cuda.synchronize()
        runs[which] = (time.perf_counter() - t0, hist, log)
    first = next(iter(runs))
    dt, hist, log = runs[first]
    out = {"what": "full svEM (ECM: estep, mstep_embedding, mstep_kernels, mstep_indpointslocs) through the model protocol",
           "config": f"{args.config}: R={cfg['R']} N={cfg['N']} K={cfg['K']} M={cfg['M']} Q={cfg['Q']} mixed={cfg['mixed']}",
           "em_iters": args.em_iters, "lbfgs_max_iter": args.lbfgs_iters, "optimizer": first,
           "seconds_per_em_iter_gpu": dt / args.em_iters,
           "closure_evals": sum(l[4] for l in log), "bound": [hist[0], hist[-1]],
           "monotone": all(b >= a - 1e-9 * abs(a) for a, b in zip(hist, hist[1:])),
           "steps": [{"iter": l[0], "step": l[1], "bound": l[2], "lbfgs_iters": l[3], "evals": l[4]} for l in log]}
    if len(runs) > 1:
        dt2, hist2, log2 = runs["b200"]
        out["b200_optimizer"] = {"seconds_per_em_iter_gpu": dt2 / args.em_iters, "closure_evals": sum(l[4] for l in log2),
                                 "bound": [hist2[0], hist2[-1]],
                                 "same_counts_as_torch": [l[3:] for l in log] == [l[3:] for l in log2],
                                 "bound_rel_diff_per_step": [abs(a[2] - b[2]) / abs(a[2]) for a, b in zip(log, log2)],
                                 "steps": [{"iter": l[0], "step": l[1], "bound": l[2], "lbfgs_iters": l[3], "evals": l[4]}
                                           for l in log2]}
    if args.cpu_trials > 0:
        r_sub = min(args.cpu_trials, cfg["R"])
        sub = synthetic.case_to_numpy(case, 0, r_sub)
        cpu_model = ecm_driver.OracleModel(sub)
        torch.set_num_threads(os.cpu_count() or 1)
        t0 = time.perf_counter()
        hist_c, log_c = ecm_driver.maximize(cpu_model, em_max_iter=1, lbfgs_kwargs=kw)
        dtc = time.perf_counter() - t0
        out["cpu_oracle"] = {"trials": r_sub, "cores": os.cpu_count(), "seconds_per_em_iter_on_sample": dtc,
                             "closure_evals": sum(l[4] for l in log_c),
                             "seconds_per_em_iter_extrapolated": dtc * cfg["R"] / r_sub,
                             "note": "oracle port of the reference algorithm; cost linear in the number of trials"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
