"""Secondary figures: milliseconds per LBFGS closure of svEM's four steps through the model API
(Python overhead and the per-evaluation host sync included), wall clock after synchronisation.
    python tools/bench_steps.py --config config2 [--trials R]"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="config2")
    ap.add_argument("--trials", type=int, default=None)
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    import torch
    from svgpfa_b200 import synthetic
    from svgpfa_b200.testing import model_from_case, set_requires_grad
    dev = torch.device("cuda")
    cfg = dict(synthetic.CONFIGS[args.config])
    if args.trials:
        cfg["R"] = args.trials
    case = synthetic.make_case_torch(cfg, dev, seed=0)
    model = model_from_case(case, device=dev)
    leaves = model._leaves()

    def timeit(fn, reps):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps * 1e3

    def closure(build):
        def f():
            for p in leaves:
                p.grad = None
            if build:
                model.buildKernelsMatrices()
            v = -model.eval()
            v.backward()
            return float(v)
        return f

    out = {}
    set_requires_grad(model)
    out["all_params"] = timeit(closure(True), args.reps)
    set_requires_grad(model, posterior=True, embedding=False, kernels=False, indlocs=False)
    out["estep (caches on)"] = timeit(closure(False), args.reps)
    set_requires_grad(model, posterior=False, embedding=False, kernels=True, indlocs=False)
    out["mstep_kernels"] = timeit(closure(True), args.reps)
    set_requires_grad(model, posterior=False, embedding=False, kernels=False, indlocs=True)
    out["mstep_indpointslocs"] = timeit(closure(True), args.reps)
    set_requires_grad(model, posterior=False, embedding=True, kernels=False, indlocs=False)
    t0 = time.perf_counter()
    stats = model.computeSVPosteriorOnLatentsStats()
    torch.cuda.synchronize()
    out["compute_stats (once per step)"] = (time.perf_counter() - t0) * 1e3

    def emb():
        for p in leaves:
            p.grad = None
        v = -model.evalELLSumAcrossTrialsAndNeurons(svPosteriorOnLatentsStats=stats)
        v.backward()
        return float(v)
    out["mstep_embedding (cached stats)"] = timeit(emb, args.reps)
    with torch.no_grad():
        out["forward only"] = timeit(lambda: float(model.eval()), args.reps)
    print(f"{args.config} R={cfg['R']} N={cfg['N']} K={cfg['K']} M={cfg['M']}: " +
          ", ".join(f"{k}={v:.3f} ms" for k, v in out.items()))


if __name__ == "__main__":
    main()
