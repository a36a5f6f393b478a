"""Times individual C-ABI stages (device time, CUDA events) on a shard of a BASELINE.json configuration.
    python tools/bench_kernels.py [--config config5] [--trials 2000] [--reps 5] [--quad-warps N] [--flags F]"""
import argparse
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--trials", type=int, default=2000)
    ap.add_argument("--config", default="config5")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--flags", type=int, default=15)
    ap.add_argument("--quad-warps", type=int, default=0)
    ap.add_argument("--spike-chunks", type=int, default=0)
    ap.add_argument("--keep", action="store_true",
                    help="no buildKernelsMatrices() between evaluations: the closure of an E-step (Kzz factors, spike-term "
                         "statistic and the V cache stay valid)")
    args = ap.parse_args()
    import torch
    from svgpfa_b200 import _cabi, synthetic
    from svgpfa_b200.testing import model_from_case, set_requires_grad
    dev = torch.device("cuda")
    cfg = dict(synthetic.CONFIGS[args.config], R=args.trials)
    case = synthetic.make_case_torch(cfg, dev, seed=0)
    model = model_from_case(case, device=dev, spike_chunks=args.spike_chunks)
    model._quad_warps = args.quad_warps
    set_requires_grad(model)
    lib = _cabi.lib()
    n_ev = len(_cabi.STAGES) + 1
    tot = [0.0] * len(_cabi.STAGES)
    v = model.eval()
    ref = v.item()
    for rep in range(args.reps + 1):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(n_ev)]
        for e in evs:
            e.record()
        arr = (ctypes.c_void_p * n_ev)(*[e.cuda_event for e in evs])
        lib.svgpfa_set_stage_events(arr)
        if not args.keep:
            model.buildKernelsMatrices()
        if args.flags == 15:
            v = model.eval()
        else:
            model._run(args.flags)
        lib.svgpfa_set_stage_events(None)
        torch.cuda.synchronize()
        if rep > 0:
            for j in range(len(_cabi.STAGES)):
                tot[j] += evs[j].elapsed_time(evs[j + 1]) / args.reps
    print(f"{args.config} R={args.trials} qw={args.quad_warps} flags={args.flags}: " +
          " ".join(f"{n}={t:.3f}" for n, t in zip(_cabi.STAGES, tot)) + f" total={sum(tot):.3f} elbo={ref:.10e}")


if __name__ == "__main__":
    main()
