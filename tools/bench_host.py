"""Host-buffer entry (svgpfa_elbo_grad_host) timing for several pipeline depths.
    python tools/bench_host.py --trials 2500 --blocks 1,2,4,8"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--trials", type=int, default=2500)
    ap.add_argument("--config", default="config5")
    ap.add_argument("--blocks", default="1,2,4,8")
    ap.add_argument("--reps", type=int, default=4)
    args = ap.parse_args()
    import torch
    from svgpfa_b200 import synthetic
    from svgpfa_b200.testing import model_from_case
    dev = torch.device("cuda")
    cfg = dict(synthetic.CONFIGS[args.config], R=args.trials)
    case = synthetic.make_case_torch(cfg, dev, seed=0)
    model = model_from_case(case, device=dev)
    io = model.makeHostIO(pin=True)
    for nb in [int(b) for b in args.blocks.split(",")]:
        for static in (True, False):
            model.evalAndGradHost(io, copy_static=static, n_blocks=nb)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.reps):
                elbo, h2d, d2h = model.evalAndGradHost(io, copy_static=static, n_blocks=nb)
            e1.record()
            e1.synchronize()
            print(f"R={args.trials} n_blocks={nb} copy_static={static}: {e0.elapsed_time(e1) / args.reps:.2f} ms "
                  f"(h2d {h2d / 1e6:.0f} MB, d2h {d2h / 1e6:.0f} MB) elbo={elbo:.10e}")


if __name__ == "__main__":
    main()
