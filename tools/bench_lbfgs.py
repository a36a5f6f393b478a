"""HBM roofline of the L-BFGS vector primitives (csrc/lbfgs.cu) and the cost of one direction update against
torch.optim.LBFGS's two-loop recursion on the same history.
    python tools/bench_lbfgs.py [--n 55000000] [--history 10] [--reps 5]
One JSON line: per kernel the device time (CUDA events), algorithmic bytes (each operand once), achieved GB/s and the
fraction of MEASURED_PEAKS.json's HBM copy bandwidth."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=55_000_000)      # config #5 E-step vector on one of four GPUs
    ap.add_argument("--history", type=int, default=10)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    import torch
    from svgpfa_b200.lbfgs import CudaVectorOps
    dev = torch.device("cuda")
    peak = 6553.9
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    ops = CudaVectorOps(dev)
    n, h = args.n, args.history
    S = [torch.randn(n, dtype=torch.float64, device=dev) for _ in range(h)]
    Y = [torch.randn(n, dtype=torch.float64, device=dev) for _ in range(h)]
    g, gp, d, x0 = (torch.randn(n, dtype=torch.float64, device=dev) for _ in range(4))
    x = torch.empty(n, dtype=torch.float64, device=dev)

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(args.reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        return best
    vecs = S + Y + [g]
    nv = len(vecs)
    coefs = [0.01 * (i + 1) for i in range(nv)]
    rows = {}

    def row(name, ms, nvec, note):
        by = nvec * n * 8
        rows[name] = {"ms": ms, "algorithmic_bytes": by, "GBps": by / ms / 1e6, "frac_of_hbm_peak": by / ms / 1e6 / peak,
                      "what": note}
    row("multidot", timed(lambda: ops.multidot(vecs, [S[-1], Y[-1], g])), nv,
        f"{nv} stored vectors x 3 probes (the probes are stored vectors: re-read once per group of 8 from L2/HBM)")
    row("combine", timed(lambda: ops.combine(d, vecs, coefs, g)), nv + 1, f"d = sum of {nv} vectors, g.d, max|d| (reads {nv}, writes 1)")
    row("update", timed(lambda: ops.update(S[0], Y[0], d, 0.5, g, gp)), 6, "s = t d, y = g - gp, gp = g (reads 3, writes 3)")
    row("step", timed(lambda: ops.step(x, x0, d, 0.5)), 3, "x = x0 + t d (reads 2, writes 1)")
    row("stats", timed(lambda: ops.stats(g, d)), 2, "g.d, max|g|, sum|g|, max|d| (reads 2)")

    ro = [1.0 / (1.0 + i) for i in range(h)]

    def torch_two_loop():
        al = [None] * h
        q = g.neg()
        for i in range(h - 1, -1, -1):
            al[i] = S[i].dot(q) * ro[i]
            q.add_(Y[i], alpha=-al[i])
        r = torch.mul(q, 0.7)
        for i in range(h):
            be = Y[i].dot(r) * ro[i]
            r.add_(S[i], alpha=al[i] - be)
        return r
    t_torch = timed(torch_two_loop)
    out = {"n": n, "history_pairs": h, "hbm_peak_GBps": peak, "kernels": rows,
           "direction_update_ms": {"b200 (update + multidot + combine)": rows["update"]["ms"] + rows["multidot"]["ms"] + rows["combine"]["ms"],
                                   "torch two-loop recursion (y, s, dots and axpys as torch.optim.LBFGS issues them)": t_torch}}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
