"""HBM roofline of the cached-statistics spike gather (north_star (iv); SURVEY.md 8d: bytes S (8K + 4) + R Q K 16):
device time of svgpfa_cached_ell_fwd_bwd with and without the gather (SVGPFA_REUSE_SPIKE), CUDA events.
    python tools/bench_gather.py [--config config5] [--trials R]"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="config5")
    ap.add_argument("--trials", type=int, default=None)
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    import torch
    from svgpfa_b200 import _cabi, synthetic
    from svgpfa_b200.testing import model_from_case
    dev = torch.device("cuda")
    cfg = dict(synthetic.CONFIGS[args.config])
    if args.trials:
        cfg["R"] = args.trials
    model = model_from_case(synthetic.make_case_torch(cfg, dev, seed=0), device=dev, spike_method="direct")
    stats = model.computeSVPosteriorOnLatentsStats()
    model.evalELLSumAcrossTrialsAndNeurons(svPosteriorOnLatentsStats=stats)       # builds the buffers
    lib = _cabi.lib()
    b = _cabi.Buffers.from_buffer_copy(model._bufs)
    shared = torch.empty(model._shared_len, dtype=torch.float64, device=dev)
    mu_q, var_q, mu_s = model._cached_keepalive
    b.shared, b.mu_q, b.var_q, b.mu_s = shared.data_ptr(), mu_q.data_ptr(), var_q.data_ptr(), mu_s.data_ptr()
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def timed(flags):
        best = 1e30
        for _ in range(args.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _cabi.check(lib.svgpfa_cached_ell_fwd_bwd(ctypes.byref(model._dims), ctypes.byref(b), flags, st))
            e1.record()
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best
    t_full, t_reuse = timed(0), timed(_cabi.REUSE_SPIKE)
    S, K, R, N = model._S, model._K, model._R, model._N
    bytes_gather = S * K * 8 + (R * N + 1) * 8
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6553.9
    gbs = bytes_gather / ((t_full - t_reuse) * 1e-3) / 1e9
    print(json.dumps({"config": args.config, "R": R, "N": N, "K": K, "spikes": S, "cached_ell_ms": t_full,
                      "cached_ell_reuse_ms": t_reuse, "gather_ms": t_full - t_reuse, "gather_bytes": bytes_gather,
                      "gather_GBps": gbs, "hbm_peak_GBps": peak, "frac_of_hbm_peak": gbs / peak}))


if __name__ == "__main__":
    main()
