"""FP64 pipe probes: DFMA, mma.m8n8k4.f64, and mixes of the two (are the pipes independent?)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svgpfa_b200 import _cabi
lib = _cabi.probes(); dev = torch.device('cuda')
blocks = 148 * 8; out = torch.empty(blocks * 256, dtype=torch.float64, device=dev)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
iters = 20000
for kind, name in ((0, "8 DFMA"), (4, "8 DMMA"), (7, "mix 8 DFMA + 0 DMMA"), (8, "mix 0 DFMA + 2 DMMA"),
                   (5, "mix 8 DFMA + 1 DMMA"), (6, "mix 8 DFMA + 2 DMMA"), (9, "mix 8 DFMA + 4 DMMA")):
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); _cabi.check_probe(lib.svgpfa_peak_probe(kind, blocks, iters, out.data_ptr(), st)); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    cyc = best * 1e-3 * 1.965e9 / iters / (blocks * 8 / (148 * 4))      # SMSP cycles per loop iteration per warp
    print(f"{name:24s} {best:8.3f} ms  ~{cyc:6.1f} cycles per warp-iteration")
