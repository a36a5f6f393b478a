"""The spike kernel's evaluation sequence in isolation (svgpfa_peak_probe kinds 20-23): cycles per warp evaluation per
SM sub-partition at 1.965 GHz, for several resident-CTA counts (128-thread CTAs, 40 KB static shared memory each)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svgpfa_b200 import _cabi

lib = _cabi.probes()
dev = torch.device("cuda")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
names = {20: "full", 21: "no moments", 22: "no table LDS", 23: "no spike-time LDS", 24: "degree-3 polynomial",
         25: "I2F range reduction", 26: "degree 3 + I2F", 27: "degree 3 + I2F + pre-scaled t"}
for per_sm in (5, 3, 2, 1):
    blocks = 148 * per_sm
    out = torch.zeros(blocks * 128, dtype=torch.float64, device=dev)
    for kind in (20, 21, 22, 23, 24, 25, 26, 27):
        iters = 40000
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _cabi.check_probe(lib.svgpfa_peak_probe(kind, blocks, iters, out.data_ptr(), st))
            e1.record()
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        warp_evals_per_smsp = per_sm * 4 * iters * 4 / 4          # 4 warps per CTA, 4 evaluations per step, 4 SMSPs
        print(f"{per_sm} CTAs/SM  {names[kind]:18s} {best:8.3f} ms  {best * 1e-3 * 1.965e9 / warp_evals_per_smsp:6.2f} cycles per warp evaluation per SMSP")
