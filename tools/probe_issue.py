"""FP64 issue-model probes (svgpfa_peak_probe kinds 10-16): thread-instructions per second of the FP64 operation and the
implied cycles per warp instruction per SM sub-partition at the sampled SM clock."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svgpfa_b200 import _cabi

lib = _cabi.probes()
dev = torch.device("cuda")
blocks = 148 * 8
out = torch.zeros(blocks * 256, dtype=torch.float64, device=dev)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
names = {0: "DFMA const operands", 10: "DFMA 3 regs", 11: "DADD 2 regs", 12: "DMUL 2 regs", 13: "DFMA + 1 int",
         14: "DFMA + 2 int", 15: "DFMA + 3 int", 16: "2 DFMA + 1 int + LDS(+addr)", 3: "svgpfa_exp_neg (+2 fp64)"}
for kind in (0, 10, 11, 12, 13, 14, 15, 16, 3):
    iters = 20000
    best = 1e9
    for _ in range(3):
        out.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _cabi.check_probe(lib.svgpfa_peak_probe(kind, blocks, iters, out.data_ptr(), st))
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    steps = blocks * 256 * iters * 8
    warp_steps_per_smsp = steps / 32 / (148 * 4)
    cyc = best * 1e-3 * 1.965e9 / warp_steps_per_smsp
    print(f"kind {kind:2d} {names[kind]:32s} {steps / (best * 1e-3) / 1e12:7.3f} T steps/s   {cyc:6.2f} cycles per warp step per SMSP")
