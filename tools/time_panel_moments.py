import ctypes, sys, torch
sys.path.insert(0, "/root/repo")
from svgpfa_b200 import _cabi, synthetic
from svgpfa_b200.testing import model_from_case
dev = torch.device("cuda")
cfg = dict(synthetic.CONFIGS["config5"], R=2000)
model = model_from_case(synthetic.make_case_torch(cfg, dev, seed=0), device=dev)
model.eval()
lib = _cabi.lib(); st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
best = 1e9
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); _cabi.check(lib.svgpfa_panel_moments(ctypes.byref(model._dims), ctypes.byref(model._bufs), st)); e1.record(); e1.synchronize()
    best = min(best, e0.elapsed_time(e1))
print("panel_moments ms (2000 trials):", best)
