import ctypes, torch, sys
sys.path.insert(0,'/root/repo')
from svgpfa_b200 import _cabi
lib=_cabi.lib(); dev=torch.device('cuda')
blocks=148*8; out=torch.empty(blocks*256,dtype=torch.float64,device=dev)
st=ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for kind,iters,per in ((0,20000,8),(4,20000,8*32)):
    best=1e9
    for _ in range(3):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); _cabi.check(lib.svgpfa_peak_probe(kind,blocks,iters,out.data_ptr(),st)); e1.record(); e1.synchronize()
        best=min(best,e0.elapsed_time(e1))
    print(kind, "FMA/s", blocks*256*iters*per/(best*1e-3)/1e12, "T  => TFLOP/s", 2*blocks*256*iters*per/(best*1e-3)/1e12)
