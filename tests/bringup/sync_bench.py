import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import vlb200
from vlb200 import _native as nv
L = nv.lib()
L.vl_debug_sync_bench.restype = ctypes.c_int32
L.vl_debug_sync_bench.argtypes = [ctypes.c_int32] * 6 + [ctypes.c_void_p, ctypes.c_void_p]
out = torch.zeros(148, dtype=torch.int64, device="cuda")
iters = 4000
names = {0: "arrive", 1: "commit", 2: "arrive+fence", 3: "commit+fence", 5: "commit+mma", 7: "commit+fence+mma", 4: "arrive+mma"}
for bm, bn in ((128, 256), (128, 224), (64, 256), (64, 224), (64, 128), (64, 64)):
    for variant in (5,):
        for stages in (4,):
            for rep in range(2):
                nv.check(L.vl_debug_sync_bench(variant, stages, iters, bn, bm, 148, out.data_ptr(),
                                               ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
                torch.cuda.synchronize()
            print("bm %3d bn %3d %-18s stages %d: %7.1f clk/iter" % (bm, bn, names[variant], stages, out.float().mean().item() / iters), flush=True)

