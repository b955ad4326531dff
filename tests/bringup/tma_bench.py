import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import vlb200
from vlb200 import _native as nv
L = nv.lib()
L.vl_debug_tma_bench.restype = ctypes.c_int32
L.vl_debug_tma_bench.argtypes = [ctypes.c_void_p] + [ctypes.c_int32] * 9 + [ctypes.c_void_p, ctypes.c_void_p]
n_img, h, w, c = 1024, 13, 13, 256
x = torch.randn(n_img, h, w, c, device="cuda").to(torch.bfloat16)
out = torch.zeros(148, dtype=torch.int64, device="cuda")
print("tensor %.0f MB" % (x.numel() * 2 / 1e6))
for im2col in (0, 1):
    for rows in (32, 64, 128, 256):
        for stages in (2, 4, 8):
            if stages * rows * 128 > 200000:
                continue
            iters = 2000
            for rep in range(2):
                nv.check(L.vl_debug_tma_bench(x.data_ptr(), im2col, n_img, h, w, c, rows, stages, iters, 148, out.data_ptr(),
                                              ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
                torch.cuda.synchronize()
            cyc = out.float().mean().item()
            print("%s rows %3d stages %d: %7.1f clk/load  %5.2f clk/row  %5.1f B/clk/SM" % (
                "im2col" if im2col else "tiled ", rows, stages, cyc / iters, cyc / iters / rows, rows * 128 * iters / cyc))
