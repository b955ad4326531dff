"""TMA feed rate (bytes per clock per SM) of tiled vs im2col boxes of 64 channels for the pixel pitches of the encoder:
48 channels (conv1 space-to-depth input, 96 B pixel rows), 64, 96 (p1 / da1) and 256 channels."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import vlb200
from vlb200 import _native as nv
L = nv.lib()
L.vl_debug_tma_bench.restype = ctypes.c_int32
L.vl_debug_tma_bench.argtypes = [ctypes.c_void_p] + [ctypes.c_int32] * 9 + [ctypes.c_void_p, ctypes.c_void_p]
out = torch.zeros(148, dtype=torch.int64, device="cuda")
for (n_img, h, w, c) in ((512, 59, 59, 48), (512, 59, 59, 64), (512, 57, 57, 96), (1024, 28, 28, 96), (1024, 13, 13, 256)):
    x = torch.randn(n_img, h, w, c, device="cuda").to(torch.bfloat16)
    print("tensor [%d,%d,%d,%d] %.0f MB" % (n_img, h, w, c, x.numel() * 2 / 1e6))
    for im2col in (0, 1):
        for rows in (64, 128):
            stages, iters = 6, 2000
            for rep in range(2):
                nv.check(L.vl_debug_tma_bench(x.data_ptr(), im2col, n_img, h, w, c, rows, stages, iters, 148, out.data_ptr(),
                                              ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
                torch.cuda.synchronize()
            cyc = out.float().mean().item()
            real = min(c, 64) * 2
            print("  %s rows %3d: %7.1f clk/load  %5.1f B/clk/SM nominal  %5.1f B/clk/SM real" % (
                "im2col" if im2col else "tiled ", rows, cyc / iters, rows * 128 * iters / cyc, rows * real * iters / cyc))
    del x
