import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import vlb200
from vlb200 import _native as nv
L = nv.lib()
L.vl_debug_shift_mma.restype = ctypes.c_int32
L.vl_debug_shift_mma.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                 ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p]
torch.manual_seed(0)
w = torch.randn(128, 64, device="cuda").to(torch.bfloat16)
x = torch.randn(512, 64, device="cuda").to(torch.bfloat16)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for mode in (0, 1, 2, 3):
    line = []
    for off in (0, 1, 2, 3, 5, 7, 8, 9, 13, 33, 100, 131):
        n = 64 if not (mode & 2) else 64
        out = torch.zeros(128, n, device="cuda")
        nv.check(L.vl_debug_shift_mma(w.data_ptr(), x.data_ptr(), 512, off, n, mode, out.data_ptr(), st))
        torch.cuda.synchronize()
        if mode & 2:   # A shifted: D[i][j] = X[off+i] . W[j]
            ref = x[off:off + 128].float() @ w[:n].float().t()
        else:          # B shifted: D[i][j] = W[i] . X[off+j]
            ref = w.float() @ x[off:off + n].float().t()
        err = (out - ref).abs().max().item() / ref.abs().max().item()
        line.append("%d:%s" % (off, "ok" if err < 1e-2 else "BAD(%.2f)" % err))
    print("mode %d (%s%s): %s" % (mode, "shift A" if mode & 2 else "shift B", ", base_offset" if mode & 1 else "", " ".join(line)), flush=True)

# mode 4: N-major B with OVERLAPPING 64-wide atoms (leading byte offset = one pixel row): three taps of a filter row
# through one descriptor.  D[m][(s, c)] = sum_{k<64} W[m][k] * X[off + k + s][c]
line = []
for off in (0, 1, 2, 3, 5, 8, 9, 59, 118, 200):
    n = 192
    out = torch.zeros(128, n, device="cuda")
    nv.check(L.vl_debug_shift_mma(w.data_ptr(), x.data_ptr(), 512, off, n, 4, out.data_ptr(), st))
    torch.cuda.synchronize()
    ref = torch.cat([w.float() @ x[off + s:off + s + 64].float() for s in range(3)], dim=1)
    err = (out - ref).abs().max().item() / ref.abs().max().item()
    line.append("%d:%s" % (off, "ok" if err < 1e-2 else "BAD(%.2f)" % err))
print("mode 4 (N-major B, overlapping atoms, LBO = 128 B): %s" % " ".join(line), flush=True)
