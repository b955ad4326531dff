"""ctypes binding of the C-ABI declared in include/vlb200.h.

There is no CPU fallback: if the CUDA library is missing every call raises (the product path must fail
loudly rather than route through the oracle).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libvlb200.so")

# enums (include/vlb200.h)
A_TILED_K, A_TILED_MN, A_IM2COL_K, A_IM2COL_MN = 0, 1, 2, 3
B_TILED_K, B_TILED_MN = 0, 1
DT_BF16, DT_F32 = 0, 1


class ConvGeom(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in (
        "n", "h", "w", "c", "kh", "kw", "stride_h", "stride_w", "pad_top", "pad_left", "p", "q", "cin_g",
        "flip_taps")]


class GemmDesc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in (
        "m", "n", "k", "groups", "a_mode", "b_mode", "a_ld", "b_ld", "a_goff", "b_goff", "c_goff",
        "b_tap_stride", "c_ld", "c_dtype", "c_atomic", "relu", "split_k", "block_n", "mask_ld")] + [
        ("conv", ConvGeom)]


class NativeError(RuntimeError):
    pass


_lib = None


def lib():
    """Load (once) and return the CUDA library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(
                "vlb200: %s is missing - build it with `python video-learning-tf_b200/build.py` "
                "(there is no CPU fallback)" % LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH)
        _declare(_lib)
    return _lib


def _declare(L):
    vp, i32, i64, f32 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_float
    L.vl_last_error.restype = ctypes.c_char_p
    L.vl_last_error.argtypes = []
    L.vl_version.restype = i32
    L.vl_device_sm_count.restype = i32
    L.vl_launch_count.restype = i64
    L.vl_gemm.restype = i32
    L.vl_gemm.argtypes = [ctypes.POINTER(GemmDesc), vp, vp, vp, vp, vp, vp]


def check(status):
    if status != 0:
        raise NativeError("vlb200 error %d: %s" % (status, lib().vl_last_error().decode()))


def ptr(t):
    """Raw device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_handle():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def gemm(desc, a, b, c, bias=None, relu_mask=None):
    check(lib().vl_gemm(ctypes.byref(desc), ptr(a), ptr(b), ptr(c), ptr(bias), ptr(relu_mask), stream_handle()))
