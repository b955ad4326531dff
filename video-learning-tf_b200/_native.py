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
B_TILED_K, B_TILED_MN, B_IM2COL_MN = 0, 1, 2
DT_BF16, DT_F32 = 0, 1


class ConvGeom(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in (
        "n", "h", "w", "c", "kh", "kw", "stride_h", "stride_w", "pad_top", "pad_left", "p", "q", "cin_g",
        "flip_taps")]


class GemmDesc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in (
        "m", "n", "k", "groups", "a_mode", "b_mode", "a_ld", "b_ld", "a_goff", "b_goff", "c_goff",
        "b_tap_stride", "b_row_goff", "b_tap_inner", "c_ld", "c_dtype", "c_atomic", "relu", "split_k", "msub", "block_n",
        "mask_ld", "d2s_sh", "d2s_sw", "d2s_c", "d2s_h", "d2s_w", "row_shift")] + [
        ("conv", ConvGeom)]


class ConvFlatDesc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in (
        "n", "h", "w", "c", "kh", "kw", "pad_top", "pad_left", "pad_bottom", "pad_right", "groups", "cin_g", "cout_g",
        "flip_taps", "w_rows", "w_ld", "c_ld", "relu")]


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
    L.vl_host_alloc.restype = ctypes.c_void_p
    L.vl_host_alloc.argtypes = [i64, i32]
    L.vl_host_free.restype = i32
    L.vl_host_free.argtypes = [vp]
    L.vl_set_smem_reserve.restype = i32
    L.vl_set_smem_reserve.argtypes = [i32]
    L.vl_grad_sqnorms_workspace.restype = i64
    L.vl_grad_sqnorms_workspace.argtypes = [i64, i32]
    u64 = ctypes.c_uint64
    sigs = {
        "vl_gemm": [ctypes.POINTER(GemmDesc), vp, vp, vp, vp, vp, vp],
        "vl_zero": [vp, i64, vp],
        "vl_conv_flat": [ctypes.POINTER(ConvFlatDesc), vp, vp, vp, vp, vp],
        "vl_pack_dgrad_kmajor": [vp, vp, i32, i32, i32, i32, vp],
        "vl_pack_dgrad_d2s": [vp, vp, i32, i32, i32, i32, i32, i32, i32, vp],
        "vl_lrn_fwd": [vp, vp, i64, i32, i32, f32, f32, f32, vp],
        "vl_lrn_bwd": [vp, vp, vp, i64, i32, i32, f32, f32, f32, i32, vp],
        "vl_maxpool_fwd": [vp, vp, vp, i32, i32, i32, i32, vp],
        "vl_maxpool_bwd": [vp, vp, vp, vp, i32, i32, i32, i32, vp],
        "vl_colsum": [vp, vp, i64, i32, i32, vp],
        "vl_frames_s2d": [vp, i32, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp],
        "vl_frames_s2d_crop": [vp, i32, vp, vp, i32, i32, i32, vp, i32, i32, i32, i32, i32, i32, i32, vp],
        "vl_s2d_pack_filter": [vp, vp, i32, i32, i32, i32, i32, i32, i32, vp],
        "vl_pack_bf16_t": [vp, i32, i32, vp, i32, i32, i32, vp],
        "vl_s2d_unpack_grad": [vp, vp, i32, i32, i32, i32, i32, vp],
        "vl_lrn_pool_fwd_generic": [vp, vp, vp, i32, i32, i32, i32, i32, f32, f32, f32, vp],
        "vl_pool_lrn_bwd_generic": [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, f32, f32, f32, vp],
        "vl_lrn_pool_fwd": [vp, vp, vp, i32, i32, i32, i32, i32, f32, f32, f32, vp],
        "vl_pool_lrn_bwd": [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, f32, f32, f32, vp],
        "vl_lstm_fwd_cluster": [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, f32, vp],
        "vl_lstm_bwd_cluster": [vp, vp, vp, vp, vp, i32, i32, i32, vp],
        "vl_pack_bf16": [vp, i32, i32, vp, i32, i32, i32, i32, vp],
        "vl_cast_f32_to_bf16": [vp, vp, i64, vp],
        "vl_transpose_f32": [vp, vp, i32, i32, vp],
        "vl_gather_bf16": [vp, vp, vp, i64, vp],
        "vl_lstm_fwd": [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, f32, vp],
        "vl_lstm_bwd": [vp, vp, vp, vp, vp, i32, i32, i32, vp],
        "vl_lstm_fwd_ex": [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, f32, vp],
        "vl_lstm_bwd_ex": [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp],
        "vl_argmax_gather": [vp, i32, i32, i32, vp, i32, vp, vp, vp, vp],
        "vl_fuse_list": [vp, i32, i64, i32, vp, vp, vp],
        "vl_segment_pool_fwd": [vp, vp, i32, i32, i32, i32, vp, vp, vp],
        "vl_segment_pool_bwd": [vp, vp, i32, i32, i32, i32, vp, vp],
        "vl_segment_pool_fwd_bf16": [vp, i32, i32, i32, i32, vp, vp, vp],
        "vl_segment_pool_bwd_relu_bf16": [vp, vp, i32, i32, i32, i32, vp, vp],
        "vl_dropout_mask": [vp, i64, f32, u64, u64, vp],
        "vl_mul": [vp, vp, vp, vp, i64, vp],
        "vl_softmax_ce": [vp, vp, i32, i32, f32, vp, vp, vp, vp, i32, vp],
        "vl_grad_sqnorms": [vp, i64, vp, i32, vp, vp, i64, vp],
        "vl_clip_scalars": [vp, i32, f32, f32, vp, vp],
        "vl_resize_bilinear_u8": [vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, i32, vp, vp, i32, vp],
        "vl_sgd_update": [vp, vp, i64, f32, vp, f32, vp],
        "vl_sgd_update_shadow": [vp, vp, i64, f32, vp, f32, i32, vp, vp, vp, vp],
        "vl_adam_update": [vp, vp, vp, vp, i64, f32, f32, f32, f32, i32, vp, f32, vp],
        "vl_split3_act": [vp, vp, i64, i32, i32, vp],
        "vl_split3_weight": [vp, vp, i64, i32, i32, vp],
        "vl_gather_split_bf16": [vp, vp, vp, i64, vp],
        "vl_frames_s2d_f32": [vp, i32, vp, vp, i32, i32, i32, vp, i32, i32, i32, i32, i32, i32, i32, vp],
        "vl_lrn_pool_fwd_f32": [vp, vp, i32, i32, i32, i32, i32, f32, f32, f32, i32, vp],
    }
    for name, argtypes in sigs.items():
        fn = getattr(L, name)
        fn.restype = i32
        fn.argtypes = argtypes


EXPORTS = ["vl_host_alloc", "vl_host_free", "vl_set_smem_reserve", "vl_last_error", "vl_version", "vl_device_sm_count", "vl_launch_count", "vl_zero", "vl_gemm", "vl_conv_flat", "vl_pack_dgrad_kmajor", "vl_pack_dgrad_d2s",
           "vl_lrn_fwd", "vl_lrn_bwd", "vl_maxpool_fwd", "vl_maxpool_bwd", "vl_colsum", "vl_pack_bf16",
           "vl_cast_f32_to_bf16", "vl_transpose_f32", "vl_gather_bf16", "vl_lstm_fwd", "vl_lstm_bwd", "vl_lstm_fwd_ex", "vl_lstm_bwd_ex", "vl_argmax_gather", "vl_fuse_list", "vl_lrn_pool_fwd", "vl_pool_lrn_bwd",
           "vl_lstm_fwd_cluster", "vl_lstm_bwd_cluster", "vl_frames_s2d", "vl_frames_s2d_crop", "vl_pack_bf16_t", "vl_s2d_pack_filter", "vl_s2d_unpack_grad",
           "vl_lrn_pool_fwd_generic", "vl_pool_lrn_bwd_generic", "vl_segment_pool_fwd",
           "vl_segment_pool_bwd", "vl_segment_pool_fwd_bf16", "vl_segment_pool_bwd_relu_bf16", "vl_dropout_mask", "vl_mul", "vl_softmax_ce", "vl_grad_sqnorms", "vl_grad_sqnorms_workspace",
           "vl_clip_scalars", "vl_resize_bilinear_u8", "vl_sgd_update", "vl_sgd_update_shadow", "vl_adam_update",
           "vl_split3_act", "vl_split3_weight", "vl_gather_split_bf16", "vl_frames_s2d_f32", "vl_lrn_pool_fwd_f32"]


def check(status):
    if status != 0:
        raise NativeError("vlb200 error %d: %s" % (status, lib().vl_last_error().decode()))


def ptr(t):
    """Raw device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_handle():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


# Measurement hook (bench.py, tests/bringup): an object with begin(name, args, meta) / end() brackets every launch that
# goes through this module, on the stream the kernel is launched on.  `meta` is what the launch wrapper of kernels.py
# declared just before the call (label + algorithmic FLOPs of a contraction).  None in normal operation.
tracer = None
pending_meta = None


def declare(label, flops):
    """Called by the launch wrappers of kernels.py: label and ALGORITHMIC FLOPs (real extents, no padding) of the
    contraction launched next; consumed by the tracer (if any)."""
    global pending_meta
    pending_meta = (label, float(flops))


def _traced(name, args, fn):
    global pending_meta
    meta, pending_meta = pending_meta, None
    if tracer is None:
        return fn()
    tracer.begin(name, args, meta)
    try:
        return fn()
    finally:
        tracer.end()


def gemm(desc, a, b, c, bias=None, relu_mask=None):
    _traced("vl_gemm", (desc,), lambda: check(lib().vl_gemm(
        ctypes.byref(desc), ptr(a), ptr(b), ptr(c), ptr(bias), ptr(relu_mask), stream_handle())))


def conv_flat(desc, x, w, bias, out):
    _traced("vl_conv_flat", (desc,), lambda: check(lib().vl_conv_flat(
        ctypes.byref(desc), ptr(x), ptr(w), ptr(bias), ptr(out), stream_handle())))


def call(name, *args):
    """Invoke `name(*args, stream)` on the current torch stream and raise on a non-zero status."""
    conv = []
    for a in args:
        if a is None:
            conv.append(None)
        elif hasattr(a, "data_ptr"):
            conv.append(ctypes.c_void_p(a.data_ptr()))
        else:
            conv.append(a)
    _traced(name, args, lambda: check(getattr(lib(), name)(*conv, stream_handle())))


def host_staging_tensor(shape, dtype, write_combined=False):
    """A pinned host tensor for the frame feed backed by vl_host_alloc (torch sees cudaHostAlloc memory as pinned, so
    `copy_(non_blocking=True)` from it is asynchronous).  write_combined: see include/vlb200.h; fill it with
    `tensor.copy_(...)` / numpy writes only, never read it back on the host."""
    import numpy as np
    import torch
    n = int(np.prod(shape))
    itemsize = torch.empty(0, dtype=dtype).element_size()
    p = lib().vl_host_alloc(n * itemsize, 1 if write_combined else 0)
    if not p:
        raise NativeError("vlb200: %s" % lib().vl_last_error().decode())
    buf = (ctypes.c_uint8 * (n * itemsize)).from_address(p)
    t = torch.frombuffer(buf, dtype=dtype, count=n).view(*shape)
    t._vl_host_ptr = p  # (kept alive for the life of the process: staging buffers are allocated once)
    return t
