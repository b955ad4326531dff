"""Host-side launch wrappers: build the POD descriptors of include/vlb200.h and enqueue the sm_100a kernels.

PyTorch tensors are used only as device buffers (`data_ptr()`) and for the current stream; every FLOP of
the hot path runs in the hand-written kernels of `csrc/`.  Each wrapper names the TensorFlow op of the
reference it stands in for (file:line into /root/reference).
"""
import math

import torch

from . import _native as nv

BF16 = torch.bfloat16
F32 = torch.float32


# ----------------------------------------------------------------------------------------------------
# geometry helpers
# ----------------------------------------------------------------------------------------------------
def same_padding(in_size, k, stride):
    """TF `padding="SAME"` (alexnet.py:76,117,159,180,201): out = ceil(in/stride), pad split low/high."""
    out = -(-in_size // stride)
    total = max((out - 1) * stride + k - in_size, 0)
    return out, total // 2, total - total // 2


def valid_out(in_size, k, stride):
    """TF `padding="VALID"` (alexnet.py:98,139,211)."""
    return (in_size - k) // stride + 1


class ConvSpec(object):
    """Static description of one (possibly grouped) convolution of the AlexNet encoder."""

    def __init__(self, h, w, cin, cout, kh, kw, stride, groups, padding="SAME"):
        self.h, self.w, self.cin, self.cout = h, w, cin, cout
        self.kh, self.kw, self.stride, self.groups = kh, kw, stride, groups
        if padding == "SAME":
            self.p, self.pad_top, self.pad_bottom = same_padding(h, kh, stride)
            self.q, self.pad_left, self.pad_right = same_padding(w, kw, stride)
        else:  # VALID: the space-to-depth form of conv1 (vl_frames_s2d materialises the SAME padding)
            self.p, self.pad_top, self.pad_bottom = valid_out(h, kh, stride), 0, 0
            self.q, self.pad_left, self.pad_right = valid_out(w, kw, stride), 0, 0
        self.cin_g = cin // groups
        self.cout_g = cout // groups
        self.taps = kh * kw
        self.cchunks = -(-self.cin_g // 64)
        self.k_packed = self.taps * self.cchunks * 64  # rows of the packed (zero padded) filter matrix

    def geom(self, n):
        g = nv.ConvGeom()
        g.n, g.h, g.w, g.c = n, self.h, self.w, self.cin
        g.kh, g.kw = self.kh, self.kw
        g.stride_h = g.stride_w = self.stride
        g.pad_top, g.pad_left = self.pad_top, self.pad_left
        g.p, g.q = self.p, self.q
        g.cin_g = self.cin_g
        g.flip_taps = 0
        return g

    def geom_dgrad(self, n):
        """im2col walk over dY for the data gradient (stride 1 only): pad' = k-1-pad, taps flipped."""
        assert self.stride == 1
        g = nv.ConvGeom()
        g.n, g.h, g.w, g.c = n, self.p, self.q, self.cout
        g.kh, g.kw = self.kh, self.kw
        g.stride_h = g.stride_w = 1
        g.pad_top, g.pad_left = self.kh - 1 - self.pad_top, self.kw - 1 - self.pad_left
        g.p, g.q = self.h, self.w
        g.cin_g = self.cout_g
        g.flip_taps = 1
        return g


def _ld(t):
    assert t.dim() == 2 and t.stride(1) == 1
    return t.stride(0)


# ----------------------------------------------------------------------------------------------------
# dense layers: tf.nn.relu_layer / tf.nn.xw_plus_b (alexnet.py:228,248,275; tf_util.py:56; lstm.py:141 x-part)
# ----------------------------------------------------------------------------------------------------
def linear_fwd(x, w, bias, out, relu=False, n=None, block_n=0, msub=0, split_k=1):
    """out[M,N] = act(x[M,K] @ w[K,N] + bias).  x, w bf16 (w in TF [in,out] layout, row pitch % 8 == 0).
    block_n / msub = 0: the library's tile choice.  split_k > 1 (fp32 output, no ReLU): the contraction is split over
    split_k CTAs per tile that red.add into the zeroed output - for a tall-K, few-tile product (the LSTM input
    projection: 32 tiles of 64 serial k-blocks on 148 SMs) the k-block chain, not the tensor pipe, is the bound."""
    m, k = x.shape
    n = out.shape[1] if n is None else n
    d = nv.GemmDesc()
    d.m, d.n, d.k, d.groups = m, n, k, 1
    d.a_mode, d.b_mode = nv.A_TILED_K, nv.B_TILED_MN
    d.a_ld, d.b_ld = _ld(x), _ld(w)
    d.c_ld = _ld(out)
    d.c_dtype = nv.DT_BF16 if out.dtype == BF16 else nv.DT_F32
    d.relu = 1 if relu else 0
    d.split_k = 1
    if split_k > 1:
        assert out.dtype == F32 and not relu and out.is_contiguous()
        nv.call("vl_zero", out, out.numel() * 4)
        d.c_atomic = 1
        d.split_k = split_k
    d.block_n = block_n
    d.msub = msub
    nv.declare("dense fwd %dx%dx%d" % (m, n, k), 2.0 * m * n * k)
    nv.gemm(d, x, w, out, bias)
    return out


def linear_dgrad(dy, w, dx, relu_mask=None, n_contract=None, block_n=0):
    """dx[M,K] = dy[M,N] @ w[K,N]^T, optionally masked by (relu_mask > 0) (tf ReluGrad of the producer)."""
    m = dy.shape[0]
    nn = dy.shape[1] if n_contract is None else n_contract
    k = dx.shape[1]
    d = nv.GemmDesc()
    d.m, d.n, d.k, d.groups = m, k, nn, 1
    d.a_mode, d.b_mode = nv.A_TILED_K, nv.B_TILED_K
    d.a_ld, d.b_ld = _ld(dy), _ld(w)
    d.c_ld = _ld(dx)
    d.c_dtype = nv.DT_BF16 if dx.dtype == BF16 else nv.DT_F32
    d.split_k = 1
    d.block_n = block_n
    if relu_mask is not None:
        d.mask_ld = _ld(relu_mask)
    nv.declare("dense dgrad %dx%dx%d" % (m, k, nn), 2.0 * m * k * nn)
    nv.gemm(d, dy, w, dx, None, relu_mask)
    return dx


def linear_wgrad(x, dy, dw, split_k=1, n=None, block_n=0):
    """dw[K,N] (fp32) += x[M,K]^T @ dy[M,N]; the contraction runs over the batch rows M (split-K red.add)."""
    m, k = x.shape
    n = dy.shape[1] if n is None else n
    d = nv.GemmDesc()
    d.m, d.n, d.k, d.groups = dw.shape[0], n, m, 1  # dw may have fewer rows than x has (zero padded) columns
    assert dw.shape[0] <= k
    d.a_mode, d.b_mode = nv.A_TILED_MN, nv.B_TILED_MN
    d.a_ld, d.b_ld = _ld(x), _ld(dy)
    d.c_ld = _ld(dw)
    d.c_dtype = nv.DT_F32
    d.c_atomic = 1
    d.split_k = split_k
    d.block_n = block_n
    assert dw.dtype == F32
    nv.declare("dense wgrad %dx%dx%d" % (dw.shape[0], n, m), 2.0 * dw.shape[0] * n * m)
    nv.gemm(d, x, dy, dw)
    return dw


def conv_flops(spec, n):
    """Algorithmic FLOPs of one pass (forward, data gradient or filter gradient) of the convolution over n frames."""
    return 2.0 * n * spec.p * spec.q * spec.cout * spec.taps * spec.cin_g


# ----------------------------------------------------------------------------------------------------
# convolutions: dcnn.conv (alexnet.py:15-31) and its gradients
# ----------------------------------------------------------------------------------------------------
def conv_fwd(spec, x, w_kmajor, bias, out, relu=True, block_n=0, msub=0):
    """out[N,P,Q,Cout] = act(conv2d(x[N,H,W,Cin], W) + bias); w_kmajor = bf16 [Cout, taps*cchunks*64] (K-major:
    row = output channel, each tap's cin_g filter rows zero-padded to a multiple of 64)."""
    n = x.shape[0]
    d = nv.GemmDesc()
    d.m, d.n, d.k, d.groups = n * spec.p * spec.q, spec.cout_g, spec.k_packed, spec.groups
    d.a_mode, d.b_mode = nv.A_IM2COL_K, nv.B_TILED_K
    d.a_goff, d.b_goff, d.c_goff = spec.cin_g, 0, spec.cout_g
    d.b_row_goff = spec.cout_g
    d.b_tap_inner = spec.cchunks * 64
    d.b_ld = spec.k_packed
    d.c_ld = spec.cout
    d.c_dtype = nv.DT_BF16 if out.dtype == BF16 else nv.DT_F32
    d.relu = 1 if relu else 0
    d.split_k = 1
    d.block_n = block_n
    d.msub = msub
    d.conv = spec.geom(n)
    nv.declare("conv fwd (im2col) %dx%dx%d k%dx%d -> %d g%d" % (spec.h, spec.w, spec.cin, spec.kh, spec.kw, spec.cout,
                                                               spec.groups), conv_flops(spec, n))
    nv.gemm(d, x, w_kmajor, out, bias)
    return out


def conv_fwd_flat(spec, x, w_kmajor, bias, out, relu=True, flops=None):
    """Same contract as conv_fwd for stride-1 convolutions with cout_g <= 128, through the tap-shifted kernel
    (csrc/conv_flat.cu): the input tile is staged once per row tile instead of once per tap."""
    assert spec.stride == 1 and out.dtype == BF16
    n = x.shape[0]
    d = nv.ConvFlatDesc()
    d.n, d.h, d.w, d.c = n, spec.h, spec.w, spec.cin
    d.kh, d.kw = spec.kh, spec.kw
    d.pad_top, d.pad_left, d.pad_bottom, d.pad_right = spec.pad_top, spec.pad_left, spec.pad_bottom, spec.pad_right
    d.groups, d.cin_g, d.cout_g = spec.groups, spec.cin_g, spec.cout_g
    d.flip_taps = 0
    d.w_rows, d.w_ld = spec.cout, spec.k_packed
    d.c_ld = spec.cout
    d.relu = 1 if relu else 0
    nv.declare("conv fwd (tap-shifted) %dx%dx%d k%dx%d -> %d g%d" % (spec.h, spec.w, spec.cin, spec.kh, spec.kw,
                                                                    spec.cout, spec.groups),
               conv_flops(spec, n) if flops is None else flops)  # flops: the layer's own count (conv1 runs as s2d)
    nv.conv_flat(d, x, w_kmajor, bias, out)
    return out


def conv_dgrad_flat(spec, dy, w_dgrad_kmajor, dx):
    """dx = conv2d_backprop_input(dy, W) for a stride-1 SAME convolution with cin_g <= 128: a "full" correlation of
    dy with the flipped filter; w_dgrad_kmajor = vl_pack_dgrad_kmajor(W) [cin, taps * roundup64(cout_g)]."""
    assert spec.stride == 1 and dx.dtype == BF16
    n = dy.shape[0]
    d = nv.ConvFlatDesc()
    d.n, d.h, d.w, d.c = n, spec.p, spec.q, spec.cout
    d.kh, d.kw = spec.kh, spec.kw
    d.pad_top, d.pad_left = spec.kh - 1 - spec.pad_top, spec.kw - 1 - spec.pad_left
    d.pad_bottom, d.pad_right = spec.kh - 1 - spec.pad_bottom, spec.kw - 1 - spec.pad_right
    d.groups, d.cin_g, d.cout_g = spec.groups, spec.cout_g, spec.cin_g
    d.flip_taps = 1
    d.w_rows, d.w_ld = spec.cin, spec.taps * (-(-spec.cout_g // 64) * 64)
    d.c_ld = spec.cin
    d.relu = 0
    nv.declare("conv dgrad (tap-shifted) %dx%dx%d k%dx%d g%d" % (spec.h, spec.w, spec.cin, spec.kh, spec.kw, spec.groups),
               conv_flops(spec, n))
    nv.conv_flat(d, dy, w_dgrad_kmajor, None, dx)
    return dx


def d2s_filter_shape(spec, sh, sw):
    """Shape of the vl_pack_dgrad_d2s operand of `spec`: [groups*sh*sw*cin_g, (kh+sh-1)*(kw+sw-1)*roundup64(cout_g)]."""
    kpad = -(-spec.cout_g // 64) * 64
    return spec.groups * sh * sw * spec.cin_g, (spec.kh + sh - 1) * (spec.kw + sw - 1) * kpad


def conv_dgrad_d2s(spec, dy, w_d2s, dx, sh=2, sw=2, block_n=0, msub=0):
    """dx = conv2d_backprop_input(dy, W) of a stride-1 convolution with a NARROW cin_g (conv2: 48), issued as a
    stride-(sh,sw) forward convolution over dy that produces the sh*sw sub-positions of every output block at once
    (N = sh*sw*cin_g columns per UMMA instead of cin_g; the epilogue scatters them back, depth-to-space).
    w_d2s = vl_pack_dgrad_d2s(W) (see include/vlb200.h)."""
    assert spec.stride == 1 and dx.dtype == BF16 and spec.cin_g % 16 == 0
    n = dy.shape[0]
    kh2, kw2 = spec.kh + sh - 1, spec.kw + sw - 1
    kpad = -(-spec.cout_g // 64) * 64
    py, qx = -(-spec.h // sh), -(-spec.w // sw)
    d = nv.GemmDesc()
    d.m, d.n, d.k, d.groups = n * py * qx, sh * sw * spec.cin_g, kh2 * kw2 * kpad, spec.groups
    d.a_mode, d.b_mode = nv.A_IM2COL_K, nv.B_TILED_K
    d.a_goff, d.b_goff, d.c_goff = spec.cout_g, 0, spec.cin_g
    d.b_row_goff = sh * sw * spec.cin_g
    d.b_tap_inner = kpad
    d.b_ld = kh2 * kw2 * kpad
    d.c_ld = spec.cin
    d.c_dtype = nv.DT_BF16
    d.split_k = 1
    d.block_n = block_n
    d.msub = msub
    d.d2s_sh, d.d2s_sw, d.d2s_c, d.d2s_h, d.d2s_w = sh, sw, spec.cin_g, spec.h, spec.w
    g = nv.ConvGeom()
    g.n, g.h, g.w, g.c = n, spec.p, spec.q, spec.cout
    g.kh, g.kw = kh2, kw2
    g.stride_h, g.stride_w = sh, sw
    g.pad_top, g.pad_left = spec.kh - 1 - spec.pad_top, spec.kw - 1 - spec.pad_left
    g.p, g.q = py, qx
    g.cin_g = spec.cout_g
    g.flip_taps = 0
    d.conv = g
    nv.declare("conv dgrad (depth-to-space) %dx%dx%d k%dx%d g%d" % (spec.h, spec.w, spec.cin, spec.kh, spec.kw,
                                                                   spec.groups), conv_flops(spec, n))
    nv.gemm(d, dy, w_d2s, dx, None, None)
    return dx


def conv_dgrad(spec, dy, w_hwio, dx, relu_mask=None, block_n=0, msub=0):
    """dx[N,H,W,Cin] = conv2d_backprop_input(dy[N,P,Q,Cout], W); w_hwio = bf16 [taps*cin_g, Cout] (HWIO as 2D)."""
    n = dy.shape[0]
    d = nv.GemmDesc()
    d.m, d.n, d.k, d.groups = n * spec.h * spec.w, spec.cin_g, spec.taps * spec.cout_g, spec.groups
    d.a_mode, d.b_mode = nv.A_IM2COL_K, nv.B_TILED_K
    d.a_goff, d.b_goff, d.c_goff = spec.cout_g, spec.cout_g, spec.cin_g
    d.b_ld = spec.cout
    d.b_tap_stride = spec.cin_g
    d.c_ld = spec.cin
    d.c_dtype = nv.DT_BF16 if dx.dtype == BF16 else nv.DT_F32
    d.split_k = 1
    d.block_n = block_n
    d.msub = msub
    d.conv = spec.geom_dgrad(n)
    if relu_mask is not None:
        d.mask_ld = spec.cin
    nv.declare("conv dgrad (im2col) %dx%dx%d k%dx%d g%d" % (spec.h, spec.w, spec.cin, spec.kh, spec.kw, spec.groups),
               conv_flops(spec, n))
    nv.gemm(d, dy, w_hwio, dx, None, relu_mask)
    return dx


def conv_wgrad(spec, x, dy, dw, split_k=0, block_n=0, msub=0):
    """dw[taps*cin_g, Cout] (fp32, HWIO as 2D) += conv2d_backprop_filter(x, dy); split-K over output pixels."""
    n = x.shape[0]
    d = nv.GemmDesc()
    d.m, d.n, d.k, d.groups = spec.taps * spec.cin_g, spec.cout_g, n * spec.p * spec.q, spec.groups
    d.a_mode, d.b_mode = nv.A_IM2COL_MN, nv.B_TILED_MN
    d.a_goff, d.b_goff, d.c_goff = spec.cin_g, spec.cout_g, spec.cout_g
    d.b_ld = spec.cout
    d.c_ld = spec.cout
    d.c_dtype = nv.DT_F32
    d.c_atomic = 1
    d.split_k = split_k
    d.block_n = block_n
    d.msub = msub
    d.conv = spec.geom(n)
    assert dw.dtype == F32
    nv.declare("conv wgrad %dx%dx%d k%dx%d g%d" % (spec.h, spec.w, spec.cin, spec.kh, spec.kw, spec.groups),
               conv_flops(spec, n))
    nv.gemm(d, x, dy, dw)
    return dw


def conv_wgrad_t(spec, x, dy, dw, split_k=0, block_n=0, a_ld=0, row_shift=False, flops=None):
    """Same contract as conv_wgrad with the operands swapped: output channels (dy^T) on the M side, the (tap, channel
    chunk) axis of im2col(x)^T on the N side.  Every UMMA is 128 x block_n x 16 with block_n = 192 / 256 instead of
    a narrow cout-wide one, and the split-K red.adds of a warp coalesce."""
    n = x.shape[0]
    chunks = spec.taps * spec.cchunks
    d = nv.GemmDesc()
    d.m, d.n, d.k, d.groups = spec.cout_g, chunks * 64, n * spec.p * spec.q, spec.groups
    d.a_mode, d.b_mode = nv.A_TILED_MN, nv.B_IM2COL_MN
    d.a_goff, d.b_goff, d.c_goff = spec.cout_g, spec.cin_g, spec.cout_g
    d.a_ld = a_ld or spec.cout  # row pitch of dy (elements); > cout when the tensor is stored channel-padded
    d.c_ld = spec.cout
    d.c_dtype = nv.DT_F32
    d.c_atomic = 1
    d.split_k = split_k
    d.block_n = block_n or (192 if chunks % 3 == 0 else 256)
    d.msub = 1
    d.conv = spec.geom(n)
    if row_shift:  # one k-block per output row, the taps of a filter row over one staged input row (conv1)
        assert spec.q <= 64 and spec.cin_g <= 64 and spec.stride == 1 and spec.groups == 1
        d.row_shift = 1
        d.block_n = spec.kw * 64
    assert dw.dtype == F32
    nv.declare("conv wgrad (swapped%s) %dx%dx%d k%dx%d g%d" % (", row-shift" if row_shift else "", spec.h, spec.w,
                                                              spec.cin, spec.kh, spec.kw, spec.groups),
               conv_flops(spec, n) if flops is None else flops)
    nv.gemm(d, dy, x, dw)
    return dw


def pack_conv_weight_host(spec, w_hwio):
    """HWIO fp32 -> K-major bf16 [Cout, taps*cchunks*64] with each tap's cin_g rows zero-padded to a 64 multiple
    (torch restatement of vl_pack_bf16_t, for tests and probes)."""
    kh, kw, cin_g, cout = w_hwio.shape
    w = w_hwio.reshape(kh * kw, cin_g, cout)
    packed = torch.zeros(kh * kw, spec.cchunks * 64, cout, dtype=BF16, device=w_hwio.device)
    packed[:, :cin_g, :] = w.to(BF16)
    return packed.reshape(spec.k_packed, cout).t().contiguous()
