/*
 * vlb200 -- C-ABI of the B200-native LRCN hot path (AlexNet-fc7 -> LSTM -> pooling -> CE -> SGD/Adam).
 *
 * The reference (npit/video-learning-tf) has no FFI of its own: its hot path is the TensorFlow graph
 * executed by the two `sess.run` calls in run_task.py:44 (train) and run_task.py:95 (validation).
 * Every entry point below replaces the TensorFlow op (or op group) named in its comment, i.e. it is
 * what a TF-free `sess.run` would bind.  SURVEY.md section 8(b) lists the contract.
 *
 * Conventions
 *   - plain C, no torch types; every pointer is a raw DEVICE pointer unless the name ends in `_host`;
 *   - the caller owns every buffer; calls only enqueue work on `stream` (a cudaStream_t passed as void*);
 *   - return 0 on success, negative on error; `vl_last_error()` returns the message (thread local);
 *   - bf16 tensors are `uint16_t`-sized `__nv_bfloat16`, NHWC activations, HWIO filters (TF layouts).
 */
#ifndef VLB200_H_
#define VLB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* vl_stream_t; /* cudaStream_t */

const char* vl_last_error(void);
int vl_version(void);
/* Number of SMs of the current device (tile schedulers size their persistent grids with it). */
int vl_device_sm_count(void);
/* Shared memory (bytes per SM, rounded up to 1 KB, <= 96 KB) that the persistent contraction kernels (vl_gemm,
 * vl_conv_flat) leave unused from now on, so that CTAs of the issue-bound LRN / pool kernels launched on another stream
 * can be resident next to a contraction CTA instead of queueing behind it.  0 (default): the ring takes everything. */
int vl_set_smem_reserve(int32_t bytes);
/* Pinned host staging memory for the frame feed (cudaHostAlloc, portable).  write_combined != 0: not snooped by the CPU
 * caches (the host only writes the frames into it, the copy engine reads them); returns NULL on failure. */
void* vl_host_alloc(int64_t bytes, int32_t write_combined);
int vl_host_free(void* p);
/* Total number of kernel launches issued through this library since load (bench.py: gpu_launches). */
int64_t vl_launch_count(void);
/* cudaMemsetAsync(ptr, 0, bytes) on `stream`: the gradient arena is cleared once per step (tf.gradients starts from
 * zero; the split-K filter gradients accumulate with red.add). */
int vl_zero(void* ptr, int64_t bytes, vl_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Tensor-core contraction core (tcgen05.mma + TMEM accumulators + TMA operand staging).
 * One kernel serves every dense contraction of the path:
 *   tf.nn.conv2d            alexnet.py:21      (A = im2col(x) via TMA im2col mode, B = HWIO filter)
 *   tf.nn.relu_layer        alexnet.py:228,248 (fc6/fc7: bias + ReLU epilogue)
 *   tf.nn.xw_plus_b         alexnet.py:275, tf_util.py:56 (fc8 / output_fc)
 *   BasicLSTMCell x-part    lstm.py:17-19,141  ([x_t] * kernel[:D] for all t at once)
 *   and their gradients as produced by tf.gradients (train.py:210): data-gradient and filter-gradient.
 * ---------------------------------------------------------------------------------------------- */
enum {
  VL_A_TILED_K = 0,   /* A[M][K] row-major (K contiguous)                                   */
  VL_A_TILED_MN = 1,  /* A stored [K][M] row-major (M contiguous): x^T for filter gradients */
  VL_A_IM2COL_K = 2,  /* A = im2col(NHWC) rows = output pixels, K = (tap, channel)          */
  VL_A_IM2COL_MN = 3  /* A = im2col(NHWC)^T: M = (tap, channel), K = output pixels          */
};
enum {
  VL_B_TILED_K = 0,   /* B stored [N][K] row-major (K contiguous)  */
  VL_B_TILED_MN = 1,  /* B stored [K][N] row-major (N contiguous)  */
  VL_B_IM2COL_MN = 2  /* B = im2col(NHWC)^T: N = (tap, channel chunk of 64), K = output pixels; with A = dy^T
                         (VL_A_TILED_MN) this is the filter gradient with the output channels on the M side; C
                         [taps*cin_g][c_ld] fp32 is written transposed (row = (tap, ci), column = group column + m) */
};
enum { VL_DT_BF16 = 0, VL_DT_F32 = 1 };

typedef struct vl_conv_geom {
  int32_t n, h, w, c;        /* NHWC tensor the im2col loads walk; c = all channels (all groups)     */
  int32_t kh, kw;            /* filter taps                                                          */
  int32_t stride_h, stride_w;
  int32_t pad_top, pad_left; /* TF SAME: pad_total/2 before, rest after (alexnet.py:76,117,...)      */
  int32_t p, q;              /* output spatial extent                                                */
  int32_t cin_g;             /* channels per group that carry data (48 for conv2, padded to 64)      */
  int32_t flip_taps;         /* 1: B is addressed with the spatially flipped tap (data gradient)     */
} vl_conv_geom;

typedef struct vl_gemm_desc {
  int32_t m, n, k;       /* per-group extents; k = contraction length in elements                   */
  int32_t groups;        /* grouped conv = `groups` independent GEMMs in one launch                 */
  int32_t a_mode, b_mode;
  int32_t a_ld, b_ld;    /* row pitch in elements of tiled operands (ignored for im2col A)           */
  int32_t a_goff, b_goff, c_goff; /* per-group offset: A channel / inner coordinate, B inner coordinate, C column */
  int32_t b_tap_stride;  /* VL_B_TILED_K with conv: rows of B per filter tap (cin_g); 0 for dense   */
  int32_t b_row_goff;    /* VL_B_TILED_K: per-group ROW offset of B (K-major forward filters [cout][K])   */
  int32_t b_tap_inner;   /* VL_B_TILED_K with conv: inner (contraction) advance of B per filter tap; 0 when
                            the taps are addressed through rows (b_tap_stride, data gradients)            */
  int32_t c_ld;          /* row pitch of C in elements                                              */
  int32_t c_dtype;       /* VL_DT_BF16 / VL_DT_F32                                                  */
  int32_t c_atomic;      /* 1: red.add into C (split-K filter gradients; C must be zeroed)          */
  int32_t relu;          /* 1: max(x,0) after bias                                                  */
  int32_t split_k;       /* >=1; 0 = choose (atomic epilogue only)                                  */
  int32_t msub;          /* 128-row sub-tiles per CTA tile: 0 = choose, 1, or 2 (needs 2*block_n <= 256)   */
  int32_t block_n;       /* 0 = choose; else multiple of 16 in [16,256]                             */
  int32_t mask_ld;       /* row pitch of relu_mask                                                  */
  /* depth-to-space epilogue (d2s_c > 0; data gradient of a stride-1 convolution issued as a stride-(sh,sw)
   * forward convolution over dy, see vl_pack_dgrad_d2s): row (n,Y,X) of the p x q grid, column (dy,dx,ch) of
   * group g -> C[n][sh*Y+dy][sw*X+dx][g*c_goff + ch] of an NHWC tensor [n][d2s_h][d2s_w][c_ld]; rows / columns
   * beyond d2s_h / d2s_w are dropped.  d2s_c % 16 == 0. */
  int32_t d2s_sh, d2s_sw, d2s_c, d2s_h, d2s_w;
  /* VL_A_TILED_MN x VL_B_IM2COL_MN (swapped filter gradient) only: 1 = one k-block per OUTPUT ROW (q <= 64), an n-block =
   * the kw taps of one filter row, which contract against ONE tiled box of 64 + kw - 1 input pixels through an N-major
   * descriptor whose atoms overlap by one pixel row (3x less operand traffic than one im2col box per tap). */
  int32_t row_shift;
  vl_conv_geom conv;     /* used when a_mode is an im2col mode                                      */
} vl_gemm_desc;

/* C[m][n] = epilogue( sum_k A[m][k] * B[k][n] ) ; bias (fp32, per global column) and relu_mask
 * (bf16, C-shaped: C is zeroed where mask <= 0, i.e. tf ReluGrad) may be NULL. */
int vl_gemm(const vl_gemm_desc* desc, const void* a, const void* b, void* c, const float* bias,
            const void* relu_mask, vl_stream_t stream);


/* Tap-shifted stride-1 convolution for the narrow-channel layers (conv1 in space-to-depth form, conv2, conv2's data
 * gradient): one tiled TMA box per (row tile, 64-channel chunk) with the padding materialised by out-of-bounds zero
 * fill, every filter tap = the same shared-memory tile shifted by (r*Wp+s) rows, output channels on the M side and
 * flat positions on the N side of tcgen05.mma.  out[n][Ho][Wo][c_ld] (bf16) = act(conv(x, w) + bias) with
 * Ho = h + pad_top + pad_bottom - kh + 1, Wo = w + pad_left + pad_right - kw + 1 (alexnet.py:15-31).
 * w_kmajor: bf16 [w_rows][w_ld], row = (group, output channel), column = (tap, input channel padded to 64). */
typedef struct vl_conv_flat_desc {
  int32_t n, h, w, c;                 /* NHWC input; c = all channels (all groups)                          */
  int32_t kh, kw;
  int32_t pad_top, pad_left, pad_bottom, pad_right;
  int32_t groups, cin_g, cout_g;      /* contraction / output channels per group                           */
  int32_t flip_taps;                  /* 1: tap (r,s) uses the filter of tap (kh-1-r, kw-1-s) (data gradient) */
  int32_t w_rows, w_ld;               /* rows and row pitch (elements) of w_kmajor                          */
  int32_t c_ld;                       /* channel pitch of out                                               */
  int32_t relu;
} vl_conv_flat_desc;
int vl_conv_flat(const vl_conv_flat_desc* desc, const void* x, const void* w_kmajor, const float* bias, void* out,
                 vl_stream_t stream);
/* K-major filter of the data gradient: dst[g*cin_g+ci][tap*kpad+co] = src[tap*cin_g+ci][g*cout_g+co], kpad =
 * cout_g rounded up to 64; src = HWIO fp32 as 2-D [taps*cin_g][groups*cout_g]. */
int vl_pack_dgrad_kmajor(const float* src, void* dst, int32_t taps, int32_t cin_g, int32_t cout_g, int32_t groups,
                         vl_stream_t stream);
/* K-major filter of the depth-to-space data gradient of a stride-1 convolution (conv2: 5x5, 48 -> 128 per group).
 * din[sh*Y+dy][sw*X+dx][c] = sum over the (kh+sh-1) x (kw+sw-1) taps (ty,tx) of a stride-(sh,sw) walk over dout and over
 * the cout_g channels k of  dout[sh*Y-(kh-1-pad_top)+ty][sw*X-(kw-1-pad_left)+tx][k] * W[kh-1+dy-ty][kw-1+dx-tx][c][k]
 * (zero where the filter index leaves [0,kh) x [0,kw)), so that one 128 x (sh*sw*cin_g) x 16 UMMA tile does the work of
 * sh*sw narrow (N = cin_g) tiles.  src = HWIO fp32 [kh][kw][cin_g][groups*cout_g];
 * dst = bf16 [groups*sh*sw*cin_g][(kh+sh-1)*(kw+sw-1)*kpad], kpad = cout_g rounded up to 64,
 * row = g*sh*sw*cin_g + (dy*sw+dx)*cin_g + c, column = (ty*(kw+sw-1)+tx)*kpad + k. */
int vl_pack_dgrad_d2s(const float* src, void* dst, int32_t kh, int32_t kw, int32_t cin_g, int32_t cout_g, int32_t groups,
                      int32_t sh, int32_t sw, vl_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Memory-bound kernels of the AlexNet encoder (HBM roofline).
 * ---------------------------------------------------------------------------------------------- */

/* Space-to-depth staging of the conv1 input (alexnet.py:60-77, dataset_.py:521-530): frames [n][h][w][3] (uint8,
 * mean subtracted here, or fp32 as fed by feeder.py:97-100) -> bf16 out[n][hb][wb][s*s*3] with
 * out[n][by][bx][(dy*s+dx)*3+c] = frame[n][s*by-pad_top+dy][s*bx-pad_left+dx][c] (0 outside the image), so that the
 * kh x kw stride-s SAME convolution equals a ceil(kh/s) x ceil(kw/s) stride-1 VALID convolution over s*s*3 channels. */
int vl_frames_s2d(const void* frames, int32_t is_u8, const float* mean3, void* out, int32_t n, int32_t h, int32_t w,
                  int32_t s, int32_t pad_top, int32_t pad_left, int32_t hb, int32_t wb, vl_stream_t stream);
/* Same staging with the reference's read-time preprocessing fused in (dataset_.py:444-461 crop_image, :498-500
 * rand_mirror, :521-530 sub_mean): frames are stored as [n][hr][wr][3]; crops = DEVICE int32 [n][3] = (y0, x0, mirror)
 * selects the h x w window at (y0, x0) of frame i, read right-to-left when mirror != 0 (y0 / x0 are clamped into
 * [0, hr - h] / [0, wr - w] by the kernel: it never reads outside the frame buffer).  crops may be NULL when
 * hr == h and wr == w. */
int vl_frames_s2d_crop(const void* frames, int32_t is_u8, const float* mean3, void* out, int32_t n, int32_t hr,
                       int32_t wr, const int32_t* crops, int32_t h, int32_t w, int32_t s, int32_t pad_top,
                       int32_t pad_left, int32_t hb, int32_t wb, vl_stream_t stream);
/* Read-time resampling of the reference (dataset_.py:238,484,491; serialize.py:425): scipy.misc.imresize(image, shape)
 * with its default bilinear filter = Pillow's 8-bit ImagingResample: horizontal pass, then vertical pass, each rounded
 * to uint8, 22-bit fixed-point coefficients.  in: uint8 [n][h_in][w_in][channels] -> out: uint8 [n][h_out][w_out][channels]
 * (bit-exact against PIL).  bounds_* = DEVICE int32 [out][2] (first input index, tap count), coeffs_* = DEVICE int32
 * [out][ksize_*], computed by the host exactly as Pillow's precompute_coeffs (resize.py: pil_bilinear_coeffs); an axis
 * whose size does not change needs none.  tmp = uint8 [n][h_in][w_out][channels] scratch when both axes change. */
int vl_resize_bilinear_u8(const void* in, void* out, void* tmp, int32_t n, int32_t h_in, int32_t w_in, int32_t h_out,
                          int32_t w_out, int32_t channels, const int32_t* bounds_w, const int32_t* coeffs_w,
                          int32_t ksize_w, const int32_t* bounds_h, const int32_t* coeffs_h, int32_t ksize_h,
                          vl_stream_t stream);
/* Filter of that convolution: HWIO fp32 [kh][kw][cin][cout] -> bf16 [taps][chunk][cout] (chunk >= s*s*cin rows per
 * tap, zero padded), and the inverse scatter of its filter gradient dws[taps*s*s*cin][cout] -> HWIO dw. */
int vl_s2d_pack_filter(const float* src, void* dst, int32_t kh, int32_t kw, int32_t cin, int32_t cout, int32_t s,
                       int32_t chunk, int32_t transpose /* 1: dst is K-major [cout][taps*chunk] */, vl_stream_t stream);
int vl_s2d_unpack_grad(const float* dws, float* dw, int32_t kh, int32_t kw, int32_t cin, int32_t cout, int32_t s,
                       vl_stream_t stream);

/* tf.nn.local_response_normalization(depth_radius=2, alpha=2e-5, beta=.75, bias=1) (alexnet.py:85,126) over
 * the channel axis of x[rows][c] (bf16). */
int vl_lrn_fwd(const void* x, void* y, int64_t rows, int32_t c, int32_t radius, float alpha, float beta,
               float bias, vl_stream_t stream);
/* Gradient of the above w.r.t. x, multiplied by the ReLU mask (x > 0) of the producing conv (relu1/relu_2). */
int vl_lrn_bwd(const void* x, const void* dy, void* dx, int64_t rows, int32_t c, int32_t radius, float alpha,
               float beta, float bias, int32_t relu_mask, vl_stream_t stream);

/* tf.nn.max_pool(ksize 3x3, stride 2, VALID) (alexnet.py:98,139,211) on NHWC bf16; `argmax` (uint8, window
 * local index r*3+s of the first maximum) is kept for the gradient. */
int vl_maxpool_fwd(const void* x, void* y, void* argmax, int32_t n, int32_t h, int32_t w, int32_t c,
                   vl_stream_t stream);
/* MaxPoolGrad: dx[n][h][w][c] = sum of dy over the windows whose argmax is (h,w); optional ReLU mask
 * (relu_of > 0, same shape as dx) for pool5 whose input is relu5. */
int vl_maxpool_bwd(const void* dy, const void* argmax, void* dx, const void* relu_of, int32_t n, int32_t h,
                   int32_t w, int32_t c, vl_stream_t stream);

/* Fused LRN -> max-pool forward (alexnet.py:85-98,126-139): y = max_pool(lrn(x)); lrn(x) is never written to HBM.
 * Same result as vl_lrn_fwd followed by vl_maxpool_fwd (lrn(x) is rounded to bf16 before the max in both). */
int vl_lrn_pool_fwd(const void* x, void* y, void* argmax, int32_t n, int32_t h, int32_t w, int32_t c,
                    int32_t radius, float alpha, float beta, float bias, vl_stream_t stream);
/* Fused MaxPoolGrad -> LRNGrad -> ReluGrad: dx = relu'(x) * lrn_grad(x, maxpool_grad(dy, argmax)); when dbias is
 * not NULL it also accumulates the bias gradient sum_pixels dx (dbias must be zeroed). */
int vl_pool_lrn_bwd(const void* x, const void* dy, const void* argmax, void* dx, float* dbias, int32_t n, int32_t h,
                    int32_t w, int32_t c, int32_t radius, float alpha, float beta, float bias, vl_stream_t stream);
/* Any-channel-count versions of the two fused kernels (one thread per 16-byte chunk, halo re-loaded from global
 * memory); vl_lrn_pool_fwd / vl_pool_lrn_bwd dispatch to them when c > 256. */
int vl_lrn_pool_fwd_generic(const void* x, void* y, void* argmax, int32_t n, int32_t h, int32_t w, int32_t c,
                            int32_t radius, float alpha, float beta, float bias, vl_stream_t stream);
int vl_pool_lrn_bwd_generic(const void* x, const void* dy, const void* argmax, void* dx, float* dbias, int32_t n,
                            int32_t h, int32_t w, int32_t c, int32_t radius, float alpha, float beta, float bias,
                            vl_stream_t stream);

/* bias gradient: out[c] += sum_rows dy[row][c]  (bf16 in, fp32 accumulate; `out` must be zeroed). */
int vl_colsum(const void* dy, float* out, int64_t rows, int32_t c, int32_t ld, vl_stream_t stream);

/* fp32 master weights -> bf16 operand copy: dst[(r/src_grp)*dst_grp + r%src_grp][col] = src[r][col], zero
 * elsewhere (dst has dst_rows x dst_ld elements).  Used for the zero-padded conv1/conv2 filter matrices and
 * for padding the class dimension of fc8/output_fc to a multiple of 8. */
int vl_pack_bf16(const float* src, int32_t rows, int32_t cols, void* dst, int32_t dst_rows, int32_t dst_ld,
                 int32_t src_grp, int32_t dst_grp, vl_stream_t stream);
/* K-major (transposed) form of the same packing: dst[col][(r/src_grp)*dst_grp + r%src_grp] = src[r][col], dst row
 * pitch dst_ld; the operand layout of the forward convolutions (one TMA box per k-block). */
int vl_pack_bf16_t(const float* src, int32_t rows, int32_t cols, void* dst, int32_t dst_ld, int32_t src_grp,
                   int32_t dst_grp, vl_stream_t stream);
int vl_cast_f32_to_bf16(const float* src, void* dst, int64_t n, vl_stream_t stream);
/* dst[i] = bf16(src[table[i]]), 0 where table[i] < 0; n % 8 == 0, table / dst 16-byte aligned.  Refreshes every
 * permuted / zero-padded bf16 operand copy of the convolution filters (the layouts of vl_s2d_pack_filter,
 * vl_pack_bf16_t, vl_pack_dgrad_d2s, vl_pack_bf16, vl_cast_f32_to_bf16) in ONE launch from a precomputed index table;
 * replaces the per-variable tf.assign traffic behind opt.apply_gradients (train.py:217). */
int vl_gather_bf16(const float* src, const int32_t* table, void* dst, int64_t n, vl_stream_t stream);

/* dst[cols][rows] = src[rows][cols]^T (fp32): recurrent weights for the BPTT kernel. */
int vl_transpose_f32(const float* src, float* dst, int32_t rows, int32_t cols, vl_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * LSTM (models/lstm/lstm.py:9-20,102-143): MultiRNNCell([BasicLSTMCell]) under dynamic_rnn, zero state.
 * gx[b*t_len+t][4h] (fp32) = x_t * kernel[:d] + bias, produced by vl_gemm; this kernel runs the recurrent
 * part for all timesteps: g = gx + h_{t-1} * w_h ; i,j,f,o = split(g) ; c = c*sig(f+forget_bias)+sig(i)*tanh(j);
 * h = tanh(c)*sig(o).  Saved for BPTT: acts[b][t][4h] (post-nonlinearity i,j,f,o), cs[b][t][h] (c_t),
 * h_seq[b][t][h] fp32 outputs, h_prev_bf16[b][t][h] (h_{t-1}, the filter-gradient operand).
 * ---------------------------------------------------------------------------------------------- */
int vl_lstm_fwd(const float* gx, const float* w_h, float* acts, float* cs, float* h_seq, void* h_seq_bf16,
                void* h_prev_bf16, int32_t batch, int32_t t_len, int32_t hidden, float forget_bias,
                vl_stream_t stream);
/* BPTT: dh_seq[b][t][h] (fp32) is d(loss)/d(h_t) from above; writes dg[b*t_len+t][4h] (bf16, gradient w.r.t.
 * the pre-activation gates) which feeds the tensor-core data/filter gradient GEMMs.  w_h_t = w_h^T [4h][h]. */
int vl_lstm_bwd(const float* dh_seq, const float* acts, const float* cs, const float* w_h_t, void* dg,
                int32_t batch, int32_t t_len, int32_t hidden, vl_stream_t stream);
/* Captioning variants (lstm.py:102-143 evaluate_sequence with nonzero_per_sequence and init_state; :145-265
 * generate_feedback_sequence): the same recurrence with an INITIAL STATE h0 / c0 [batch][h] (NULL = zero; the
 * reference builds LSTMStateTuple(v, v) per layer from the visual vector, lstm.py:34-42), per-sequence LENGTHS
 * (int32 [batch], NULL = all t_len; beyond its length a sequence carries its state through unchanged and emits zero
 * outputs, as dynamic_rnn does) and the FINAL STATE h_last / c_last [batch][h].  t_len = 1 with h0 / c0 is one step
 * of the greedy feedback decode.  Backward: dh_last / dc_last = gradient w.r.t. the final state, dh0 / dc0 (written
 * when non-NULL) = gradient w.r.t. the initial state. */
int vl_lstm_fwd_ex(const float* gx, const float* w_h, const float* h0, const float* c0, const int32_t* lengths,
                   float* acts, float* cs, float* h_seq, void* h_seq_bf16, void* h_prev_bf16, float* h_last,
                   float* c_last, int32_t batch, int32_t t_len, int32_t hidden, float forget_bias, vl_stream_t stream);
int vl_lstm_bwd_ex(const float* dh_seq, const float* dh_last, const float* dc_last, const float* acts, const float* cs,
                   const float* c0, const float* w_h_t, const int32_t* lengths, void* dg, float* dh0, float* dc0,
                   int32_t batch, int32_t t_len, int32_t hidden, vl_stream_t stream);
/* get_embedding_from_logits (lstm.py:257-265): index[r] = argmax(logits[r][:v]) (lowest index on ties, tf.arg_max),
 * out[r][:] = embedding[index[r]][:] (fp32 and / or bf16 copy).  logits fp32 with row pitch ld. */
int vl_argmax_gather(const float* logits, int32_t rows, int32_t v, int32_t ld, const float* embedding, int32_t e,
                     int64_t* index, float* out, void* out_bf16, vl_stream_t stream);
/* Persistent variants for hidden == 256: a cluster of 8 CTAs keeps the recurrent weights (fp32, kernel[d:], NOT
 * transposed for both directions) resident in shared memory for all timesteps of its 8 clips and exchanges h_t /
 * dh_t through distributed shared memory.  Same results as vl_lstm_fwd / vl_lstm_bwd. */
int vl_lstm_fwd_cluster(const float* gx, const float* w_h, float* acts, float* cs, float* h_seq, void* h_seq_bf16,
                        void* h_prev_bf16, int32_t batch, int32_t t_len, int32_t hidden, float forget_bias,
                        vl_stream_t stream);
int vl_lstm_bwd_cluster(const float* dh_seq, const float* acts, const float* cs, const float* w_h, void* dg,
                        int32_t batch, int32_t t_len, int32_t hidden, vl_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Pooling / loss head.
 * ---------------------------------------------------------------------------------------------- */
enum { VL_POOL_AVG = 0, VL_POOL_LAST = 1, VL_POOL_MAX = 2 /* extension: not in the reference */ };

/* Segmented reduction over rows: y[s][d] = pool(x[seg[s] .. seg[s+1])[d]).  Serves apply_temporal_fusion
 * (tf_util.py:4-30), aggregate_clip_vectors (tf_util.py:126-133) and the clip->video fusion of
 * val.py:158-167.  Rows are accumulated sequentially in fp32 and divided once, i.e. the operation order of
 * numpy's mean(axis=0), so video-level results are bit-exact with the reference's host code.
 * seg == NULL means equal segments of `fixed_len` rows. */
int vl_segment_pool_fwd(const float* x, const int32_t* seg, int32_t fixed_len, int32_t num_seg, int32_t d,
                        int32_t mode, float* y, void* y_bf16, vl_stream_t stream);
int vl_segment_pool_bwd(const float* dy, const int32_t* seg, int32_t fixed_len, int32_t num_seg, int32_t d,
                        int32_t mode, float* dx, vl_stream_t stream);

/* Early frame fusion (models/model.py:103-108: aggregate_clip_vectors on the dcnn features, tf_util.py:126-133):
 * x bf16 [num_seg * fixed_len][d] -> y fp32 / bf16 [num_seg][d], avg (rows added in order, one division) or last; and
 * its gradient fused with the ReLU gradient of the producing layer: dx bf16 = (act > 0) ? pool'(dy) : 0. */
int vl_segment_pool_fwd_bf16(const void* x, int32_t fixed_len, int32_t num_seg, int32_t d, int32_t mode, float* y,
                             void* y_bf16, vl_stream_t stream);
int vl_segment_pool_bwd_relu_bf16(const float* dy, const void* act, int32_t fixed_len, int32_t num_seg, int32_t d,
                                  int32_t mode, void* dx, vl_stream_t stream);

/* Multi-input pipelines (tf_util.py:136-147 apply_tensor_list_fusion, methods avg / maximum): element-wise mean
 * (sum in list order, one division) or maximum over a HOST array of k <= 8 DEVICE pointers to fp32 tensors of n
 * elements each; mode = VL_POOL_AVG / VL_POOL_MAX. */
int vl_fuse_list(const float* const* inputs, int32_t k, int64_t n, int32_t mode, float* y, void* y_bf16,
                 vl_stream_t stream);

/* tf.nn.dropout (lstm.py:50-56): y = x * mask, mask in {0, 1/keep}.  The mask is generated on device from
 * (seed, offset) with Philox4x32-10 and returned so that the backward pass (and the oracle) can reuse it. */
int vl_dropout_mask(float* mask, int64_t n, float keep_prob, uint64_t seed, uint64_t offset, vl_stream_t stream);
int vl_mul(const float* a, const float* b, float* y, void* y_bf16, int64_t n, vl_stream_t stream);

/* mean(softmax_cross_entropy_with_logits) (train.py:121-123) + accuracy (train.py:145-147) + d(loss)/d(logits).
 * logits[rows][c] fp32, labels[rows][c] int32 one/multi-hot; row_loss is a scratch/result buffer of 2*rows floats
 * (per-row loss, then per-row correct flag); out_scalars[0] = grad_scale * sum of the row losses, [1] = #correct;
 * dlogits = (softmax * sum(labels) - labels) * grad_scale.  grad_scale = 1 / (rows of the GLOBAL batch): on one
 * rank out_scalars[0] is the mean loss, over data-parallel ranks the SUM of out_scalars[0] is. */
int vl_softmax_ce(const float* logits, const int32_t* labels, int32_t rows, int32_t c, float grad_scale,
                  float* row_loss, float* out_scalars, float* dlogits, void* dlogits_bf16, int32_t dl_ld,
                  vl_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Optimiser (train.py:199-222) on one flat fp32 parameter / gradient arena.
 * seg_offsets[num_vars+1] (DEVICE array) delimit the variables inside the arena of n floats.
 * vl_grad_sqnorms writes sum(g^2) per variable (fp32) through a deterministic two-stage reduction (no float atomics:
 * every data-parallel rank must derive the same clip scale from the same reduced gradients); workspace = DEVICE
 * scratch of vl_grad_sqnorms_workspace(n, num_vars) floats, private to the call while it runs.
 * vl_sgd_update / vl_adam_update apply  g' = g * clip_scale  where clip_scale = clip/max(gnorm, clip) is read
 * from DEVICE memory (scale_dev[0]) so the step needs no host sync.
 * ---------------------------------------------------------------------------------------------- */
int64_t vl_grad_sqnorms_workspace(int64_t n, int32_t num_vars);
int vl_grad_sqnorms(const float* grads, int64_t n, const int64_t* seg_offsets, int32_t num_vars, float* sqnorms,
                    float* workspace, int64_t workspace_floats, vl_stream_t stream);
/* scalars[0]=global norm, [1]=clip scale, [2]=mean_i ||g_i * scale||  (grads_norm summary). clip<=0: no clip. */
int vl_clip_scalars(const float* sqnorms, int32_t num_vars, float clip_norm, float grad_prescale, float* scalars,
                    vl_stream_t stream);
int vl_sgd_update(float* params, const float* grads, int64_t n, float lr, const float* scalars,
                  float grad_prescale, vl_stream_t stream);
/* vl_sgd_update that also stores the bf16 operand copy of up to 4 arena ranges [seg_begin[k], seg_end[k]) (HOST arrays,
 * float4 aligned) into seg_dst[k] (device, same layout as the master: fc / LSTM kernels), saving the separate cast. */
int vl_sgd_update_shadow(float* params, const float* grads, int64_t n, float lr, const float* scalars,
                         float grad_prescale, int32_t num_segs, const int64_t* seg_begin, const int64_t* seg_end,
                         void* const* seg_dst, vl_stream_t stream);
int vl_adam_update(float* params, const float* grads, float* m, float* v, int64_t n, float lr, float beta1,
                   float beta2, float eps, int32_t step, const float* scalars, float grad_prescale,
                   vl_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * fp32-accuracy forward mode (csrc/fp32_path.cu; host side: video-learning-tf_b200/fp32_path.py).
 * The reference evaluates its graph in fp32 (models/alexnet/alexnet.py:60-280, models/lstm/lstm.py:59-143); the tensor
 * cores take bf16 operands.  For the "<= 1e-3 relative in fp32" tolerance every fp32 tensor X is split into
 * X_hi = bf16(X), X_lo = bf16(X - X_hi) and X * W is evaluated as X_hi W_hi + X_lo W_hi + X_hi W_lo by ONE vl_gemm over
 * operands concatenated along the contraction axis: [X_hi | X_lo | X_hi] x [W_hi ; W_hi ; W_lo] (fp32 accumulation and
 * output, bias / ReLU in the epilogue).  These entry points build those operands and run the non-contraction layers in
 * fp32.
 * ---------------------------------------------------------------------------------------------- */
/* x[rows][c] fp32 (c = groups * cg) -> out[rows][3c] bf16, group g = [hi(cg) | lo(cg) | hi(cg)] at column g * 3 * cg. */
int vl_split3_act(const float* x, void* out, int64_t rows, int32_t c, int32_t groups, vl_stream_t stream);
/* w[rows][cols] fp32 ([in, out] layout of tf.nn.xw_plus_b, alexnet.py:228,248,275) -> out[3 * rows][dst_ld] bf16 =
 * [hi ; hi ; lo], columns >= cols zero. */
int vl_split3_weight(const float* w, void* out, int64_t rows, int32_t cols, int32_t dst_ld, vl_stream_t stream);
/* vl_gather_bf16 with a hi / lo selector: dst[i] = hi or lo part (bit 30 of table[i] set: lo) of src[table[i] & 0x3fffffff],
 * 0 where table[i] < 0.  Builds the permuted K-major convolution filters [W_hi | W_hi | W_lo] of the mode in one launch. */
int vl_gather_split_bf16(const float* src, const int32_t* table, void* dst, int64_t n, vl_stream_t stream);
/* vl_frames_s2d_crop with an fp32 result (same arguments and semantics; dataset_.py:444-461,498-500,521-530). */
int vl_frames_s2d_f32(const void* frames, int32_t is_u8, const float* mean3, float* out, int32_t n, int32_t hr, int32_t wr,
                      const int32_t* crops, int32_t h, int32_t w, int32_t s, int32_t pad_top, int32_t pad_left, int32_t hb,
                      int32_t wb, vl_stream_t stream);
/* tf.nn.lrn (alexnet.py:80-89,121-130) + tf.nn.max_pool 3x3 / 2 VALID (alexnet.py:98,139,211) on fp32 NHWC tensors:
 * y[n][(h-3)/2+1][(w-3)/2+1][c]; with_lrn == 0: max-pool only (pool5). */
int vl_lrn_pool_fwd_f32(const float* x, float* y, int32_t n, int32_t h, int32_t w, int32_t c, int32_t radius, float alpha,
                        float beta, float bias, int32_t with_lrn, vl_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VLB200_H_ */
