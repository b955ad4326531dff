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
/* Total number of kernel launches issued through this library since load (bench.py: gpu_launches). */
int64_t vl_launch_count(void);

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
  VL_B_TILED_K = 0,  /* B stored [N][K] row-major (K contiguous)  */
  VL_B_TILED_MN = 1  /* B stored [K][N] row-major (N contiguous)  */
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
  int32_t c_ld;          /* row pitch of C in elements                                              */
  int32_t c_dtype;       /* VL_DT_BF16 / VL_DT_F32                                                  */
  int32_t c_atomic;      /* 1: red.add into C (split-K filter gradients; C must be zeroed)          */
  int32_t relu;          /* 1: max(x,0) after bias                                                  */
  int32_t split_k;       /* >=1                                                                     */
  int32_t block_n;       /* 0 = choose; else multiple of 16 in [16,256]                             */
  int32_t mask_ld;       /* row pitch of relu_mask                                                  */
  vl_conv_geom conv;     /* used when a_mode is an im2col mode                                      */
} vl_gemm_desc;

/* C[m][n] = epilogue( sum_k A[m][k] * B[k][n] ) ; bias (fp32, per global column) and relu_mask
 * (bf16, C-shaped: C is zeroed where mask <= 0, i.e. tf ReluGrad) may be NULL. */
int vl_gemm(const vl_gemm_desc* desc, const void* a, const void* b, void* c, const float* bias,
            const void* relu_mask, vl_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VLB200_H_ */
