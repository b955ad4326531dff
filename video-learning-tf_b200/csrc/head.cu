// Pooling / loss head of the LRCN path.
//
//  - segmented avg/last(/max) reduction: apply_temporal_fusion (tf_util.py:4-30), aggregate_clip_vectors
//    (tf_util.py:126-133) and the clip->video fusion of Validation.apply_clip_fusion (val.py:158-167);
//  - dropout (lstm.py:50-56) with a Philox4x32-10 mask generated on device;
//  - softmax cross-entropy + accuracy + d(logits) (train.py:117-124,142-149), one warp per clip with shuffle
//    reductions.
#include "common.cuh"
#include "../../include/vlb200.h"

#include <atomic>
#include <cstring>

namespace vl {
extern std::atomic<long long> g_launches;
}

namespace {

typedef __nv_bfloat16 bf16;
using vl::ptx::warp_max;
using vl::ptx::warp_sum;

// y[s][d] = pool over rows [lo, hi) -- rows accumulated in order, one rounding per add, one division:
// exactly numpy's `np.mean(rows, axis=0)` on float32 (val.py:161).
__global__ void segment_pool_fwd_kernel(const float* __restrict__ x, const int32_t* __restrict__ seg, int fixed_len,
                                        int num_seg, int d, int mode, float* __restrict__ y, bf16* __restrict__ y_bf16) {
  const long long total = (long long)num_seg * d;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int s = (int)(idx / d);
    const int col = (int)(idx - (long long)s * d);
    const long long lo = seg ? seg[s] : (long long)s * fixed_len;
    const long long hi = seg ? seg[s + 1] : lo + fixed_len;
    float out = 0.f;
    if (hi > lo) {
      if (mode == VL_POOL_LAST) {
        out = x[(hi - 1) * d + col];
      } else if (mode == VL_POOL_MAX) {
        out = x[lo * d + col];
        for (long long r = lo + 1; r < hi; ++r) out = fmaxf(out, x[r * d + col]);
      } else {
        float acc = x[lo * d + col];
        for (long long r = lo + 1; r < hi; ++r) acc = __fadd_rn(acc, x[r * d + col]);
        out = __fdiv_rn(acc, (float)(hi - lo));
      }
    }
    if (y) y[idx] = out;
    if (y_bf16) y_bf16[idx] = __float2bfloat16_rn(out);
  }
}

__global__ void segment_pool_bwd_kernel(const float* __restrict__ dy, const int32_t* __restrict__ seg, int fixed_len,
                                        int num_seg, int d, int mode, float* __restrict__ dx) {
  const int s = blockIdx.x;
  const long long lo = seg ? seg[s] : (long long)s * fixed_len;
  const long long hi = seg ? seg[s + 1] : lo + fixed_len;
  const float inv = hi > lo ? 1.0f / (float)(hi - lo) : 0.f;
  for (long long r = lo; r < hi; ++r) {
    for (int col = threadIdx.x; col < d; col += blockDim.x) {
      float g = dy[(long long)s * d + col];
      dx[r * d + col] = (mode == VL_POOL_LAST) ? (r == hi - 1 ? g : 0.f) : g * inv;
    }
  }
}

// Early frame fusion (model.py:103-108: aggregate_clip_vectors on the dcnn FEATURES): the same pooling over bf16 rows
// (the encoder stores its activations in bf16), fp32 accumulation in row order.
__global__ void segment_pool_fwd_bf16_kernel(const bf16* __restrict__ x, int fixed_len, int num_seg, int d, int mode,
                                             float* __restrict__ y, bf16* __restrict__ y_bf16) {
  const long long total = (long long)num_seg * d;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int s = (int)(idx / d);
    const int col = (int)(idx - (long long)s * d);
    const long long lo = (long long)s * fixed_len, hi = lo + fixed_len;
    float out;
    if (mode == VL_POOL_LAST) {
      out = __bfloat162float(x[(hi - 1) * d + col]);
    } else {
      float acc = __bfloat162float(x[lo * d + col]);
      for (long long r = lo + 1; r < hi; ++r) acc = __fadd_rn(acc, __bfloat162float(x[r * d + col]));
      out = __fdiv_rn(acc, (float)fixed_len);
    }
    if (y) y[idx] = out;
    if (y_bf16) y_bf16[idx] = __float2bfloat16_rn(out);
  }
}

// ... and its gradient fused with the ReLU gradient of the layer that produced the features:
// dx[r][col] = (act[r][col] > 0) ? pool'(dy[s][col]) : 0, stored in bf16 (the operand of the fc filter gradient).
__global__ void segment_pool_bwd_relu_bf16_kernel(const float* __restrict__ dy, const bf16* __restrict__ act,
                                                  int fixed_len, int num_seg, int d, int mode, bf16* __restrict__ dx) {
  const long long total = (long long)num_seg * fixed_len * d;
  const float inv = 1.0f / (float)fixed_len;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx / d;
    const int col = (int)(idx - r * d);
    const int s = (int)(r / fixed_len);
    const int t = (int)(r - (long long)s * fixed_len);
    float g = dy[(long long)s * d + col];
    g = (mode == VL_POOL_LAST) ? (t == fixed_len - 1 ? g : 0.f) : g * inv;
    if (!(__bfloat162float(act[idx]) > 0.f)) g = 0.f;
    dx[idx] = __float2bfloat16_rn(g);
  }
}

// get_embedding_from_logits (lstm.py:257-265): one CTA per row; strided scan, then a block-wide (value, index)
// reduction with the lowest index winning ties (tf.arg_max); the winner's embedding row is copied out.
__global__ void __launch_bounds__(256)
    argmax_gather_kernel(const float* __restrict__ logits, int v, int ld, const float* __restrict__ embedding, int e,
                         long long* __restrict__ index, float* __restrict__ out, bf16* __restrict__ out_bf16) {
  __shared__ float sval[8];
  __shared__ int sidx[8];
  __shared__ int winner;
  const int r = blockIdx.x;
  const float* z = logits + (long long)r * ld;
  float m = -INFINITY;
  int mi = 0x7fffffff;
  for (int k = threadIdx.x; k < v; k += blockDim.x) {
    const float x = z[k];
    if (x > m) {  // strictly greater: within a thread indices grow, so the first maximum is kept
      m = x;
      mi = k;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o);
    const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
    if (om > m || (om == m && oi < mi)) {
      m = om;
      mi = oi;
    }
  }
  if ((threadIdx.x & 31) == 0) {
    sval[threadIdx.x >> 5] = m;
    sidx[threadIdx.x >> 5] = mi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float bm = sval[0];
    int bi = sidx[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
      if (sval[w] > bm || (sval[w] == bm && sidx[w] < bi)) {
        bm = sval[w];
        bi = sidx[w];
      }
    winner = bi;
    if (index != nullptr) index[r] = bi;
  }
  __syncthreads();
  if (embedding != nullptr) {
    const float* src = embedding + (long long)winner * e;
    for (int k = threadIdx.x; k < e; k += blockDim.x) {
      const float x = src[k];
      if (out != nullptr) out[(long long)r * e + k] = x;
      if (out_bf16 != nullptr) out_bf16[(long long)r * e + k] = __float2bfloat16_rn(x);
    }
  }
}

// apply_tensor_list_fusion (tf_util.py:136-147): element-wise mean / maximum over a list of k same-shaped tensors
// (tf.reduce_mean / tf.reduce_max over axis 0 of the stacked list).  mean = sum in list order, then one division.
struct FuseList {
  const float* src[8];
  int k;
};
__global__ void fuse_list_kernel(const FuseList in, long long n, int mode, float* __restrict__ y,
                                 bf16* __restrict__ y_bf16) {
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
       idx += (long long)gridDim.x * blockDim.x) {
    float acc = in.src[0][idx];
    for (int j = 1; j < in.k; ++j) {
      const float v = in.src[j][idx];
      acc = mode == VL_POOL_MAX ? fmaxf(acc, v) : __fadd_rn(acc, v);
    }
    if (mode != VL_POOL_MAX) acc = __fdiv_rn(acc, (float)in.k);
    if (y) y[idx] = acc;
    if (y_bf16) y_bf16[idx] = __float2bfloat16_rn(acc);
  }
}

// ---- Philox4x32-10 ----
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}

__global__ void dropout_mask_kernel(float* __restrict__ mask, long long n, float keep, uint64_t seed, uint64_t offset) {
  const float inv = 1.0f / keep;
  const long long n4 = (n + 3) >> 2;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n4;
       idx += (long long)gridDim.x * blockDim.x) {
    const uint64_t cnt = offset + (uint64_t)idx;
    uint4 r = philox4x32_10(make_uint4((uint32_t)cnt, (uint32_t)(cnt >> 32), 0u, 0u),
                            make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long e = idx * 4 + k;
      if (e < n) {
        const float u = (float)(rr[k] >> 8) * (1.0f / 16777216.0f);  // [0,1)
        mask[e] = u < keep ? inv : 0.f;
      }
    }
  }
}

__global__ void mul_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y,
                           bf16* __restrict__ y_bf16, long long n) {
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
       idx += (long long)gridDim.x * blockDim.x) {
    const float v = a[idx] * (b ? b[idx] : 1.0f);
    if (y) y[idx] = v;
    if (y_bf16) y_bf16[idx] = __float2bfloat16_rn(v);
  }
}

// one warp per row
__global__ void softmax_ce_kernel(const float* __restrict__ logits, const int32_t* __restrict__ labels, int rows, int c,
                                  float grad_scale, float* __restrict__ row_loss, float* __restrict__ row_correct,
                                  float* __restrict__ dlogits, bf16* __restrict__ dlogits_bf16, int dl_ld) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float* z = logits + (long long)warp * c;
  const int32_t* yl = labels + (long long)warp * c;
  float m = -INFINITY;
  int zi = 0x7fffffff;   // argmax index of logits (lowest index on ties)
  int lbest = -1, li = 0x7fffffff;  // argmax of labels
  for (int k = lane; k < c; k += 32) {
    const float v = z[k];
    if (v > m) {
      m = v;
      zi = k;
    }
    const int l = yl[k];
    if (l > lbest) {
      lbest = l;
      li = k;
    }
  }
  // warp argmax with lowest-index tie break
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o);
    const int oi = __shfl_xor_sync(0xffffffffu, zi, o);
    if (om > m || (om == m && oi < zi)) {
      m = om;
      zi = oi;
    }
    const int ol = __shfl_xor_sync(0xffffffffu, lbest, o);
    const int oli = __shfl_xor_sync(0xffffffffu, li, o);
    if (ol > lbest || (ol == lbest && oli < li)) {
      lbest = ol;
      li = oli;
    }
  }
  float se = 0.f, ysum = 0.f, yz = 0.f;
  for (int k = lane; k < c; k += 32) {
    const float v = z[k] - m;
    se += expf(v);
    const float y = (float)yl[k];
    ysum += y;
    yz += y * v;
  }
  se = warp_sum(se);
  ysum = warp_sum(ysum);
  yz = warp_sum(yz);
  const float lse = logf(se);
  // -sum_c y_c * ((z_c - m) - lse)
  const float loss = ysum * lse - yz;
  if (lane == 0) {
    row_loss[warp] = loss;
    row_correct[warp] = (zi == li) ? 1.f : 0.f;
  }
  if (dlogits != nullptr || dlogits_bf16 != nullptr) {
    for (int k = lane; k < dl_ld; k += 32) {
      float g = 0.f;
      if (k < c) g = (expf(z[k] - m - lse) * ysum - (float)yl[k]) * grad_scale;
      if (dlogits) dlogits[(long long)warp * dl_ld + k] = g;
      if (dlogits_bf16) dlogits_bf16[(long long)warp * dl_ld + k] = __float2bfloat16_rn(g);
    }
  }
}

// deterministic in-order sums of the per-row results
// out[0] = loss_scale * sum of the row losses: with loss_scale = 1 / (clips of the GLOBAL batch) the sum of out[0] over
// the data-parallel ranks is the mean loss of the global batch (train.py:123), whatever the shard sizes are
__global__ void ce_finalize_kernel(const float* __restrict__ row_loss, const float* __restrict__ row_correct, int rows,
                                   float loss_scale, float* __restrict__ out_scalars) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float l = 0.f, cor = 0.f;
    for (int r = 0; r < rows; ++r) {
      l += row_loss[r];
      cor += row_correct[r];
    }
    out_scalars[0] = l * loss_scale;
    out_scalars[1] = cor;
  }
}

int sweep_grid(long long work, int block) {
  long long g = (work + block - 1) / block;
  long long cap = (long long)vl::num_sms() * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace

#define VL_LAUNCHED()                  \
  do {                                 \
    vl::g_launches.fetch_add(1);       \
    VL_CHECK_CUDA(cudaGetLastError()); \
  } while (0)

extern "C" int vl_segment_pool_fwd(const float* x, const int32_t* seg, int32_t fixed_len, int32_t num_seg, int32_t d,
                                   int32_t mode, float* y, void* y_bf16, vl_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VL_REQUIRE(x && (y || y_bf16) && num_seg > 0 && d > 0, "vl_segment_pool_fwd: bad arguments");
  VL_REQUIRE(mode == VL_POOL_AVG || mode == VL_POOL_LAST || mode == VL_POOL_MAX, "Undefined frame fusion type : %d",
             mode);
  VL_REQUIRE(seg != nullptr || fixed_len > 0, "vl_segment_pool_fwd: need seg offsets or fixed_len");
  segment_pool_fwd_kernel<<<sweep_grid((long long)num_seg * d, 256), 256, 0, stream>>>(
      x, seg, fixed_len, num_seg, d, mode, y, reinterpret_cast<bf16*>(y_bf16));
  VL_LAUNCHED();
  return 0;
}

extern "C" int vl_segment_pool_bwd(const float* dy, const int32_t* seg, int32_t fixed_len, int32_t num_seg, int32_t d,
                                   int32_t mode, float* dx, vl_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VL_REQUIRE(dy && dx && num_seg > 0 && d > 0, "vl_segment_pool_bwd: bad arguments");
  VL_REQUIRE(mode == VL_POOL_AVG || mode == VL_POOL_LAST, "vl_segment_pool_bwd: only avg/last have gradients");
  segment_pool_bwd_kernel<<<num_seg, 256, 0, stream>>>(dy, seg, fixed_len, num_seg, d, mode, dx);
  VL_LAUNCHED();
  return 0;
}

extern "C" int vl_segment_pool_fwd_bf16(const void* x, int32_t fixed_len, int32_t num_seg, int32_t d, int32_t mode,
                                        float* y, void* y_bf16, vl_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VL_REQUIRE(x && (y || y_bf16) && num_seg > 0 && d > 0 && fixed_len > 0, "vl_segment_pool_fwd_bf16: bad arguments");
  VL_REQUIRE(mode == VL_POOL_AVG || mode == VL_POOL_LAST, "Undefined frame fusion type : %d", mode);
  segment_pool_fwd_bf16_kernel<<<sweep_grid((long long)num_seg * d, 256), 256, 0, stream>>>(
      reinterpret_cast<const bf16*>(x), fixed_len, num_seg, d, mode, y, reinterpret_cast<bf16*>(y_bf16));
  VL_LAUNCHED();
  return 0;
}

extern "C" int vl_segment_pool_bwd_relu_bf16(const float* dy, const void* act, int32_t fixed_len, int32_t num_seg,
                                             int32_t d, int32_t mode, void* dx, vl_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VL_REQUIRE(dy && act && dx && num_seg > 0 && d > 0 && fixed_len > 0, "vl_segment_pool_bwd_relu_bf16: bad arguments");
  VL_REQUIRE(mode == VL_POOL_AVG || mode == VL_POOL_LAST, "vl_segment_pool_bwd_relu_bf16: only avg/last have gradients");
  segment_pool_bwd_relu_bf16_kernel<<<sweep_grid((long long)num_seg * fixed_len * d, 256), 256, 0, stream>>>(
      dy, reinterpret_cast<const bf16*>(act), fixed_len, num_seg, d, mode, reinterpret_cast<bf16*>(dx));
  VL_LAUNCHED();
  return 0;
}

extern "C" int vl_fuse_list(const float* const* inputs, int32_t k, int64_t n, int32_t mode, float* y, void* y_bf16,
                            vl_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VL_REQUIRE(inputs && k >= 1 && k <= 8 && n > 0 && (y || y_bf16), "vl_fuse_list: bad arguments (1..8 inputs)");
  VL_REQUIRE(mode == VL_POOL_AVG || mode == VL_POOL_MAX, "Unknown fusion method: [%d]", mode);
  FuseList in;
  memset(&in, 0, sizeof(in));
  in.k = k;
  for (int j = 0; j < k; ++j) {
    VL_REQUIRE(inputs[j] != nullptr, "vl_fuse_list: input %d is NULL", j);
    in.src[j] = inputs[j];
  }
  fuse_list_kernel<<<sweep_grid(n, 256), 256, 0, stream>>>(in, n, mode, y, reinterpret_cast<bf16*>(y_bf16));
  VL_LAUNCHED();
  return 0;
}

extern "C" int vl_argmax_gather(const float* logits, int32_t rows, int32_t v, int32_t ld, const float* embedding,
                                int32_t e, int64_t* index, float* out, void* out_bf16, vl_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VL_REQUIRE(logits && rows > 0 && v > 0 && ld >= v, "vl_argmax_gather: bad arguments");
  VL_REQUIRE(embedding == nullptr || (e > 0 && (out != nullptr || out_bf16 != nullptr)),
             "vl_argmax_gather: an embedding needs an output");
  argmax_gather_kernel<<<rows, 256, 0, stream>>>(logits, v, ld, embedding, e, reinterpret_cast<long long*>(index), out,
                                                 reinterpret_cast<bf16*>(out_bf16));
  VL_LAUNCHED();
  return 0;
}

extern "C" int vl_dropout_mask(float* mask, int64_t n, float keep_prob, uint64_t seed, uint64_t offset,
                               vl_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VL_REQUIRE(mask && n > 0 && keep_prob > 0.f && keep_prob <= 1.f, "vl_dropout_mask: bad arguments");
  dropout_mask_kernel<<<sweep_grid((n + 3) / 4, 256), 256, 0, stream>>>(mask, n, keep_prob, seed, offset);
  VL_LAUNCHED();
  return 0;
}

extern "C" int vl_mul(const float* a, const float* b, float* y, void* y_bf16, int64_t n, vl_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VL_REQUIRE(a && (y || y_bf16) && n > 0, "vl_mul: bad arguments");
  mul_kernel<<<sweep_grid(n, 256), 256, 0, stream>>>(a, b, y, reinterpret_cast<bf16*>(y_bf16), n);
  VL_LAUNCHED();
  return 0;
}

extern "C" int vl_softmax_ce(const float* logits, const int32_t* labels, int32_t rows, int32_t c, float grad_scale,
                             float* row_loss, float* out_scalars, float* dlogits, void* dlogits_bf16, int32_t dl_ld,
                             vl_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VL_REQUIRE(logits && labels && row_loss && out_scalars && rows > 0 && c > 0, "vl_softmax_ce: bad arguments");
  VL_REQUIRE((dlogits == nullptr && dlogits_bf16 == nullptr) || dl_ld >= c, "vl_softmax_ce: dl_ld must be >= c");
  // row_loss holds [rows] losses followed by [rows] correct flags
  const int warps_per_block = 4;
  const int grid = (rows + warps_per_block - 1) / warps_per_block;
  softmax_ce_kernel<<<grid, warps_per_block * 32, 0, stream>>>(logits, labels, rows, c, grad_scale, row_loss,
                                                               row_loss + rows, dlogits,
                                                               reinterpret_cast<bf16*>(dlogits_bf16), dl_ld);
  VL_LAUNCHED();
  ce_finalize_kernel<<<1, 32, 0, stream>>>(row_loss, row_loss + rows, rows, grad_scale, out_scalars);
  VL_LAUNCHED();
  return 0;
}
