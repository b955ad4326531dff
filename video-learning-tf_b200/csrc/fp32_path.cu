// fp32-accuracy forward mode (sm_100a): the tensor cores take bf16 operands, so every fp32 tensor X of the forward pass is
// split into X_hi = bf16(X) and X_lo = bf16(X - X_hi), and a product X * W is evaluated as
//     X_hi * W_hi + X_lo * W_hi + X_hi * W_lo        (the dropped X_lo * W_lo term is <= 2^-18 relative)
// in ONE launch of the tcgen05 contraction kernel by concatenating along the contraction axis:
//     [X_hi | X_lo | X_hi] (3K columns)  x  [W_hi ; W_hi ; W_lo] (3K rows),   fp32 accumulation in TMEM, fp32 output.
// This file holds the element-wise helpers of that mode; the contractions are vl_gemm (csrc/gemm_umma.cu).  It exists for
// the north-star tolerance "per-frame logits and losses <= 1e-3 relative in fp32" against the reference's fp32 TensorFlow
// graph (models/alexnet/alexnet.py:60-280, models/lstm/lstm.py:59-143); the bf16 path stays the fast one.
#include "common.cuh"
#include "../../include/vlb200.h"

#include <atomic>

namespace vl {
extern std::atomic<long long> g_launches;
}

namespace {

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ void split_hi_lo(float v, bf16& hi, bf16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// out[row][g*3*cg + {0, cg, 2cg} + j] = {hi, lo, hi} of x[row][g*cg + j]
__global__ void split3_act_kernel(const float* __restrict__ x, bf16* __restrict__ out, long long total, int c, int cg) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / c;
    const int ch = (int)(i - row * c);
    const int g = ch / cg, j = ch - g * cg;
    bf16 hi, lo;
    split_hi_lo(x[i], hi, lo);
    bf16* o = out + row * 3 * c + (long long)g * 3 * cg + j;
    o[0] = hi;
    o[cg] = lo;
    o[2 * cg] = hi;
  }
}

// out[k][col], out[rows + k][col] = hi(w[k][col]); out[2*rows + k][col] = lo(w[k][col]); zero for col >= cols
__global__ void split3_weight_kernel(const float* __restrict__ w, bf16* __restrict__ out, long long rows, int cols,
                                     int dst_ld) {
  const long long total = rows * dst_ld;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long k = i / dst_ld;
    const int col = (int)(i - k * dst_ld);
    bf16 hi = __float2bfloat16_rn(0.f), lo = hi;
    if (col < cols) split_hi_lo(w[k * cols + col], hi, lo);
    out[i] = hi;
    out[total + i] = hi;
    out[2 * total + i] = lo;
  }
}

// dst[i] = hi / lo part of src[table[i] & 0x3fffffff] (bit 30 selects lo), 0 where table[i] < 0
__global__ void gather_split_kernel(const float* __restrict__ src, const int32_t* __restrict__ table,
                                    bf16* __restrict__ dst, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int32_t t = table[i];
    bf16 hi = __float2bfloat16_rn(0.f), lo = hi;
    if (t >= 0) split_hi_lo(src[t & 0x3fffffff], hi, lo);
    dst[i] = (t >= 0 && (t & 0x40000000)) ? lo : hi;
  }
}

// fp32 twin of frames_s2d_kernel (csrc/encoder_fused.cu): one thread per output element
__global__ void frames_s2d_f32_kernel(const void* __restrict__ frames_, int is_u8, const float* __restrict__ mean3,
                                      float* __restrict__ out, long long total, int hr, int wr,
                                      const int32_t* __restrict__ crops, int h, int w, int s, int pad_top, int pad_left,
                                      int hb, int wb) {
  const int cblk = s * s * 3;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i % cblk);
    long long r = i / cblk;
    const int bx = (int)(r % wb);
    r /= wb;
    const int by = (int)(r % hb);
    const int nn = (int)(r / hb);
    const int c = j % 3, dd = j / 3;
    const int dy = dd / s, dx = dd - dy * s;
    const int y = by * s - pad_top + dy, x = bx * s - pad_left + dx;
    float v = 0.f;
    if (y >= 0 && y < h && x >= 0 && x < w) {
      int y0 = 0, x0 = 0, mirror = 0;
      if (crops != nullptr) {
        y0 = max(0, min(crops[nn * 3], hr - h));
        x0 = max(0, min(crops[nn * 3 + 1], wr - w));
        mirror = crops[nn * 3 + 2] != 0;
      }
      const int xs = x0 + (mirror ? (w - 1 - x) : x);
      const long long off = (((long long)nn * hr + (y0 + y)) * wr + xs) * 3 + c;
      if (is_u8)
        v = (float)reinterpret_cast<const uint8_t*>(frames_)[off] - (mean3 != nullptr ? mean3[c] : 0.f);
      else
        v = reinterpret_cast<const float*>(frames_)[off];
    }
    out[i] = v;
  }
}

// y[n][p][q][c] = max over the 3x3 stride-2 VALID window of lrn(x) (tf.nn.lrn: x / (bias + alpha * sum_{|d|<=radius} x^2)^beta)
// or of x itself (with_lrn == 0); one thread per output element, fp32 throughout (powf)
__global__ void lrn_pool_f32_kernel(const float* __restrict__ x, float* __restrict__ y, long long total, int h, int w, int c,
                                    int p, int q, int radius, float alpha, float beta, float bias, int with_lrn) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c);
    long long r = i / c;
    const int qq = (int)(r % q);
    r /= q;
    const int pp = (int)(r % p);
    const long long nn = r / p;
    float best = -INFINITY;
    for (int dy = 0; dy < 3; ++dy)
      for (int dx = 0; dx < 3; ++dx) {
        const float* px = x + ((nn * h + (2 * pp + dy)) * w + (2 * qq + dx)) * c;
        float v = px[ch];
        if (with_lrn) {
          float ss = 0.f;
          const int lo = max(0, ch - radius), hi = min(c - 1, ch + radius);
          for (int d = lo; d <= hi; ++d) ss += px[d] * px[d];
          v = v * powf(bias + alpha * ss, -beta);
        }
        best = fmaxf(best, v);
      }
    y[i] = best;
  }
}

int grid_for(long long total) {
  long long g = (total + 255) / 256;
  const long long cap = (long long)vl::num_sms() * 16;
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

}  // namespace

#define VL_LAUNCHED()                  \
  do {                                 \
    vl::g_launches.fetch_add(1);       \
    VL_CHECK_CUDA(cudaGetLastError()); \
  } while (0)

extern "C" int vl_split3_act(const float* x, void* out, int64_t rows, int32_t c, int32_t groups, vl_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VL_REQUIRE(x && out && rows > 0 && c > 0 && groups >= 1 && c % groups == 0, "vl_split3_act: bad arguments");
  const long long total = (long long)rows * c;
  split3_act_kernel<<<grid_for(total), 256, 0, stream>>>(x, reinterpret_cast<bf16*>(out), total, c, c / groups);
  VL_LAUNCHED();
  return 0;
}

extern "C" int vl_split3_weight(const float* w, void* out, int64_t rows, int32_t cols, int32_t dst_ld, vl_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VL_REQUIRE(w && out && rows > 0 && cols > 0 && dst_ld >= cols, "vl_split3_weight: bad arguments");
  split3_weight_kernel<<<grid_for((long long)rows * dst_ld), 256, 0, stream>>>(w, reinterpret_cast<bf16*>(out), rows, cols,
                                                                              dst_ld);
  VL_LAUNCHED();
  return 0;
}

extern "C" int vl_gather_split_bf16(const float* src, const int32_t* table, void* dst, int64_t n, vl_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VL_REQUIRE(src && table && dst && n > 0, "vl_gather_split_bf16: bad arguments");
  gather_split_kernel<<<grid_for(n), 256, 0, stream>>>(src, table, reinterpret_cast<bf16*>(dst), n);
  VL_LAUNCHED();
  return 0;
}

extern "C" int vl_frames_s2d_f32(const void* frames, int32_t is_u8, const float* mean3, float* out, int32_t n, int32_t hr,
                                 int32_t wr, const int32_t* crops, int32_t h, int32_t w, int32_t s, int32_t pad_top,
                                 int32_t pad_left, int32_t hb, int32_t wb, vl_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VL_REQUIRE(frames && out && n > 0 && hb > 0 && wb > 0 && s >= 1, "vl_frames_s2d_f32: bad arguments");
  VL_REQUIRE(hr >= h && wr >= w, "vl_frames_s2d_f32: stored frame %dx%d smaller than the network input %dx%d", hr, wr, h, w);
  VL_REQUIRE(crops != nullptr || (hr == h && wr == w), "vl_frames_s2d_f32: crop offsets are required when the stored frame is larger");
  const long long total = (long long)n * hb * wb * s * s * 3;
  frames_s2d_f32_kernel<<<grid_for(total), 256, 0, stream>>>(frames, is_u8, mean3, out, total, hr, wr, crops, h, w, s,
                                                            pad_top, pad_left, hb, wb);
  VL_LAUNCHED();
  return 0;
}

extern "C" int vl_lrn_pool_fwd_f32(const float* x, float* y, int32_t n, int32_t h, int32_t w, int32_t c, int32_t radius,
                                   float alpha, float beta, float bias, int32_t with_lrn, vl_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VL_REQUIRE(x && y && n > 0 && h >= 3 && w >= 3 && c > 0 && radius >= 0, "vl_lrn_pool_fwd_f32: bad arguments");
  const int p = (h - 3) / 2 + 1, q = (w - 3) / 2 + 1;
  const long long total = (long long)n * p * q * c;
  lrn_pool_f32_kernel<<<grid_for(total), 256, 0, stream>>>(x, y, total, h, w, c, p, q, radius, alpha, beta, bias, with_lrn);
  VL_LAUNCHED();
  return 0;
}
