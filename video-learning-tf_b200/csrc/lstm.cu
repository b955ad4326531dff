// Recurrent part of the LSTM temporal model (models/lstm/lstm.py:9-20,102-143).
//
// The input projection x_t * kernel[:D] for ALL timesteps is one tensor-core GEMM (vl_gemm, 94% of the LSTM
// FLOPs).  What is left is the inherently sequential h_{t-1} * kernel[D:] + gate math.  Clips are independent,
// so each CTA owns CB clips and walks the whole sequence without leaving the SM: h lives in shared memory, c in
// registers, and the recurrent weights (fp32 [H][4H], 1 MB at H=256) are streamed from L2 with coalesced loads
// shared by the CB clips of the CTA.  One launch per layer instead of TensorFlow's per-timestep while_loop.
//
// BasicLSTMCell semantics (TF 1.x): gates i, j, f, o = split(g, 4); c' = c*sigmoid(f + forget_bias) +
// sigmoid(i)*tanh(j); h' = tanh(c')*sigmoid(o).
#include "common.cuh"
#include "../../include/vlb200.h"

#include <atomic>

namespace vl {
extern std::atomic<long long> g_launches;
}

namespace {

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

template <int CB>
__global__ void lstm_fwd_kernel(const float* __restrict__ gx, const float* __restrict__ w_h, float* __restrict__ acts,
                                float* __restrict__ cs, float* __restrict__ h_seq, bf16* __restrict__ h_seq_bf16,
                                bf16* __restrict__ h_prev_bf16, int batch, int t_len, int hidden, float forget_bias) {
  extern __shared__ float hs[];  // [2][CB][hidden]
  const int j = threadIdx.x;
  const int b0 = blockIdx.x * CB;
  const int h4 = 4 * hidden;
  float c[CB];
#pragma unroll
  for (int cb = 0; cb < CB; ++cb) {
    c[cb] = 0.f;
    hs[cb * hidden + j] = 0.f;
  }
  __syncthreads();
  for (int t = 0; t < t_len; ++t) {
    const float* hcur = hs + (t & 1) * CB * hidden;
    float* hnext = hs + ((t & 1) ^ 1) * CB * hidden;
    float acc[CB][4];
#pragma unroll
    for (int cb = 0; cb < CB; ++cb) {
      const int b = b0 + cb;
      if (b < batch) {
        const float* g = gx + ((long long)b * t_len + t) * h4;
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[cb][q] = g[q * hidden + j];
        if (h_prev_bf16 != nullptr)
          h_prev_bf16[((long long)b * t_len + t) * hidden + j] = __float2bfloat16_rn(hcur[cb * hidden + j]);
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[cb][q] = 0.f;
      }
    }
    if (t > 0) {
#pragma unroll 4
      for (int k = 0; k < hidden; ++k) {
        const float* wr = w_h + (long long)k * h4 + j;
        const float w0 = __ldg(wr), w1 = __ldg(wr + hidden), w2 = __ldg(wr + 2 * hidden), w3 = __ldg(wr + 3 * hidden);
#pragma unroll
        for (int cb = 0; cb < CB; ++cb) {
          const float hv = hcur[cb * hidden + k];
          acc[cb][0] = fmaf(hv, w0, acc[cb][0]);
          acc[cb][1] = fmaf(hv, w1, acc[cb][1]);
          acc[cb][2] = fmaf(hv, w2, acc[cb][2]);
          acc[cb][3] = fmaf(hv, w3, acc[cb][3]);
        }
      }
    }
#pragma unroll
    for (int cb = 0; cb < CB; ++cb) {
      const int b = b0 + cb;
      const float si = sigmoidf_(acc[cb][0]);
      const float tj = tanhf(acc[cb][1]);
      const float sf = sigmoidf_(acc[cb][2] + forget_bias);
      const float so = sigmoidf_(acc[cb][3]);
      c[cb] = c[cb] * sf + si * tj;
      const float h = tanhf(c[cb]) * so;
      hnext[cb * hidden + j] = h;
      if (b < batch) {
        const long long row = (long long)b * t_len + t;
        if (acts != nullptr) {
          float* a = acts + row * h4;
          a[j] = si;
          a[hidden + j] = tj;
          a[2 * hidden + j] = sf;
          a[3 * hidden + j] = so;
        }
        if (cs != nullptr) cs[row * hidden + j] = c[cb];
        if (h_seq != nullptr) h_seq[row * hidden + j] = h;
        if (h_seq_bf16 != nullptr) h_seq_bf16[row * hidden + j] = __float2bfloat16_rn(h);
      }
    }
    __syncthreads();
  }
}

// One CTA per clip, reverse time.  dgs (this step's gate gradients) is exchanged through shared memory so that
// thread j can form dh_{t-1}[j] = sum_n dg[n] * w_h[j][n] with coalesced reads of w_h^T.
__global__ void lstm_bwd_kernel(const float* __restrict__ dh_seq, const float* __restrict__ acts,
                                const float* __restrict__ cs, const float* __restrict__ w_h_t, bf16* __restrict__ dg,
                                int batch, int t_len, int hidden) {
  extern __shared__ float dgs[];  // [4*hidden]
  const int j = threadIdx.x;
  const int b = blockIdx.x;
  const int h4 = 4 * hidden;
  float dh_next = 0.f, dc_next = 0.f;
  for (int t = t_len - 1; t >= 0; --t) {
    const long long row = (long long)b * t_len + t;
    const float* a = acts + row * h4;
    const float si = a[j], tj = a[hidden + j], sf = a[2 * hidden + j], so = a[3 * hidden + j];
    const float ct = cs[row * hidden + j];
    const float cprev = t > 0 ? cs[(row - 1) * hidden + j] : 0.f;
    const float tc = tanhf(ct);
    const float dh = dh_seq[row * hidden + j] + dh_next;
    const float d_o = dh * tc * so * (1.f - so);
    const float dc = dh * so * (1.f - tc * tc) + dc_next;
    const float d_i = dc * tj * si * (1.f - si);
    const float d_j = dc * si * (1.f - tj * tj);
    const float d_f = dc * cprev * sf * (1.f - sf);
    dc_next = dc * sf;
    dgs[j] = d_i;
    dgs[hidden + j] = d_j;
    dgs[2 * hidden + j] = d_f;
    dgs[3 * hidden + j] = d_o;
    bf16* out = dg + row * h4;
    out[j] = __float2bfloat16_rn(d_i);
    out[hidden + j] = __float2bfloat16_rn(d_j);
    out[2 * hidden + j] = __float2bfloat16_rn(d_f);
    out[3 * hidden + j] = __float2bfloat16_rn(d_o);
    __syncthreads();
    float acc = 0.f;
    if (t > 0) {
#pragma unroll 8
      for (int n = 0; n < h4; ++n) acc = fmaf(dgs[n], __ldg(w_h_t + (long long)n * hidden + j), acc);
    }
    dh_next = acc;
    __syncthreads();
  }
}

}  // namespace

extern "C" int vl_lstm_fwd(const float* gx, const float* w_h, float* acts, float* cs, float* h_seq, void* h_seq_bf16,
                           void* h_prev_bf16, int32_t batch, int32_t t_len, int32_t hidden, float forget_bias,
                           vl_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VL_REQUIRE(gx && w_h && batch > 0 && t_len > 0, "vl_lstm_fwd: bad arguments");
  VL_REQUIRE(hidden % 32 == 0 && hidden <= 1024, "vl_lstm_fwd: hidden must be a multiple of 32 and <= 1024");
  const int sms = vl::num_sms();
  int cb = 1;
  if (batch > sms * 4) cb = 8;
  else if (batch > sms * 2) cb = 4;
  else if (batch > sms) cb = 2;
  const int grid = (batch + cb - 1) / cb;
  const size_t smem = (size_t)2 * cb * hidden * sizeof(float);
  bf16* hb = reinterpret_cast<bf16*>(h_seq_bf16);
  bf16* hp = reinterpret_cast<bf16*>(h_prev_bf16);
#define VL_LSTM_LAUNCH(CB)                                                                                          \
  lstm_fwd_kernel<CB><<<grid, hidden, smem, stream>>>(gx, w_h, acts, cs, h_seq, hb, hp, batch, t_len, hidden, \
                                                      forget_bias)
  if (cb == 1) VL_LSTM_LAUNCH(1);
  else if (cb == 2) VL_LSTM_LAUNCH(2);
  else if (cb == 4) VL_LSTM_LAUNCH(4);
  else VL_LSTM_LAUNCH(8);
#undef VL_LSTM_LAUNCH
  vl::g_launches.fetch_add(1);
  VL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int vl_lstm_bwd(const float* dh_seq, const float* acts, const float* cs, const float* w_h_t, void* dg,
                           int32_t batch, int32_t t_len, int32_t hidden, vl_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VL_REQUIRE(dh_seq && acts && cs && w_h_t && dg && batch > 0 && t_len > 0, "vl_lstm_bwd: bad arguments");
  VL_REQUIRE(hidden % 32 == 0 && hidden <= 1024, "vl_lstm_bwd: hidden must be a multiple of 32 and <= 1024");
  const size_t smem = (size_t)4 * hidden * sizeof(float);
  lstm_bwd_kernel<<<batch, hidden, smem, stream>>>(dh_seq, acts, cs, w_h_t, reinterpret_cast<bf16*>(dg), batch, t_len,
                                                   hidden);
  vl::g_launches.fetch_add(1);
  VL_CHECK_CUDA(cudaGetLastError());
  return 0;
}
