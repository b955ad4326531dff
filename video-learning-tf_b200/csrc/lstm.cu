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

// Optional (captioning variants, lstm.py:102-143): h0 / c0 = initial state [batch][hidden] (zero when NULL),
// lengths[b] = valid timesteps of sequence b (dynamic_rnn's sequence_length: beyond it the state is carried
// through unchanged and the emitted output is zero), h_last / c_last = final state [batch][hidden].
template <int CB>
__global__ void lstm_fwd_kernel(const float* __restrict__ gx, const float* __restrict__ w_h, float* __restrict__ acts,
                                float* __restrict__ cs, float* __restrict__ h_seq, bf16* __restrict__ h_seq_bf16,
                                bf16* __restrict__ h_prev_bf16, int batch, int t_len, int hidden, float forget_bias,
                                const float* __restrict__ h0, const float* __restrict__ c0,
                                const int32_t* __restrict__ lengths, float* __restrict__ h_last,
                                float* __restrict__ c_last) {
  extern __shared__ float hs[];  // [2][CB][hidden]
  const int j = threadIdx.x;
  const int b0 = blockIdx.x * CB;
  const int h4 = 4 * hidden;
  float c[CB];
  int len[CB];
#pragma unroll
  for (int cb = 0; cb < CB; ++cb) {
    const int b = b0 + cb;
    const bool live = b < batch;
    c[cb] = (live && c0 != nullptr) ? c0[(long long)b * hidden + j] : 0.f;
    hs[cb * hidden + j] = (live && h0 != nullptr) ? h0[(long long)b * hidden + j] : 0.f;
    len[cb] = (live && lengths != nullptr) ? min(lengths[b], t_len) : t_len;
  }
  __syncthreads();
  const bool recur_at_0 = h0 != nullptr;
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
    if (t > 0 || recur_at_0) {
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
      const bool valid = t < len[cb];
      const float si = sigmoidf_(acc[cb][0]);
      const float tj = tanhf(acc[cb][1]);
      const float sf = sigmoidf_(acc[cb][2] + forget_bias);
      const float so = sigmoidf_(acc[cb][3]);
      const float c_new = c[cb] * sf + si * tj;
      const float h_new = tanhf(c_new) * so;
      if (valid) c[cb] = c_new;
      const float h_state = valid ? h_new : hcur[cb * hidden + j];  // beyond the length: state copied through
      const float h_out = valid ? h_new : 0.f;                       // ... and a zero output (dynamic_rnn)
      hnext[cb * hidden + j] = h_state;
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
        if (h_seq != nullptr) h_seq[row * hidden + j] = h_out;
        if (h_seq_bf16 != nullptr) h_seq_bf16[row * hidden + j] = __float2bfloat16_rn(h_out);
        if (t == t_len - 1) {
          if (h_last != nullptr) h_last[(long long)b * hidden + j] = h_state;
          if (c_last != nullptr) c_last[(long long)b * hidden + j] = c[cb];
        }
      }
    }
    __syncthreads();
  }
}

// One CTA per clip, reverse time.  dgs (this step's gate gradients) is exchanged through shared memory so that
// thread j can form dh_{t-1}[j] = sum_n dg[n] * w_h[j][n] with coalesced reads of w_h^T.
// Optional: lengths (steps beyond a sequence's length pass the state gradient through and emit zero gate gradients),
// c0 (initial cell state), dh_last / dc_last (gradient w.r.t. the final state), dh0 / dc0 (gradient w.r.t. the initial
// state, written when non-NULL: the recursion then also runs at t = 0).
__global__ void lstm_bwd_kernel(const float* __restrict__ dh_seq, const float* __restrict__ acts,
                                const float* __restrict__ cs, const float* __restrict__ w_h_t, bf16* __restrict__ dg,
                                int batch, int t_len, int hidden, const int32_t* __restrict__ lengths,
                                const float* __restrict__ c0, const float* __restrict__ dh_last,
                                const float* __restrict__ dc_last, float* __restrict__ dh0, float* __restrict__ dc0) {
  extern __shared__ float dgs[];  // [4*hidden]
  const int j = threadIdx.x;
  const int b = blockIdx.x;
  const int h4 = 4 * hidden;
  const int len = lengths != nullptr ? min(lengths[b], t_len) : t_len;
  float dh_next = dh_last != nullptr ? dh_last[(long long)b * hidden + j] : 0.f;
  float dc_next = dc_last != nullptr ? dc_last[(long long)b * hidden + j] : 0.f;
  const bool want_init = dh0 != nullptr;
  for (int t = t_len - 1; t >= 0; --t) {
    const long long row = (long long)b * t_len + t;
    bf16* out = dg + row * h4;
    if (t >= len) {  // block-uniform: padded step, the state (and its gradient) passes through
      const bf16 z = __float2bfloat16_rn(0.f);
      out[j] = z;
      out[hidden + j] = z;
      out[2 * hidden + j] = z;
      out[3 * hidden + j] = z;
      continue;
    }
    const float* a = acts + row * h4;
    const float si = a[j], tj = a[hidden + j], sf = a[2 * hidden + j], so = a[3 * hidden + j];
    const float ct = cs[row * hidden + j];
    const float cprev = t > 0 ? cs[(row - 1) * hidden + j] : (c0 != nullptr ? c0[(long long)b * hidden + j] : 0.f);
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
    out[j] = __float2bfloat16_rn(d_i);
    out[hidden + j] = __float2bfloat16_rn(d_j);
    out[2 * hidden + j] = __float2bfloat16_rn(d_f);
    out[3 * hidden + j] = __float2bfloat16_rn(d_o);
    __syncthreads();
    float acc = 0.f;
    if (t > 0 || want_init) {
#pragma unroll 8
      for (int n = 0; n < h4; ++n) acc = fmaf(dgs[n], __ldg(w_h_t + (long long)n * hidden + j), acc);
    }
    dh_next = acc;
    __syncthreads();
  }
  if (want_init) {
    dh0[(long long)b * hidden + j] = dh_next;
    if (dc0 != nullptr) dc0[(long long)b * hidden + j] = dc_next;
  }
}

}  // namespace

extern "C" int vl_lstm_fwd_ex(const float* gx, const float* w_h, const float* h0, const float* c0,
                              const int32_t* lengths, float* acts, float* cs, float* h_seq, void* h_seq_bf16,
                              void* h_prev_bf16, float* h_last, float* c_last, int32_t batch, int32_t t_len,
                              int32_t hidden, float forget_bias, vl_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VL_REQUIRE(gx && w_h && batch > 0 && t_len > 0, "vl_lstm_fwd: bad arguments");
  VL_REQUIRE(hidden % 32 == 0 && hidden <= 1024, "vl_lstm_fwd: hidden must be a multiple of 32 and <= 1024");
  VL_REQUIRE((h0 == nullptr) == (c0 == nullptr), "vl_lstm_fwd: h0 and c0 come together (LSTMStateTuple)");
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
  lstm_fwd_kernel<CB><<<grid, hidden, smem, stream>>>(gx, w_h, acts, cs, h_seq, hb, hp, batch, t_len, hidden,       \
                                                      forget_bias, h0, c0, lengths, h_last, c_last)
  if (cb == 1) VL_LSTM_LAUNCH(1);
  else if (cb == 2) VL_LSTM_LAUNCH(2);
  else if (cb == 4) VL_LSTM_LAUNCH(4);
  else VL_LSTM_LAUNCH(8);
#undef VL_LSTM_LAUNCH
  vl::g_launches.fetch_add(1);
  VL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int vl_lstm_fwd(const float* gx, const float* w_h, float* acts, float* cs, float* h_seq, void* h_seq_bf16,
                           void* h_prev_bf16, int32_t batch, int32_t t_len, int32_t hidden, float forget_bias,
                           vl_stream_t stream_) {
  return vl_lstm_fwd_ex(gx, w_h, nullptr, nullptr, nullptr, acts, cs, h_seq, h_seq_bf16, h_prev_bf16, nullptr, nullptr,
                        batch, t_len, hidden, forget_bias, stream_);
}

extern "C" int vl_lstm_bwd_ex(const float* dh_seq, const float* dh_last, const float* dc_last, const float* acts,
                              const float* cs, const float* c0, const float* w_h_t, const int32_t* lengths, void* dg,
                              float* dh0, float* dc0, int32_t batch, int32_t t_len, int32_t hidden,
                              vl_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VL_REQUIRE(dh_seq && acts && cs && w_h_t && dg && batch > 0 && t_len > 0, "vl_lstm_bwd: bad arguments");
  VL_REQUIRE(hidden % 32 == 0 && hidden <= 1024, "vl_lstm_bwd: hidden must be a multiple of 32 and <= 1024");
  const size_t smem = (size_t)4 * hidden * sizeof(float);
  lstm_bwd_kernel<<<batch, hidden, smem, stream>>>(dh_seq, acts, cs, w_h_t, reinterpret_cast<bf16*>(dg), batch, t_len,
                                                   hidden, lengths, c0, dh_last, dc_last, dh0, dc0);
  vl::g_launches.fetch_add(1);
  VL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int vl_lstm_bwd(const float* dh_seq, const float* acts, const float* cs, const float* w_h_t, void* dg,
                           int32_t batch, int32_t t_len, int32_t hidden, vl_stream_t stream_) {
  return vl_lstm_bwd_ex(dh_seq, nullptr, nullptr, acts, cs, nullptr, w_h_t, nullptr, dg, nullptr, nullptr, batch, t_len,
                        hidden, stream_);
}

// =================================================================================================
// Persistent cluster kernels (hidden == 256): the recurrent weights stay resident in shared memory.
//
// A cluster of 8 CTAs owns CB = 8 clips for the whole sequence.  CTA r holds the fp32 slice of kernel[D:] for its
// 32 hidden units (4 gates x 32 = 128 columns, 256 x 129 floats = 132 KB, pitch 129 against bank conflicts) for all
// timesteps.  Per step each CTA forms its 128 gate columns for the 8 clips (K split over two thread halves),
// applies the cell update for its 32 units and publishes h_t to every CTA of the cluster through distributed
// shared memory; one cluster barrier per step replaces the per-timestep launches of a while_loop.
// The backward kernel walks time in reverse with the same residency: gate gradients stay fp32 on chip for the
// dh_{t-1} = dg * W_h^T recursion (partial sums exchanged all-to-all over DSMEM), bf16 copies go to HBM for the
// tensor-core filter/data-gradient GEMMs.
// =================================================================================================
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace {

constexpr int CL = 8;          // CTAs per cluster
constexpr int CH = 256;        // hidden size served by the cluster kernels
constexpr int CU = CH / CL;    // hidden units per CTA (32)
constexpr int CN = 4 * CU;     // gate columns per CTA (128)
// CCB = clips per cluster (template parameter: 8 by default, 4 as a measured alternative)
constexpr int WP = CN + 1;     // padded pitch of the weight slice
constexpr int LSTM_CL_THREADS = 256;

template <int CCB>
struct ClusterSmem {
  float w[CH * WP];            // w[k][g*CU+u] = kernel[D+k][g*CH + rank*CU + u]
  float h[2][CH][CCB];         // h_{t-1} of the 8 clips, k-major, double buffered (fwd) / recv partials (bwd)
  float part[2][CN][CCB];      // K-half partial sums (fwd) / dg of this step [n][cb] in part[0] (bwd)
};

__device__ __forceinline__ void load_w_slice(float* w, const float* __restrict__ w_h, int rank) {
  for (int idx = threadIdx.x; idx < CH * CN; idx += LSTM_CL_THREADS) {
    const int k = idx / CN, col = idx - k * CN;
    const int g = col / CU, u = col - g * CU;
    w[k * WP + col] = w_h[(long long)k * (4 * CH) + g * CH + rank * CU + u];
  }
}

template <int CCB>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(LSTM_CL_THREADS, 1)
    lstm_fwd_cluster_kernel(const float* __restrict__ gx, const float* __restrict__ w_h, float* __restrict__ acts,
                            float* __restrict__ cs, float* __restrict__ h_seq, bf16* __restrict__ h_seq_bf16,
                            bf16* __restrict__ h_prev_bf16, int batch, int t_len, float forget_bias) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  ClusterSmem<CCB>& S = *reinterpret_cast<ClusterSmem<CCB>*>(smem_raw);
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int b0 = (blockIdx.x / CL) * CCB;
  const int tid = threadIdx.x;
  load_w_slice(S.w, w_h, rank);
  for (int idx = tid; idx < CH * CCB; idx += LSTM_CL_THREADS) (&S.h[0][0][0])[idx] = 0.f;
  // matmul role: column `col`, K half `kh`
  const int col = tid & (CN - 1);
  const int khalf = tid >> 7;
  // cell role: unit u, clip cb
  const int u = tid & (CU - 1);
  const int cb = tid >> 5;
  const bool cell = cb < CCB;  // CCB = 4: the upper half of the CTA only takes part in the matmul role
  const int b = b0 + cb;
  const bool live = cell && b < batch;
  const int junit = rank * CU + u;
  float c = 0.f;
  float gxr[4] = {0.f, 0.f, 0.f, 0.f};
  if (live) {
    const float* g = gx + ((long long)b * t_len) * (4 * CH) + junit;
#pragma unroll
    for (int q = 0; q < 4; ++q) gxr[q] = g[q * CH];
  }
  cluster.sync();
  for (int t = 0; t < t_len; ++t) {
    const int cur = t & 1;
    float pre[4] = {gxr[0], gxr[1], gxr[2], gxr[3]};
    if (live && t + 1 < t_len) {  // prefetch the next step's input projection
      const float* g = gx + ((long long)b * t_len + t + 1) * (4 * CH) + junit;
#pragma unroll
      for (int q = 0; q < 4; ++q) gxr[q] = g[q * CH];
    }
    if (t > 0) {
      float acc[CCB];
#pragma unroll
      for (int i = 0; i < CCB; ++i) acc[i] = 0.f;
      const float* hk = &S.h[cur][khalf * (CH / 2)][0];
      const float* wk = S.w + khalf * (CH / 2) * WP + col;
#pragma unroll 4
      for (int k = 0; k < CH / 2; ++k) {
        const float wv = wk[k * WP];
#pragma unroll
        for (int v = 0; v < CCB / 4; ++v) {
          const float4 h4 = *reinterpret_cast<const float4*>(hk + k * CCB + 4 * v);
          acc[4 * v] = fmaf(h4.x, wv, acc[4 * v]);
          acc[4 * v + 1] = fmaf(h4.y, wv, acc[4 * v + 1]);
          acc[4 * v + 2] = fmaf(h4.z, wv, acc[4 * v + 2]);
          acc[4 * v + 3] = fmaf(h4.w, wv, acc[4 * v + 3]);
        }
      }
      float4* dst = reinterpret_cast<float4*>(&S.part[khalf][col][0]);
#pragma unroll
      for (int v = 0; v < CCB / 4; ++v) dst[v] = make_float4(acc[4 * v], acc[4 * v + 1], acc[4 * v + 2], acc[4 * v + 3]);
      __syncthreads();
      if (cell) {
#pragma unroll
        for (int q = 0; q < 4; ++q) pre[q] += S.part[0][q * CU + u][cb] + S.part[1][q * CU + u][cb];
      }
    }
    const float hprev = cell ? S.h[cur][junit][cb] : 0.f;
    const float si = sigmoidf_(pre[0]);
    const float tj = tanhf(pre[1]);
    const float sf = sigmoidf_(pre[2] + forget_bias);
    const float so = sigmoidf_(pre[3]);
    c = c * sf + si * tj;
    const float h = tanhf(c) * so;
    // publish h_t[junit][cb] to every CTA of the cluster (next step's operand)
    if (cell) {
#pragma unroll
      for (int r = 0; r < CL; ++r) {
        float* remote = cluster.map_shared_rank(&S.h[cur ^ 1][0][0], r);
        remote[junit * CCB + cb] = h;
      }
    }
    if (live) {
      const long long row = (long long)b * t_len + t;
      if (acts != nullptr) {
        float* a = acts + row * (4 * CH) + junit;
        a[0] = si;
        a[CH] = tj;
        a[2 * CH] = sf;
        a[3 * CH] = so;
      }
      if (cs != nullptr) cs[row * CH + junit] = c;
      if (h_seq != nullptr) h_seq[row * CH + junit] = h;
      if (h_seq_bf16 != nullptr) h_seq_bf16[row * CH + junit] = __float2bfloat16_rn(h);
      if (h_prev_bf16 != nullptr) h_prev_bf16[row * CH + junit] = __float2bfloat16_rn(hprev);
    }
    cluster.sync();  // h_t visible cluster-wide; also orders the reuse of S.part and S.h[cur]
  }
}

template <int CCB>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(LSTM_CL_THREADS, 1)
    lstm_bwd_cluster_kernel(const float* __restrict__ dh_seq, const float* __restrict__ acts,
                            const float* __restrict__ cs, const float* __restrict__ w_h, bf16* __restrict__ dg,
                            int batch, int t_len) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  ClusterSmem<CCB>& S = *reinterpret_cast<ClusterSmem<CCB>*>(smem_raw);
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int b0 = (blockIdx.x / CL) * CCB;
  const int tid = threadIdx.x;
  load_w_slice(S.w, w_h, rank);
  // recv[parity][src][u][cb] aliases S.h[parity] (CL * CU * CCB = CH * CCB floats)
  for (int idx = tid; idx < 2 * CH * CCB; idx += LSTM_CL_THREADS) (&S.h[0][0][0])[idx] = 0.f;
  const int u = tid & (CU - 1);
  const int cb = tid >> 5;
  const bool cell = cb < CCB;
  const int b = b0 + cb;
  const bool live = cell && b < batch;
  const int junit = rank * CU + u;
  float dc_next = 0.f;
  cluster.sync();
  for (int t = t_len - 1; t >= 0; --t) {
    const int par = t & 1;
    // dh_next = sum over source CTAs of the partials they sent for my units (zero at t = T-1)
    float dh = 0.f;
    if (cell) {
      const float* recv = &S.h[par ^ 1][0][0];
#pragma unroll
      for (int src = 0; src < CL; ++src) dh += recv[(src * CU + u) * CCB + cb];
    }
    float d_i = 0.f, d_j = 0.f, d_f = 0.f, d_o = 0.f;
    if (live) {
      const long long row = (long long)b * t_len + t;
      const float* a = acts + row * (4 * CH) + junit;
      const float si = a[0], tj = a[CH], sf = a[2 * CH], so = a[3 * CH];
      const float ct = cs[row * CH + junit];
      const float cprev = t > 0 ? cs[(row - 1) * CH + junit] : 0.f;
      const float tc = tanhf(ct);
      dh += dh_seq[row * CH + junit];
      d_o = dh * tc * so * (1.f - so);
      const float dc = dh * so * (1.f - tc * tc) + dc_next;
      d_i = dc * tj * si * (1.f - si);
      d_j = dc * si * (1.f - tj * tj);
      d_f = dc * cprev * sf * (1.f - sf);
      dc_next = dc * sf;
      bf16* out = dg + row * (4 * CH) + junit;
      out[0] = __float2bfloat16_rn(d_i);
      out[CH] = __float2bfloat16_rn(d_j);
      out[2 * CH] = __float2bfloat16_rn(d_f);
      out[3 * CH] = __float2bfloat16_rn(d_o);
    }
    if (cell) {
      S.part[0][u][cb] = d_i;
      S.part[0][CU + u][cb] = d_j;
      S.part[0][2 * CU + u][cb] = d_f;
      S.part[0][3 * CU + u][cb] = d_o;
    }
    __syncthreads();
    if (t > 0) {
      // partial dh_{t-1}[k][cb] over my 128 gate columns, thread = k
      const int k = tid;
      float acc[CCB];
#pragma unroll
      for (int i = 0; i < CCB; ++i) acc[i] = 0.f;
      const float* wr = S.w + k * WP;
#pragma unroll 4
      for (int n = 0; n < CN; ++n) {
        const float wv = wr[n];
#pragma unroll
        for (int v = 0; v < CCB / 4; ++v) {
          const float4 g4 = *reinterpret_cast<const float4*>(&S.part[0][n][4 * v]);
          acc[4 * v] = fmaf(g4.x, wv, acc[4 * v]);
          acc[4 * v + 1] = fmaf(g4.y, wv, acc[4 * v + 1]);
          acc[4 * v + 2] = fmaf(g4.z, wv, acc[4 * v + 2]);
          acc[4 * v + 3] = fmaf(g4.w, wv, acc[4 * v + 3]);
        }
      }
      // send to the CTA that owns unit k: recv[par][my rank][k % CU][cb]
      float* remote = cluster.map_shared_rank(&S.h[par][0][0], k / CU) + (rank * CU + (k % CU)) * CCB;
#pragma unroll
      for (int v = 0; v < CCB / 4; ++v)
        reinterpret_cast<float4*>(remote)[v] = make_float4(acc[4 * v], acc[4 * v + 1], acc[4 * v + 2], acc[4 * v + 3]);
    }
    cluster.sync();
  }
}


// ------------------------------------------------------------------------------------------------
// Wide variants (1024 threads per CTA, 8 clips per cluster).  The 256-thread kernels above spend a step in a
// 128-iteration dependent FMA loop with two warps per scheduler (7.3 us per step, latency bound); here the
// contraction axis is split eight ways (forward: K = 256 -> 8 x 32) / four ways (backward: 128 gate columns -> 4 x 32),
// the partial sums meet in shared memory, and eight warps per scheduler hide the LDS / FMA latency.
// ------------------------------------------------------------------------------------------------
constexpr int WIDE_T = 1024;
constexpr int WCB = 8;                 // clips per cluster
constexpr int KPARTS = WIDE_T / CN;    // 8 K slices of 32 (forward)
constexpr int NPARTS = WIDE_T / CH;    // 4 column slices of 32 (backward)

struct WideSmem {
  float w[CH * WP];
  float h[2][CH][WCB];
  float part[KPARTS * CN * WCB];       // fwd: [KPARTS][CN][WCB]; bwd: [NPARTS][CH][WCB] (same size)
  float dg[CN][WCB];                   // bwd: gate gradients of this step
};

__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(WIDE_T, 1)
    lstm_fwd_cluster_wide_kernel(const float* __restrict__ gx, const float* __restrict__ w_h, float* __restrict__ acts,
                                 float* __restrict__ cs, float* __restrict__ h_seq, bf16* __restrict__ h_seq_bf16,
                                 bf16* __restrict__ h_prev_bf16, int batch, int t_len, float forget_bias) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  WideSmem& S = *reinterpret_cast<WideSmem*>(smem_raw);
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int b0 = (blockIdx.x / CL) * WCB;
  const int tid = threadIdx.x;
  for (int idx = tid; idx < CH * CN; idx += WIDE_T) {
    const int k = idx / CN, col = idx - k * CN;
    const int g = col / CU, u = col - g * CU;
    S.w[k * WP + col] = w_h[(long long)k * (4 * CH) + g * CH + rank * CU + u];
  }
  for (int idx = tid; idx < CH * WCB; idx += WIDE_T) (&S.h[0][0][0])[idx] = 0.f;
  const int col = tid & (CN - 1);
  const int kp = tid >> 7;
  const bool cell = tid < CU * WCB;  // first 256 threads: unit u of clip cb
  const int u = tid & (CU - 1);
  const int cb = (tid >> 5) & (WCB - 1);
  const int b = b0 + cb;
  const bool live = cell && b < batch;
  const int junit = rank * CU + u;
  float c = 0.f;
  float gxr[4] = {0.f, 0.f, 0.f, 0.f};
  if (live) {
    const float* g = gx + ((long long)b * t_len) * (4 * CH) + junit;
#pragma unroll
    for (int q = 0; q < 4; ++q) gxr[q] = g[q * CH];
  }
  cluster.sync();
  for (int t = 0; t < t_len; ++t) {
    const int cur = t & 1;
    float pre[4] = {gxr[0], gxr[1], gxr[2], gxr[3]};
    if (live && t + 1 < t_len) {
      const float* g = gx + ((long long)b * t_len + t + 1) * (4 * CH) + junit;
#pragma unroll
      for (int q = 0; q < 4; ++q) gxr[q] = g[q * CH];
    }
    if (t > 0) {
      float acc[WCB];
#pragma unroll
      for (int i = 0; i < WCB; ++i) acc[i] = 0.f;
      constexpr int KS = CH / KPARTS;
      const float* hk = &S.h[cur][kp * KS][0];
      const float* wk = S.w + kp * KS * WP + col;
#pragma unroll 8
      for (int k = 0; k < KS; ++k) {
        const float wv = wk[k * WP];
        const float4 h0 = *reinterpret_cast<const float4*>(hk + k * WCB);
        const float4 h1 = *reinterpret_cast<const float4*>(hk + k * WCB + 4);
        acc[0] = fmaf(h0.x, wv, acc[0]);
        acc[1] = fmaf(h0.y, wv, acc[1]);
        acc[2] = fmaf(h0.z, wv, acc[2]);
        acc[3] = fmaf(h0.w, wv, acc[3]);
        acc[4] = fmaf(h1.x, wv, acc[4]);
        acc[5] = fmaf(h1.y, wv, acc[5]);
        acc[6] = fmaf(h1.z, wv, acc[6]);
        acc[7] = fmaf(h1.w, wv, acc[7]);
      }
      float4* dst = reinterpret_cast<float4*>(S.part + (kp * CN + col) * WCB);
      dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
      dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
      __syncthreads();
      if (cell) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float sum = 0.f;
#pragma unroll
          for (int pp = 0; pp < KPARTS; ++pp) sum += S.part[(pp * CN + q * CU + u) * WCB + cb];
          pre[q] += sum;
        }
      }
    }
    float hprev = 0.f, si = 0.f, tj = 0.f, sf = 0.f, so = 0.f, h = 0.f;
    float* hstage = &S.dg[0][0];  // [CU][WCB]: this CTA's slice of h_t, laid out like the remote h[.][rank*CU + u][cb]
    if (cell) {
      hprev = S.h[cur][junit][cb];
      si = sigmoidf_(pre[0]);
      tj = tanhf(pre[1]);
      sf = sigmoidf_(pre[2] + forget_bias);
      so = sigmoidf_(pre[3]);
      c = c * sf + si * tj;
      h = tanhf(c) * so;
      hstage[u * WCB + cb] = h;
    }
    __syncthreads();
    // publish h_t to every CTA of the cluster: 8 destinations x 64 float4 (one 16-byte DSMEM store per thread of the
    // lower half; 2048 scalar remote stores per step were the dominant cost of the 256-thread kernel)
    if (tid < CL * (CU * WCB / 4)) {
      const int r = tid / (CU * WCB / 4), i = tid % (CU * WCB / 4);
      float* remote = cluster.map_shared_rank(&S.h[cur ^ 1][0][0], r) + rank * (CU * WCB);
      reinterpret_cast<float4*>(remote)[i] = reinterpret_cast<const float4*>(hstage)[i];
    }
    // arrive (release) NOW: only the DSMEM publication above has to be visible cluster-wide.  The global stores of
    // this step follow the arrive, so the barrier does not wait for their acknowledgement.
    cluster.barrier_arrive();
    if (cell) {
      if (live) {
        const long long row = (long long)b * t_len + t;
        if (acts != nullptr) {
          float* a = acts + row * (4 * CH) + junit;
          a[0] = si;
          a[CH] = tj;
          a[2 * CH] = sf;
          a[3 * CH] = so;
        }
        if (cs != nullptr) cs[row * CH + junit] = c;
        if (h_seq != nullptr) h_seq[row * CH + junit] = h;
        if (h_seq_bf16 != nullptr) h_seq_bf16[row * CH + junit] = __float2bfloat16_rn(h);
        if (h_prev_bf16 != nullptr) h_prev_bf16[row * CH + junit] = __float2bfloat16_rn(hprev);
      }
    }
    cluster.barrier_wait();  // h_t visible cluster-wide; also orders the reuse of S.part and S.h[cur]
  }
}

__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(WIDE_T, 1)
    lstm_bwd_cluster_wide_kernel(const float* __restrict__ dh_seq, const float* __restrict__ acts,
                                 const float* __restrict__ cs, const float* __restrict__ w_h, bf16* __restrict__ dg,
                                 int batch, int t_len) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  WideSmem& S = *reinterpret_cast<WideSmem*>(smem_raw);
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int b0 = (blockIdx.x / CL) * WCB;
  const int tid = threadIdx.x;
  for (int idx = tid; idx < CH * CN; idx += WIDE_T) {
    const int k = idx / CN, col = idx - k * CN;
    const int g = col / CU, u = col - g * CU;
    S.w[k * WP + col] = w_h[(long long)k * (4 * CH) + g * CH + rank * CU + u];
  }
  for (int idx = tid; idx < 2 * CH * WCB; idx += WIDE_T) (&S.h[0][0][0])[idx] = 0.f;
  const bool cell = tid < CU * WCB;
  const int u = tid & (CU - 1);
  const int cb = (tid >> 5) & (WCB - 1);
  const int b = b0 + cb;
  const bool live = cell && b < batch;
  const int junit = rank * CU + u;
  const int kk = tid & (CH - 1);  // matmul role: hidden unit k, column slice np
  const int np = tid >> 8;
  float dc_next = 0.f;
  // operands of a step (gate activations, cell states, incoming dh) are fetched one step ahead of their use
  float nx[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  auto fetch = [&](int t) {
    if (live && t >= 0) {
      const long long row = (long long)b * t_len + t;
      const float* a = acts + row * (4 * CH) + junit;
      nx[0] = a[0], nx[1] = a[CH], nx[2] = a[2 * CH], nx[3] = a[3 * CH];
      nx[4] = cs[row * CH + junit];
      nx[5] = t > 0 ? cs[(row - 1) * CH + junit] : 0.f;
      nx[6] = dh_seq[row * CH + junit];
    }
  };
  fetch(t_len - 1);
  cluster.sync();
  for (int t = t_len - 1; t >= 0; --t) {
    const int par = t & 1;
    float d_i = 0.f, d_j = 0.f, d_f = 0.f, d_o = 0.f;
    if (cell) {
      float dh = 0.f;
      const float* recv = &S.h[par ^ 1][0][0];
#pragma unroll
      for (int src = 0; src < CL; ++src) dh += recv[(src * CU + u) * WCB + cb];
      if (live) {
        const float si = nx[0], tj = nx[1], sf = nx[2], so = nx[3];
        const float ct = nx[4];
        const float cprev = nx[5];
        const float tc = tanhf(ct);
        dh += nx[6];
        d_o = dh * tc * so * (1.f - so);
        const float dc = dh * so * (1.f - tc * tc) + dc_next;
        d_i = dc * tj * si * (1.f - si);
        d_j = dc * si * (1.f - tj * tj);
        d_f = dc * cprev * sf * (1.f - sf);
        dc_next = dc * sf;
      }
      fetch(t - 1);
      S.dg[u][cb] = d_i;
      S.dg[CU + u][cb] = d_j;
      S.dg[2 * CU + u][cb] = d_f;
      S.dg[3 * CU + u][cb] = d_o;
    }
    __syncthreads();
    if (t > 0) {
      float acc[WCB];
#pragma unroll
      for (int i = 0; i < WCB; ++i) acc[i] = 0.f;
      constexpr int NS = CN / NPARTS;
      const float* wr = S.w + kk * WP + np * NS;
#pragma unroll 8
      for (int n = 0; n < NS; ++n) {
        const float wv = wr[n];
        const float4 g0 = *reinterpret_cast<const float4*>(&S.dg[np * NS + n][0]);
        const float4 g1 = *reinterpret_cast<const float4*>(&S.dg[np * NS + n][4]);
        acc[0] = fmaf(g0.x, wv, acc[0]);
        acc[1] = fmaf(g0.y, wv, acc[1]);
        acc[2] = fmaf(g0.z, wv, acc[2]);
        acc[3] = fmaf(g0.w, wv, acc[3]);
        acc[4] = fmaf(g1.x, wv, acc[4]);
        acc[5] = fmaf(g1.y, wv, acc[5]);
        acc[6] = fmaf(g1.z, wv, acc[6]);
        acc[7] = fmaf(g1.w, wv, acc[7]);
      }
      float4* dst = reinterpret_cast<float4*>(S.part + (np * CH + kk) * WCB);
      dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
      dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
      __syncthreads();
      if (tid < CH) {
        // dh_{t-1}[k][.] partial over my 128 gate columns -> the CTA that owns unit k: recv[par][my rank][k % CU][cb]
        float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
#pragma unroll
        for (int pp = 0; pp < NPARTS; ++pp) {
          const float4 a0 = *reinterpret_cast<const float4*>(S.part + (pp * CH + tid) * WCB);
          const float4 a1 = *reinterpret_cast<const float4*>(S.part + (pp * CH + tid) * WCB + 4);
          s0.x += a0.x, s0.y += a0.y, s0.z += a0.z, s0.w += a0.w;
          s1.x += a1.x, s1.y += a1.y, s1.z += a1.z, s1.w += a1.w;
        }
        float* remote = cluster.map_shared_rank(&S.h[par][0][0], tid / CU) + (rank * CU + (tid % CU)) * WCB;
        reinterpret_cast<float4*>(remote)[0] = s0;
        reinterpret_cast<float4*>(remote)[1] = s1;
      }
    }
    cluster.barrier_arrive();  // release covers the DSMEM partials; the global stores below are not waited for
    if (live) {
      bf16* out = dg + ((long long)b * t_len + t) * (4 * CH) + junit;
      out[0] = __float2bfloat16_rn(d_i);
      out[CH] = __float2bfloat16_rn(d_j);
      out[2 * CH] = __float2bfloat16_rn(d_f);
      out[3 * CH] = __float2bfloat16_rn(d_o);
    }
    cluster.barrier_wait();
  }
}

}  // namespace

// clips per cluster: 8.  Four clips per cluster (16 clusters for 64 clips) halve the FMA work per step but measured
// SLOWER (forward 149 us against 117 us, backward 95 against 69): the recurrence is bound by the per-step cluster
// barrier / DSMEM exchange, not by arithmetic, and 16 clusters of 8 CTAs are not all co-resident.  VL_LSTM_CCB=4 selects it.
static int lstm_clips_per_cluster(int batch) {
  (void)batch;
  const char* e = getenv("VL_LSTM_CCB");
  if (e && (atoi(e) == 4 || atoi(e) == 8)) return atoi(e);
  return 8;
}

template <int CCB>
static int launch_lstm_fwd_cluster(const float* gx, const float* w_h, float* acts, float* cs, float* h_seq,
                                   void* h_seq_bf16, void* h_prev_bf16, int32_t batch, int32_t t_len,
                                   float forget_bias, cudaStream_t stream) {
  static bool attr = false;
  if (!attr) {
    VL_CHECK_CUDA(cudaFuncSetAttribute(lstm_fwd_cluster_kernel<CCB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(ClusterSmem<CCB>)));
    attr = true;
  }
  const int clusters = (batch + CCB - 1) / CCB;
  lstm_fwd_cluster_kernel<CCB><<<clusters * CL, LSTM_CL_THREADS, sizeof(ClusterSmem<CCB>), stream>>>(
      gx, w_h, acts, cs, h_seq, reinterpret_cast<bf16*>(h_seq_bf16), reinterpret_cast<bf16*>(h_prev_bf16), batch, t_len,
      forget_bias);
  vl::g_launches.fetch_add(1);
  VL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <int CCB>
static int launch_lstm_bwd_cluster(const float* dh_seq, const float* acts, const float* cs, const float* w_h, void* dg,
                                   int32_t batch, int32_t t_len, cudaStream_t stream) {
  static bool attr = false;
  if (!attr) {
    VL_CHECK_CUDA(cudaFuncSetAttribute(lstm_bwd_cluster_kernel<CCB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(ClusterSmem<CCB>)));
    attr = true;
  }
  const int clusters = (batch + CCB - 1) / CCB;
  lstm_bwd_cluster_kernel<CCB><<<clusters * CL, LSTM_CL_THREADS, sizeof(ClusterSmem<CCB>), stream>>>(
      dh_seq, acts, cs, w_h, reinterpret_cast<bf16*>(dg), batch, t_len);
  vl::g_launches.fetch_add(1);
  VL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int vl_lstm_fwd_cluster(const float* gx, const float* w_h, float* acts, float* cs, float* h_seq,
                                   void* h_seq_bf16, void* h_prev_bf16, int32_t batch, int32_t t_len, int32_t hidden,
                                   float forget_bias, vl_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VL_REQUIRE(gx && w_h && batch > 0 && t_len > 0, "vl_lstm_fwd_cluster: bad arguments");
  VL_REQUIRE(hidden == CH, "vl_lstm_fwd_cluster: the resident-weight kernel serves hidden == %d", CH);
  if (!(getenv("VL_LSTM_WIDE") && atoi(getenv("VL_LSTM_WIDE")) == 0) && !getenv("VL_LSTM_CCB")) {
    static bool attr = false;
    if (!attr) {
      VL_CHECK_CUDA(cudaFuncSetAttribute(lstm_fwd_cluster_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(WideSmem)));
      attr = true;
    }
    const int clusters = (batch + WCB - 1) / WCB;
    lstm_fwd_cluster_wide_kernel<<<clusters * CL, WIDE_T, sizeof(WideSmem), stream>>>(
        gx, w_h, acts, cs, h_seq, reinterpret_cast<bf16*>(h_seq_bf16), reinterpret_cast<bf16*>(h_prev_bf16), batch, t_len,
        forget_bias);
    vl::g_launches.fetch_add(1);
    VL_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  if (lstm_clips_per_cluster(batch) == 4)
    return launch_lstm_fwd_cluster<4>(gx, w_h, acts, cs, h_seq, h_seq_bf16, h_prev_bf16, batch, t_len, forget_bias, stream);
  return launch_lstm_fwd_cluster<8>(gx, w_h, acts, cs, h_seq, h_seq_bf16, h_prev_bf16, batch, t_len, forget_bias, stream);
}

extern "C" int vl_lstm_bwd_cluster(const float* dh_seq, const float* acts, const float* cs, const float* w_h, void* dg,
                                   int32_t batch, int32_t t_len, int32_t hidden, vl_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VL_REQUIRE(dh_seq && acts && cs && w_h && dg && batch > 0 && t_len > 0, "vl_lstm_bwd_cluster: bad arguments");
  VL_REQUIRE(hidden == CH, "vl_lstm_bwd_cluster: the resident-weight kernel serves hidden == %d", CH);
  if (!(getenv("VL_LSTM_WIDE") && atoi(getenv("VL_LSTM_WIDE")) == 0) && !getenv("VL_LSTM_CCB")) {
    static bool attr = false;
    if (!attr) {
      VL_CHECK_CUDA(cudaFuncSetAttribute(lstm_bwd_cluster_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(WideSmem)));
      attr = true;
    }
    const int clusters = (batch + WCB - 1) / WCB;
    lstm_bwd_cluster_wide_kernel<<<clusters * CL, WIDE_T, sizeof(WideSmem), stream>>>(
        dh_seq, acts, cs, w_h, reinterpret_cast<bf16*>(dg), batch, t_len);
    vl::g_launches.fetch_add(1);
    VL_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  if (lstm_clips_per_cluster(batch) == 4)
    return launch_lstm_bwd_cluster<4>(dh_seq, acts, cs, w_h, dg, batch, t_len, stream);
  return launch_lstm_bwd_cluster<8>(dh_seq, acts, cs, w_h, dg, batch, t_len, stream);
}
