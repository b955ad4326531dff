// Optimiser step of Train.create_single_tier_learning (train.py:199-222) on one flat fp32 arena:
// per-variable squared norms -> global-norm clip scale (kept on device) -> SGD / Adam apply.
// HBM-bound: SGD reads w,g and writes w (12 B/param); Adam reads w,g,m,v and writes w,m,v (28 B/param).
#include "common.cuh"
#include <cstring>
#include "../../include/vlb200.h"

#include <atomic>

namespace vl {
extern std::atomic<long long> g_launches;
}

namespace {

constexpr int SQ_THREADS = 256;
constexpr long long SQ_CHUNK = 1 << 14;  // elements per block (64 per thread: the 9 MB conv slice still fills the GPU)

// Deterministic two-stage reduction (no floating-point atomics): every data-parallel rank must derive the SAME clip
// scale from the same all-reduced gradients, or the replicas drift apart by an ulp per step.
// stage 1: block b owns elements [b * SQ_CHUNK, (b+1) * SQ_CHUNK) and writes partial[b][v] for every variable v.
__global__ void grad_sqnorms_partial_kernel(const float* __restrict__ g, const int64_t* __restrict__ seg, int num_vars,
                                            float* __restrict__ partial) {
  __shared__ float red[SQ_THREADS / 32];
  const long long c0 = (long long)blockIdx.x * SQ_CHUNK;
  const long long c1 = c0 + SQ_CHUNK;
  for (int v = 0; v < num_vars; ++v) {
    const long long lo = max((long long)seg[v], c0);
    const long long hi = min((long long)seg[v + 1], c1);
    if (hi <= lo) {  // block-uniform
      if (threadIdx.x == 0) partial[(long long)blockIdx.x * num_vars + v] = 0.f;
      continue;
    }
    float acc = 0.f;
    for (long long i = lo + threadIdx.x; i < hi; i += SQ_THREADS) {
      const float x = g[i];
      acc = fmaf(x, x, acc);
    }
    acc = vl::ptx::warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int w = 0; w < SQ_THREADS / 32; ++w) s += red[w];
      partial[(long long)blockIdx.x * num_vars + v] = s;
    }
    __syncthreads();
  }
}

// stage 2: block v sums partial[0..nblocks)[v] in a fixed order (strided per thread, then a shared-memory tree).
__global__ void grad_sqnorms_final_kernel(const float* __restrict__ partial, int nblocks, int num_vars,
                                          float* __restrict__ sq) {
  __shared__ float red[SQ_THREADS];
  const int v = blockIdx.x;
  float acc = 0.f;
  for (int b = threadIdx.x; b < nblocks; b += SQ_THREADS) acc += partial[(long long)b * num_vars + v];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = SQ_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) sq[v] = red[0];
}

__global__ void clip_scalars_kernel(const float* __restrict__ sq, int num_vars, float clip, float prescale,
                                    float* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double tot = 0.0;
  for (int v = 0; v < num_vars; ++v) tot += (double)sq[v];
  const float gn = (float)sqrt(tot) * prescale;
  const float scale = clip > 0.f ? clip / fmaxf(gn, clip) : 1.0f;
  double mean = 0.0;
  for (int v = 0; v < num_vars; ++v) mean += sqrt((double)sq[v]);
  out[0] = gn;
  out[1] = scale;
  out[2] = (float)(mean / num_vars) * prescale * scale;
}

__global__ void sgd_kernel(float* __restrict__ w, const float* __restrict__ g, long long n, float lr,
                           const float* __restrict__ scalars, float prescale) {
  const float s = lr * prescale * (scalars ? scalars[1] : 1.0f);
  const long long n4 = n >> 2;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n4;
       idx += (long long)gridDim.x * blockDim.x) {
    float4 wv = reinterpret_cast<float4*>(w)[idx];
    const float4 gv = reinterpret_cast<const float4*>(g)[idx];
    wv.x -= s * gv.x;
    wv.y -= s * gv.y;
    wv.z -= s * gv.z;
    wv.w -= s * gv.w;
    reinterpret_cast<float4*>(w)[idx] = wv;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long long i = n4 << 2; i < n; ++i) w[i] -= s * g[i];
}

// SGD step that also refreshes the bf16 tensor-core operand ("shadow") of up to 4 arena ranges whose shadow has the
// layout of the master (fc6 / fc7 / LSTM kernels = 97 % of the parameters): the updated weights are rounded and
// stored while they are in registers instead of being re-read by a cast kernel.  Ranges are float4 aligned.
struct ShadowSegs {
  long long begin4[4], end4[4];  // [begin, end) in float4 units
  __nv_bfloat16* dst[4];
  int n;
};
__global__ void sgd_shadow_kernel(float* __restrict__ w, const float* __restrict__ g, long long n, float lr,
                                  const float* __restrict__ scalars, float prescale, const ShadowSegs segs) {
  const float s = lr * prescale * (scalars ? scalars[1] : 1.0f);
  const long long n4 = n >> 2;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n4;
       idx += (long long)gridDim.x * blockDim.x) {
    float4 wv = reinterpret_cast<float4*>(w)[idx];
    const float4 gv = reinterpret_cast<const float4*>(g)[idx];
    wv.x -= s * gv.x;
    wv.y -= s * gv.y;
    wv.z -= s * gv.z;
    wv.w -= s * gv.w;
    reinterpret_cast<float4*>(w)[idx] = wv;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (k < segs.n && idx >= segs.begin4[k] && idx < segs.end4[k]) {
        const __nv_bfloat162 lo = __floats2bfloat162_rn(wv.x, wv.y), hi = __floats2bfloat162_rn(wv.z, wv.w);
        uint2 o;
        o.x = *reinterpret_cast<const uint32_t*>(&lo);
        o.y = *reinterpret_cast<const uint32_t*>(&hi);
        reinterpret_cast<uint2*>(segs.dst[k])[idx - segs.begin4[k]] = o;
      }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long long i = n4 << 2; i < n; ++i) w[i] -= s * g[i];
}

__global__ void adam_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long n, float lr_t, float b1, float b2, float eps,
                            const float* __restrict__ scalars, float prescale) {
  const float gs = prescale * (scalars ? scalars[1] : 1.0f);
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
       idx += (long long)gridDim.x * blockDim.x) {
    const float gi = g[idx] * gs;
    const float mi = b1 * m[idx] + (1.f - b1) * gi;
    const float vi = b2 * v[idx] + (1.f - b2) * gi * gi;
    m[idx] = mi;
    v[idx] = vi;
    w[idx] -= lr_t * mi / (sqrtf(vi) + eps);
  }
}

int sweep_grid(long long work, int block) {
  long long g = (work + block - 1) / block;
  long long cap = (long long)vl::num_sms() * 16;
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

extern "C" int64_t vl_grad_sqnorms_workspace(int64_t n, int32_t num_vars) {
  return ((n + SQ_CHUNK - 1) / SQ_CHUNK) * (int64_t)num_vars;
}

extern "C" int vl_grad_sqnorms(const float* grads, int64_t n, const int64_t* seg_offsets, int32_t num_vars,
                               float* sqnorms, float* workspace, int64_t workspace_floats, vl_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VL_REQUIRE(grads && seg_offsets && sqnorms && workspace && num_vars > 0 && n > 0, "vl_grad_sqnorms: bad arguments");
  const int grid = (int)((n + SQ_CHUNK - 1) / SQ_CHUNK);
  VL_REQUIRE(workspace_floats >= (int64_t)grid * num_vars,
             "vl_grad_sqnorms: workspace of %lld floats, %lld needed (vl_grad_sqnorms_workspace)",
             (long long)workspace_floats, (long long)grid * num_vars);
  grad_sqnorms_partial_kernel<<<grid, SQ_THREADS, 0, stream>>>(grads, seg_offsets, num_vars, workspace);
  VL_LAUNCHED();
  grad_sqnorms_final_kernel<<<num_vars, SQ_THREADS, 0, stream>>>(workspace, grid, num_vars, sqnorms);
  VL_LAUNCHED();
  return 0;
}

extern "C" int vl_clip_scalars(const float* sqnorms, int32_t num_vars, float clip_norm, float grad_prescale,
                               float* scalars, vl_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VL_REQUIRE(sqnorms && scalars && num_vars > 0, "vl_clip_scalars: bad arguments");
  clip_scalars_kernel<<<1, 32, 0, stream>>>(sqnorms, num_vars, clip_norm, grad_prescale, scalars);
  VL_LAUNCHED();
  return 0;
}

extern "C" int vl_sgd_update(float* params, const float* grads, int64_t n, float lr, const float* scalars,
                             float grad_prescale, vl_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VL_REQUIRE(params && grads && n > 0, "vl_sgd_update: bad arguments");
  sgd_kernel<<<sweep_grid(n / 4 + 1, 256), 256, 0, stream>>>(params, grads, n, lr, scalars, grad_prescale);
  VL_LAUNCHED();
  return 0;
}

extern "C" int vl_adam_update(float* params, const float* grads, float* m, float* v, int64_t n, float lr, float beta1,
                              float beta2, float eps, int32_t step, const float* scalars, float grad_prescale,
                              vl_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VL_REQUIRE(params && grads && m && v && n > 0 && step >= 1, "vl_adam_update: bad arguments");
  const double lr_t = (double)lr * sqrt(1.0 - pow((double)beta2, step)) / (1.0 - pow((double)beta1, step));
  adam_kernel<<<sweep_grid(n, 256), 256, 0, stream>>>(params, grads, m, v, n, (float)lr_t, beta1, beta2, eps, scalars,
                                                      grad_prescale);
  VL_LAUNCHED();
  return 0;
}

extern "C" int vl_sgd_update_shadow(float* params, const float* grads, int64_t n, float lr, const float* scalars,
                                    float grad_prescale, int32_t num_segs, const int64_t* seg_begin,
                                    const int64_t* seg_end, void* const* seg_dst, vl_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VL_REQUIRE(params && grads && n > 0 && num_segs >= 0 && num_segs <= 4, "vl_sgd_update_shadow: bad arguments");
  ShadowSegs segs;
  memset(&segs, 0, sizeof(segs));
  segs.n = num_segs;
  for (int k = 0; k < num_segs; ++k) {
    VL_REQUIRE(seg_begin[k] % 4 == 0 && seg_end[k] % 4 == 0 && seg_begin[k] <= seg_end[k] && seg_end[k] <= (n & ~3LL) &&
                   (reinterpret_cast<uintptr_t>(seg_dst[k]) & 7) == 0,
               "vl_sgd_update_shadow: range %d must be float4 aligned inside the arena, its shadow 8-byte aligned", k);
    segs.begin4[k] = seg_begin[k] / 4;
    segs.end4[k] = seg_end[k] / 4;
    segs.dst[k] = reinterpret_cast<__nv_bfloat16*>(seg_dst[k]);
  }
  sgd_shadow_kernel<<<sweep_grid(n / 4 + 1, 256), 256, 0, stream>>>(params, grads, n, lr, scalars, grad_prescale, segs);
  VL_LAUNCHED();
  return 0;
}
