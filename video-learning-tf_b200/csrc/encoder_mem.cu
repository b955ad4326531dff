// HBM-bound kernels of the AlexNet encoder: conv1 patch staging, LRN, 3x3/2 max-pool, bias gradients and
// the fp32 -> bf16 operand packing.  All loads/stores are 16-byte vectors over the channel axis (NHWC), grids
// are plain 1-D sweeps (these kernels have no reuse to tile for; they are judged on achieved GB/s).
//
// Reference call sites: models/alexnet/alexnet.py:76 (conv1 input), :85-89,:126-130 (LRN), :98,:139,:211
// (max_pool), bias_add gradients of :31, make_w_b (:40-46) variables feeding the tensor-core operands.
#include "common.cuh"
#include "../../include/vlb200.h"

#include <atomic>

namespace vl {
extern std::atomic<long long> g_launches;
}

namespace {

typedef __nv_bfloat16 bf16;

struct alignas(16) Bf16x8 {
  __nv_bfloat162 v[4];
};

// 16-byte accesses are spelled through uint4: a plain copy of the 4 x bf16x2 struct compiles to four 32-bit LDG/STG
// (LDS/STS) instructions, i.e. 4x the LSU wavefronts.
__device__ __forceinline__ Bf16x8 ld8(const void* p) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  Bf16x8 r;
  r.v[0] = *reinterpret_cast<const __nv_bfloat162*>(&u.x);
  r.v[1] = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
  r.v[2] = *reinterpret_cast<const __nv_bfloat162*>(&u.z);
  r.v[3] = *reinterpret_cast<const __nv_bfloat162*>(&u.w);
  return r;
}
__device__ __forceinline__ void st8(void* p, const Bf16x8& v) {
  uint4 u;
  u.x = *reinterpret_cast<const uint32_t*>(&v.v[0]);
  u.y = *reinterpret_cast<const uint32_t*>(&v.v[1]);
  u.z = *reinterpret_cast<const uint32_t*>(&v.v[2]);
  u.w = *reinterpret_cast<const uint32_t*>(&v.v[3]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void unpack8(const Bf16x8& in, float (&f)[8]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(in.v[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ Bf16x8 pack8(const float (&f)[8]) {
  Bf16x8 o;
#pragma unroll
  for (int i = 0; i < 4; ++i) o.v[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return o;
}

// ------------------------------------------------------------------------------------------------
// LRN
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float pow_neg_beta(float s, float beta) {
  if (beta == 0.75f) {
    float r = rsqrtf(s);
    return r * sqrtf(r);
  }
  return __powf(s, -beta);
}

template <int R>
__global__ void lrn_fwd_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, long long rows, int c, int radius_rt,
                               float alpha, float beta, float bias, long long total_chunks) {
  const int radius = R > 0 ? R : radius_rt;  // R = 2 (the reference's value) keeps the window in registers
  const int cpr = c >> 3;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total_chunks;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long row = idx / cpr;
    const int c0 = (int)(idx - row * cpr) * 8;
    const bf16* xr = x + row * c;
    float v[8 + 8];  // window [c0-4, c0+12), only [c0-radius, c0+8+radius) is used (radius <= 4)
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = 0.f;
    {
      float own[8];
      unpack8(ld8(xr + c0), own);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[4 + j] = own[j];
    }
#pragma unroll
    for (int j = 1; j <= 4; ++j) {
      if (j <= radius) {
        if (c0 - j >= 0) v[4 - j] = __bfloat162float(xr[c0 - j]);
        if (c0 + 7 + j < c) v[11 + j] = __bfloat162float(xr[c0 + 7 + j]);
      }
    }
    float out[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float acc = 0.f;
#pragma unroll
      for (int o = -4; o <= 4; ++o)
        if (o >= -radius && o <= radius) acc += v[4 + j + o] * v[4 + j + o];
      const float s = bias + alpha * acc;
      out[j] = v[4 + j] * pow_neg_beta(s, beta);
    }
    st8(y + row * c + c0, pack8(out));
  }
}

// radius is fixed to 2 in the gradient kernel (the only value the reference uses).
__global__ void lrn_bwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, bf16* __restrict__ dx,
                               long long rows, int c, float alpha, float beta, float bias, int relu_mask,
                               long long total_chunks) {
  const int cpr = c >> 3;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total_chunks;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long row = idx / cpr;
    const int c0 = (int)(idx - row * cpr) * 8;
    const bf16* xr = x + row * c;
    const bf16* gr = dy + row * c;
    float xv[16];  // channels [c0-4, c0+12)
    float gv[12];  // channels [c0-2, c0+10)
#pragma unroll
    for (int j = 0; j < 16; ++j) xv[j] = 0.f;
#pragma unroll
    for (int j = 0; j < 12; ++j) gv[j] = 0.f;
    {
      float own[8];
      unpack8(ld8(xr + c0), own);
#pragma unroll
      for (int j = 0; j < 8; ++j) xv[4 + j] = own[j];
      unpack8(ld8(gr + c0), own);
#pragma unroll
      for (int j = 0; j < 8; ++j) gv[2 + j] = own[j];
    }
#pragma unroll
    for (int j = 1; j <= 4; ++j) {
      if (c0 - j >= 0) xv[4 - j] = __bfloat162float(xr[c0 - j]);
      if (c0 + 7 + j < c) xv[11 + j] = __bfloat162float(xr[c0 + 7 + j]);
    }
#pragma unroll
    for (int j = 1; j <= 2; ++j) {
      if (c0 - j >= 0) gv[2 - j] = __bfloat162float(gr[c0 - j]);
      if (c0 + 7 + j < c) gv[9 + j] = __bfloat162float(gr[c0 + 7 + j]);
    }
    // t_d = dy_d * x_d * s_d^(-beta-1) for d in [c0-2, c0+10); s_i^-beta for the 8 own channels
    float t[12];
    float sp[8];
#pragma unroll
    for (int d = 0; d < 12; ++d) {
      // channel c0-2+d sits at xv[2+d]; its window is xv[d .. d+4]
      float acc = 0.f;
#pragma unroll
      for (int o = 0; o < 5; ++o) acc += xv[d + o] * xv[d + o];
      const float s = bias + alpha * acc;
      const float pw = pow_neg_beta(s, beta);
      t[d] = gv[d] * xv[2 + d] * pw / s;
      if (d >= 2 && d < 10) sp[d - 2] = pw;
    }
    float out[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float sum = t[j] + t[j + 1] + t[j + 2] + t[j + 3] + t[j + 4];
      float g = gv[2 + j] * sp[j] - 2.0f * alpha * beta * xv[4 + j] * sum;
      if (relu_mask && !(xv[4 + j] > 0.f)) g = 0.f;
      out[j] = g;
    }
    st8(dx + row * c + c0, pack8(out));
  }
}

// ------------------------------------------------------------------------------------------------
// max-pool 3x3 stride 2 VALID
// ------------------------------------------------------------------------------------------------
__global__ void maxpool_fwd_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, uint8_t* __restrict__ arg, int n,
                                   int h, int w, int c, int p, int q, long long total_chunks) {
  const int cpr = c >> 3;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total_chunks;
       idx += (long long)gridDim.x * blockDim.x) {
    long long pix = idx / cpr;
    const int c0 = (int)(idx - pix * cpr) * 8;
    const int qq = (int)(pix % q);
    const long long t = pix / q;
    const int pp = (int)(t % p);
    const int nn = (int)(t / p);
    float best[8];
    int bi[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      best[j] = -INFINITY;
      bi[j] = 0;
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const bf16* src = x + (((long long)nn * h + (pp * 2 + r)) * w + (qq * 2 + s)) * c + c0;
        float v[8];
        unpack8(ld8(src), v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (v[j] > best[j]) {  // strict: the first maximum in (h, w) scan order wins, like TF
            best[j] = v[j];
            bi[j] = r * 3 + s;
          }
        }
      }
    }
    st8(y + pix * c + c0, pack8(best));
    if (arg != nullptr) {
      uint32_t lo = bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24);
      uint32_t hi = bi[4] | (bi[5] << 8) | (bi[6] << 16) | (bi[7] << 24);
      *reinterpret_cast<uint2*>(arg + pix * c + c0) = make_uint2(lo, hi);
    }
  }
}

// Second generation of the pooling gradient (pool5, alexnet.py:211): a thread owns 8 channels of a 2 x 2 pixel block,
// so the list of (window, window position) pairs is fixed (see pool_lrn_bwd_kernel4 in encoder_fused.cu) and the gather
// runs on bf16 pairs: argmax byte b -> half-word b << 8, HSET2.EQ against the position code -> 1.0 / 0.0, HFMA2
// accumulates dy; the ReLU mask of the producing layer is one HSET2 + LOP3 per pair.  32-bit index arithmetic.
__device__ __forceinline__ uint32_t mp_heq2(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("set.eq.bf16x2.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ uint32_t mp_hfma2(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ uint32_t mp_gt0_mask(uint32_t a) {
  uint32_t d;
  asm("set.gt.u32.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(0u));
  return d;
}
__device__ __forceinline__ void mp_gather(const bf16* __restrict__ g, const uint8_t* __restrict__ a, uint32_t code2,
                                          uint32_t (&acc)[4]) {
  const uint4 gv = __ldg(reinterpret_cast<const uint4*>(g));
  const uint2 av = __ldg(reinterpret_cast<const uint2*>(a));
  const uint32_t gw[4] = {gv.x, gv.y, gv.z, gv.w};
  const uint32_t h[4] = {__byte_perm(av.x, 0u, 0x1404), __byte_perm(av.x, 0u, 0x3424), __byte_perm(av.y, 0u, 0x1404),
                         __byte_perm(av.y, 0u, 0x3424)};
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i] = mp_hfma2(mp_heq2(h[i], code2), gw[i], acc[i]);
}

__global__ void __launch_bounds__(256)
    maxpool_bwd_kernel2(const bf16* __restrict__ dy, const uint8_t* __restrict__ arg, bf16* __restrict__ dx,
                        const bf16* __restrict__ relu_of, int n, int h, int w, int c, int p, int q, int total) {
  const int cpr = c >> 3;
  const int jp = (h + 1) >> 1, cp = (w + 1) >> 1;
  constexpr uint32_t DEAD = 0xFF00FF00u;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int ch = idx % cpr;
    int t = idx / cpr;
    const int b = t % cp;
    t /= cp;
    const int j = t % jp;
    const int nn = t / jp;
    const int c0 = ch * 8;
    const bool pa_ok = j >= 1 && j - 1 < p, pb_ok = j < p, qa_ok = b >= 1 && b - 1 < q, qb_ok = b < q;
    const int pa = min(max(j - 1, 0), p - 1), pb = min(j, p - 1), qa = min(max(b - 1, 0), q - 1), qb = min(b, q - 1);
    const int pooled = nn * (p * q * c) + c0;  // < 2^31 elements for every tensor this kernel serves (host checks)
    const int o_aa = pooled + (pa * q + qa) * c, o_ab = pooled + (pa * q + qb) * c;
    const int o_ba = pooled + (pb * q + qa) * c, o_bb = pooled + (pb * q + qb) * c;
#define MP_CODE(k) ((uint32_t)(k) * 0x01000100u)
    const uint32_t c_aa = pa_ok && qa_ok ? MP_CODE(8) : DEAD;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
#pragma unroll
      for (int s2 = 0; s2 < 2; ++s2) {
        const int row = 2 * j + r, col = 2 * b + s2;
        if (row >= h || col >= w) continue;
        uint32_t acc[4] = {0u, 0u, 0u, 0u};
        // windows of pixel (r, s2) of the block: rows (j-1: only r = 0, window row 2) and j (window row r);
        // columns (b-1: only s2 = 0, window column 2) and b (window column s2)
        if (r == 0 && s2 == 0) mp_gather(dy + o_aa, arg + o_aa, c_aa, acc);
        if (r == 0) mp_gather(dy + o_ab, arg + o_ab, pa_ok && qb_ok ? MP_CODE(6 + s2) : DEAD, acc);
        if (s2 == 0) mp_gather(dy + o_ba, arg + o_ba, pb_ok && qa_ok ? MP_CODE(3 * r + 2) : DEAD, acc);
        mp_gather(dy + o_bb, arg + o_bb, pb_ok && qb_ok ? MP_CODE(3 * r + s2) : DEAD, acc);
        const int pix = ((nn * h + row) * w + col) * c + c0;
        if (relu_of != nullptr) {
          const uint4 mv = __ldg(reinterpret_cast<const uint4*>(relu_of + pix));
          acc[0] &= mp_gt0_mask(mv.x);
          acc[1] &= mp_gt0_mask(mv.y);
          acc[2] &= mp_gt0_mask(mv.z);
          acc[3] &= mp_gt0_mask(mv.w);
        }
        *reinterpret_cast<uint4*>(dx + pix) = make_uint4(acc[0], acc[1], acc[2], acc[3]);
      }
    }
#undef MP_CODE
  }
}

__global__ void maxpool_bwd_kernel(const bf16* __restrict__ dy, const uint8_t* __restrict__ arg, bf16* __restrict__ dx,
                                   const bf16* __restrict__ relu_of, int n, int h, int w, int c, int p, int q,
                                   long long total_chunks) {
  const int cpr = c >> 3;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total_chunks;
       idx += (long long)gridDim.x * blockDim.x) {
    long long pix = idx / cpr;
    const int c0 = (int)(idx - pix * cpr) * 8;
    const int ww = (int)(pix % w);
    const long long t = pix / w;
    const int hh = (int)(t % h);
    const int nn = (int)(t / h);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const int p_lo = max(0, (hh - 1) >> 1), p_hi = min(p - 1, hh >> 1);
    const int q_lo = max(0, (ww - 1) >> 1), q_hi = min(q - 1, ww >> 1);
    for (int pp = p_lo; pp <= p_hi; ++pp) {
      const int r = hh - 2 * pp;
      if (r < 0 || r > 2) continue;
      for (int qq = q_lo; qq <= q_hi; ++qq) {
        const int s = ww - 2 * qq;
        if (s < 0 || s > 2) continue;
        const long long opix = ((long long)nn * p + pp) * q + qq;
        const uint2 a = *reinterpret_cast<const uint2*>(arg + opix * c + c0);
        float g[8];
        unpack8(ld8(dy + opix * c + c0), g);
        const uint32_t code = r * 3 + s;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t aj = ((j < 4 ? a.x : a.y) >> (8 * (j & 3))) & 0xffu;
          if (aj == code) acc[j] += g[j];
        }
      }
    }
    if (relu_of != nullptr) {
      float m[8];
      unpack8(ld8(relu_of + pix * c + c0), m);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (!(m[j] > 0.f)) acc[j] = 0.f;
    }
    st8(dx + pix * c + c0, pack8(acc));
  }
}

// ------------------------------------------------------------------------------------------------
// Fused LRN + max-pool (forward) and max-pool-grad + LRN-grad + ReLU-grad (+ bias gradient) (backward).
// The normalised tensor n = lrn(a) and its gradient dn are never written to HBM: the forward recomputes the LRN
// inside each 3x3 window (radius 2, beta .75 fast path), the backward rebuilds dn for the 12 channels an 8-channel
// chunk depends on from the pooled gradient and the saved argmax codes.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_window16(const bf16* __restrict__ row, int c0, int c, float (&v)[16]) {
  // channels [c0-4, c0+12) of one pixel, zero outside [0, c)
  float t[8];
  unpack8(ld8(row + c0), t);
#pragma unroll
  for (int j = 0; j < 8; ++j) v[4 + j] = t[j];
  if (c0 >= 8) {
    const uint2 l = *reinterpret_cast<const uint2*>(row + c0 - 4);
    const __nv_bfloat162* lp = reinterpret_cast<const __nv_bfloat162*>(&l);
    float2 a = __bfloat1622float2(lp[0]), b = __bfloat1622float2(lp[1]);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  } else {
    v[0] = v[1] = v[2] = v[3] = 0.f;
  }
  if (c0 + 8 < c) {
    const uint2 r = *reinterpret_cast<const uint2*>(row + c0 + 8);
    const __nv_bfloat162* rp = reinterpret_cast<const __nv_bfloat162*>(&r);
    float2 a = __bfloat1622float2(rp[0]), b = __bfloat1622float2(rp[1]);
    v[12] = a.x; v[13] = a.y; v[14] = b.x; v[15] = b.y;
  } else {
    v[12] = v[13] = v[14] = v[15] = 0.f;
  }
}

__global__ void lrn_pool_fwd_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, uint8_t* __restrict__ arg, int n,
                                    int h, int w, int c, int p, int q, float alpha, float bias,
                                    long long total_chunks) {
  const int cpr = c >> 3;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total_chunks;
       idx += (long long)gridDim.x * blockDim.x) {
    long long pix = idx / cpr;
    const int c0 = (int)(idx - pix * cpr) * 8;
    const int qq = (int)(pix % q);
    const long long t = pix / q;
    const int pp = (int)(t % p);
    const int nn = (int)(t / p);
    float best[8];
    int bi[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      best[j] = -INFINITY;
      bi[j] = 0;
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const bf16* src = x + (((long long)nn * h + (pp * 2 + r)) * w + (qq * 2 + s)) * c;
        float v[16];
        load_window16(src, c0, c, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float acc = 0.f;
#pragma unroll
          for (int o = 2; o <= 6; ++o) acc = fmaf(v[j + o], v[j + o], acc);
          const float sc = bias + alpha * acc;
          const float rs = rsqrtf(sc);
          // the unfused path stores lrn(a) in bf16 before pooling: round here too so both paths agree bit for bit
          const float val = __bfloat162float(__float2bfloat16_rn(v[4 + j] * (rs * sqrtf(rs))));
          if (val > best[j]) {
            best[j] = val;
            bi[j] = r * 3 + s;
          }
        }
      }
    }
    st8(y + pix * c + c0, pack8(best));
    uint32_t lo = bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24);
    uint32_t hi = bi[4] | (bi[5] << 8) | (bi[6] << 16) | (bi[7] << 24);
    *reinterpret_cast<uint2*>(arg + pix * c + c0) = make_uint2(lo, hi);
  }
}

__global__ void pool_lrn_bwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy,
                                    const uint8_t* __restrict__ arg, bf16* __restrict__ dx, float* __restrict__ dbias,
                                    int n, int h, int w, int c, int p, int q, float alpha, float beta, float bias,
                                    long long total_chunks) {
  extern __shared__ float bsum[];  // [c] per-block bias-gradient partials
  if (dbias != nullptr) {
    for (int i = threadIdx.x; i < c; i += blockDim.x) bsum[i] = 0.f;
    __syncthreads();
  }
  const int cpr = c >> 3;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total_chunks;
       idx += (long long)gridDim.x * blockDim.x) {
    long long pix = idx / cpr;
    const int c0 = (int)(idx - pix * cpr) * 8;
    const int ww = (int)(pix % w);
    const long long t = pix / w;
    const int hh = (int)(t % h);
    const int nn = (int)(t / h);
    // dn for channels [c0-2, c0+10): pooled gradient routed through the saved argmax codes (bf16 rounded, as the
    // unfused path stores it)
    float gv[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) gv[j] = 0.f;
    const int p_lo = max(0, (hh - 1) >> 1), p_hi = min(p - 1, hh >> 1);
    const int q_lo = max(0, (ww - 1) >> 1), q_hi = min(q - 1, ww >> 1);
    for (int pp = p_lo; pp <= p_hi; ++pp) {
      const int r = hh - 2 * pp;
      if (r < 0 || r > 2) continue;
      for (int qq = q_lo; qq <= q_hi; ++qq) {
        const int s = ww - 2 * qq;
        if (s < 0 || s > 2) continue;
        const long long opix = ((long long)nn * p + pp) * q + qq;
        const uint32_t code = r * 3 + s;
        const uint8_t* ap = arg + opix * c;
        const bf16* gp = dy + opix * c;
        float g[8];
        unpack8(ld8(gp + c0), g);
        const uint2 a = *reinterpret_cast<const uint2*>(ap + c0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t aj = ((j < 4 ? a.x : a.y) >> (8 * (j & 3))) & 0xffu;
          if (aj == code) gv[2 + j] += g[j];
        }
        if (c0 >= 8) {
#pragma unroll
          for (int j = 0; j < 2; ++j)
            if (ap[c0 - 2 + j] == code) gv[j] += __bfloat162float(gp[c0 - 2 + j]);
        }
        if (c0 + 8 < c) {
#pragma unroll
          for (int j = 0; j < 2; ++j)
            if (ap[c0 + 8 + j] == code) gv[10 + j] += __bfloat162float(gp[c0 + 8 + j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 12; ++j) gv[j] = __bfloat162float(__float2bfloat16_rn(gv[j]));
    float xv[16];
    load_window16(x + pix * c, c0, c, xv);
    float tt[12];
    float sp[8];
#pragma unroll
    for (int d = 0; d < 12; ++d) {
      float acc = 0.f;
#pragma unroll
      for (int o = 0; o < 5; ++o) acc = fmaf(xv[d + o], xv[d + o], acc);
      const float sc = bias + alpha * acc;
      float pw, inv;
      if (beta == 0.75f) {
        const float rs = rsqrtf(sc);
        pw = rs * sqrtf(rs);
        inv = rs * rs;
      } else {
        pw = __powf(sc, -beta);
        inv = 1.0f / sc;
      }
      tt[d] = gv[d] * xv[2 + d] * pw * inv;
      if (d >= 2 && d < 10) sp[d - 2] = pw;
    }
    float out[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float sum = tt[j] + tt[j + 1] + tt[j + 2] + tt[j + 3] + tt[j + 4];
      float g = gv[2 + j] * sp[j] - 2.0f * alpha * beta * xv[4 + j] * sum;
      if (!(xv[4 + j] > 0.f)) g = 0.f;  // ReLU gradient of the producing conv
      out[j] = g;
    }
    const Bf16x8 packed = pack8(out);
    st8(dx + pix * c + c0, packed);
    if (dbias != nullptr) {
      float rb[8];
      unpack8(packed, rb);  // the bias gradient sums the bf16 values that are stored (as vl_colsum would)
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&bsum[c0 + j], rb[j]);
    }
  }
  if (dbias != nullptr) {
    __syncthreads();
    for (int i = threadIdx.x; i < c; i += blockDim.x) atomicAdd(dbias + i, bsum[i]);
  }
}

// ------------------------------------------------------------------------------------------------
// column sums (bias gradients)
// ------------------------------------------------------------------------------------------------
constexpr int COLSUM_THREADS = 256;

// Vector path (c % 8 == 0, ld % 8 == 0, c/8 <= 256): each thread owns one 16-byte column chunk and strides over
// rows, four independent loads in flight; partial sums meet in shared memory, one atomicAdd per column and CTA.
__global__ void __launch_bounds__(COLSUM_THREADS)
    colsum_vec_kernel(const bf16* __restrict__ dy, float* __restrict__ out, long long rows, int c, int ld,
                      long long rows_per_block) {
  __shared__ float red[2048];
  const int cpr = c >> 3;
  const int lanes = COLSUM_THREADS / cpr;  // rows handled in parallel by one CTA
  const int ch = threadIdx.x % cpr;
  const int rl = threadIdx.x / cpr;
  for (int i = threadIdx.x; i < c; i += COLSUM_THREADS) red[i] = 0.f;
  __syncthreads();
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = min(rows, r0 + rows_per_block);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (rl < lanes) {
    const bf16* base = dy + ch * 8;
    long long r = r0 + rl;
    for (; r + 3LL * lanes < r1; r += 4LL * lanes) {
      Bf16x8 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = ld8(base + (r + (long long)u * lanes) * ld);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float f[8];
        unpack8(v[u], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += f[j];
      }
    }
    for (; r < r1; r += lanes) {
      float f[8];
      unpack8(ld8(base + r * ld), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&red[ch * 8 + j], acc[j]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < c; i += COLSUM_THREADS) atomicAdd(out + i, red[i]);
}

__global__ void colsum_kernel(const bf16* __restrict__ dy, float* __restrict__ out, long long rows, int c, int ld,
                              int ct, int ty_count, long long rows_per_block) {
  __shared__ float2 red[COLSUM_THREADS];
  const int tx = threadIdx.x % ct;
  const int ty = threadIdx.x / ct;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = min(rows, r0 + rows_per_block);
  for (int cbase = 0; cbase < c; cbase += 2 * ct) {
    const int col = cbase + 2 * tx;
    float2 acc = make_float2(0.f, 0.f);
    if (ty < ty_count && col < c) {
      if (col + 1 < c && (ld & 1) == 0) {
        for (long long r = r0 + ty; r < r1; r += ty_count) {
          float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(dy + r * ld + col));
          acc.x += v.x;
          acc.y += v.y;
        }
      } else {
        for (long long r = r0 + ty; r < r1; r += ty_count) {
          acc.x += __bfloat162float(dy[r * ld + col]);
          if (col + 1 < c) acc.y += __bfloat162float(dy[r * ld + col + 1]);
        }
      }
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    if (ty == 0 && col < c) {
      float2 s = red[tx];
      for (int j = 1; j < ty_count; ++j) {
        s.x += red[j * ct + tx].x;
        s.y += red[j * ct + tx].y;
      }
      atomicAdd(out + col, s.x);
      if (col + 1 < c) atomicAdd(out + col + 1, s.y);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// operand packing
// ------------------------------------------------------------------------------------------------
__global__ void pack_bf16_kernel(const float* __restrict__ src, int rows, int cols, bf16* __restrict__ dst,
                                 int dst_rows, int dst_ld, int src_grp, int dst_grp, long long total) {
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int R = (int)(idx / dst_ld);
    const int col = (int)(idx - (long long)R * dst_ld);
    const int grp = R / dst_grp;
    const int rr = R - grp * dst_grp;
    float v = 0.f;
    const long long r = (long long)grp * src_grp + rr;
    if (rr < src_grp && col < cols && r < rows) v = src[r * cols + col];
    dst[idx] = __float2bfloat16_rn(v);
  }
}

// Transposed variant: dst[col][(r/src_grp)*dst_grp + r%src_grp] = src[r][col] (K-major operand of the forward
// convolutions: one TMA box per k-block instead of one per 64 output channels).  32x32 tiles through shared memory.
__global__ void pack_bf16_t_kernel(const float* __restrict__ src, int rows, int cols, bf16* __restrict__ dst, int dst_ld,
                                   int src_grp, int dst_grp) {
  __shared__ float tile[32][33];
  const int k0 = blockIdx.x * 32, c0 = blockIdx.y * 32;  // k = packed row index, c = column of src
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int kk = k0 + i, col = c0 + threadIdx.x;
    const int grp = kk / dst_grp, rr = kk - grp * dst_grp;
    const long long r = (long long)grp * src_grp + rr;
    float v = 0.f;
    if (kk < dst_ld && rr < src_grp && r < rows && col < cols) v = src[r * cols + col];
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int col = c0 + i, kk = k0 + threadIdx.x;
    if (col < cols && kk < dst_ld) dst[(long long)col * dst_ld + kk] = __float2bfloat16_rn(tile[threadIdx.x][i]);
  }
}

__global__ void cast_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long long n) {
  const long long n8 = n >> 3;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n8;
       idx += (long long)gridDim.x * blockDim.x) {
    const float4 a = reinterpret_cast<const float4*>(src)[2 * idx];
    const float4 b = reinterpret_cast<const float4*>(src)[2 * idx + 1];
    float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    st8(reinterpret_cast<Bf16x8*>(dst) + idx, pack8(f));
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long long i = n8 << 3; i < n; ++i) dst[i] = __float2bfloat16_rn(src[i]);
}

// dst[i] = bf16(src[table[i]]) (0 where table[i] < 0): every operand copy ("shadow") of the convolution filters is a
// fixed permutation / zero padding of the fp32 master, so ONE launch over a precomputed index table refreshes all of
// them after the optimiser step (Engine.refresh_shadows) instead of one pack kernel per layer and layout.
// A thread produces 8 consecutive outputs: two 16-byte table loads, 8 gathers (the masters are L2 resident), one
// 16-byte store.
__global__ void __launch_bounds__(256)
    gather_bf16_kernel(const float* __restrict__ src, const int32_t* __restrict__ table, bf16* __restrict__ dst,
                       long long n8) {
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n8;
       idx += (long long)gridDim.x * blockDim.x) {
    const int4 t0 = __ldg(reinterpret_cast<const int4*>(table) + 2 * idx);
    const int4 t1 = __ldg(reinterpret_cast<const int4*>(table) + 2 * idx + 1);
    const int t[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = t[j] >= 0 ? __ldg(src + t[j]) : 0.f;
    st8(reinterpret_cast<Bf16x8*>(dst) + idx, pack8(f));
  }
}

// dst[c][r] = src[r][c]  (fp32, 32x32 smem tiles; used for the recurrent weights w_h -> w_h^T of the BPTT kernel)
__global__ void transpose_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[i][threadIdx.x] = src[(long long)r * cols + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) dst[(long long)c * rows + r] = tile[threadIdx.x][i];
  }
}

int sweep_grid(long long work_items, int block) {
  long long g = (work_items + block - 1) / block;
  long long cap = (long long)vl::num_sms() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace

#define VL_LAUNCHED()            \
  do {                           \
    vl::g_launches.fetch_add(1); \
    VL_CHECK_CUDA(cudaGetLastError()); \
  } while (0)

extern "C" int vl_lrn_fwd(const void* x, void* y, int64_t rows, int32_t c, int32_t radius, float alpha, float beta,
                          float bias, vl_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VL_REQUIRE(x && y && c % 8 == 0 && radius >= 0 && radius <= 4, "vl_lrn_fwd: c must be a multiple of 8, radius <= 4");
  const long long total = rows * (c / 8);
  if (radius == 2)
    lrn_fwd_kernel<2><<<sweep_grid(total, 256), 256, 0, stream>>>(
        reinterpret_cast<const bf16*>(x), reinterpret_cast<bf16*>(y), rows, c, radius, alpha, beta, bias, total);
  else
    lrn_fwd_kernel<0><<<sweep_grid(total, 256), 256, 0, stream>>>(
        reinterpret_cast<const bf16*>(x), reinterpret_cast<bf16*>(y), rows, c, radius, alpha, beta, bias, total);
  VL_LAUNCHED();
  return 0;
}

extern "C" int vl_lrn_bwd(const void* x, const void* dy, void* dx, int64_t rows, int32_t c, int32_t radius,
                          float alpha, float beta, float bias, int32_t relu_mask, vl_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VL_REQUIRE(x && dy && dx && c % 8 == 0, "vl_lrn_bwd: c must be a multiple of 8");
  VL_REQUIRE(radius == 2, "vl_lrn_bwd: only depth_radius 2 (alexnet.py:80,121) is implemented");
  const long long total = rows * (c / 8);
  lrn_bwd_kernel<<<sweep_grid(total, 256), 256, 0, stream>>>(
      reinterpret_cast<const bf16*>(x), reinterpret_cast<const bf16*>(dy), reinterpret_cast<bf16*>(dx), rows, c, alpha,
      beta, bias, relu_mask, total);
  VL_LAUNCHED();
  return 0;
}

extern "C" int vl_maxpool_fwd(const void* x, void* y, void* argmax, int32_t n, int32_t h, int32_t w, int32_t c,
                              vl_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VL_REQUIRE(x && y && c % 8 == 0 && h >= 3 && w >= 3, "vl_maxpool_fwd: bad arguments");
  const int p = (h - 3) / 2 + 1, q = (w - 3) / 2 + 1;
  const long long total = (long long)n * p * q * (c / 8);
  maxpool_fwd_kernel<<<sweep_grid(total, 256), 256, 0, stream>>>(reinterpret_cast<const bf16*>(x),
                                                                 reinterpret_cast<bf16*>(y),
                                                                 reinterpret_cast<uint8_t*>(argmax), n, h, w, c, p, q,
                                                                 total);
  VL_LAUNCHED();
  return 0;
}

extern "C" int vl_maxpool_bwd(const void* dy, const void* argmax, void* dx, const void* relu_of, int32_t n, int32_t h,
                              int32_t w, int32_t c, vl_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VL_REQUIRE(dy && argmax && dx && c % 8 == 0, "vl_maxpool_bwd: bad arguments");
  const int p = (h - 3) / 2 + 1, q = (w - 3) / 2 + 1;
  const long long total = (long long)n * h * w * (c / 8);
  if ((long long)n * h * w * c < (1LL << 31) && !getenv("VL_MAXPOOL_BWD_V1")) {
    // 2 x 2 pixel blocks, packed gather (the accumulation in bf16 is exact for one contribution and rounds once per
    // further one, like the fused LRN / pool backward)
    const long long blocks2 = (long long)n * ((h + 1) / 2) * ((w + 1) / 2) * (c / 8);
    maxpool_bwd_kernel2<<<sweep_grid(blocks2, 256), 256, 0, stream>>>(
        reinterpret_cast<const bf16*>(dy), reinterpret_cast<const uint8_t*>(argmax), reinterpret_cast<bf16*>(dx),
        reinterpret_cast<const bf16*>(relu_of), n, h, w, c, p, q, (int)blocks2);
    VL_LAUNCHED();
    return 0;
  }
  maxpool_bwd_kernel<<<sweep_grid(total, 256), 256, 0, stream>>>(
      reinterpret_cast<const bf16*>(dy), reinterpret_cast<const uint8_t*>(argmax), reinterpret_cast<bf16*>(dx),
      reinterpret_cast<const bf16*>(relu_of), n, h, w, c, p, q, total);
  VL_LAUNCHED();
  return 0;
}

extern "C" int vl_colsum(const void* dy, float* out, int64_t rows, int32_t c, int32_t ld, vl_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VL_REQUIRE(dy && out && rows > 0 && c > 0 && ld >= c, "vl_colsum: bad arguments");
  if (c % 8 == 0 && ld % 8 == 0 && (reinterpret_cast<uintptr_t>(dy) & 15) == 0) {
    // wide matrices (fc layers: 4096 columns) go through in column blocks of up to 2048
    for (int c0 = 0; c0 < c; c0 += 2048) {
      const int cb = c - c0 < 2048 ? c - c0 : 2048;
      const int lanes = COLSUM_THREADS / (cb / 8);  // cb / 8 <= 256
      long long blocks = (long long)vl::num_sms() * 8;
      long long rpb = (rows + blocks - 1) / blocks;
      if (rpb < 4LL * lanes) rpb = 4LL * lanes;
      blocks = (rows + rpb - 1) / rpb;
      colsum_vec_kernel<<<(int)blocks, COLSUM_THREADS, 0, stream>>>(reinterpret_cast<const bf16*>(dy) + c0, out + c0,
                                                                    rows, cb, ld, rpb);
      VL_LAUNCHED();
    }
    return 0;
  }
  int ct = (c + 1) / 2;
  if (ct > 128) ct = 128;
  int ty = COLSUM_THREADS / ct;
  long long blocks = (long long)vl::num_sms() * 4;
  long long min_rows = (long long)ty * 8;
  long long rpb = (rows + blocks - 1) / blocks;
  if (rpb < min_rows) rpb = min_rows;
  blocks = (rows + rpb - 1) / rpb;
  colsum_kernel<<<(int)blocks, COLSUM_THREADS, 0, stream>>>(reinterpret_cast<const bf16*>(dy), out, rows, c, ld, ct, ty,
                                                            rpb);
  VL_LAUNCHED();
  return 0;
}

extern "C" int vl_pack_bf16(const float* src, int32_t rows, int32_t cols, void* dst, int32_t dst_rows, int32_t dst_ld,
                            int32_t src_grp, int32_t dst_grp, vl_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VL_REQUIRE(src && dst && src_grp > 0 && dst_grp >= src_grp && dst_ld >= cols, "vl_pack_bf16: bad arguments");
  const long long total = (long long)dst_rows * dst_ld;
  pack_bf16_kernel<<<sweep_grid(total, 256), 256, 0, stream>>>(src, rows, cols, reinterpret_cast<bf16*>(dst), dst_rows,
                                                               dst_ld, src_grp, dst_grp, total);
  VL_LAUNCHED();
  return 0;
}

extern "C" int vl_pack_bf16_t(const float* src, int32_t rows, int32_t cols, void* dst, int32_t dst_ld, int32_t src_grp,
                              int32_t dst_grp, vl_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VL_REQUIRE(src && dst && src_grp > 0 && dst_grp >= src_grp && dst_ld > 0, "vl_pack_bf16_t: bad arguments");
  dim3 grid((dst_ld + 31) / 32, (cols + 31) / 32), block(32, 8);
  pack_bf16_t_kernel<<<grid, block, 0, stream>>>(src, rows, cols, reinterpret_cast<bf16*>(dst), dst_ld, src_grp, dst_grp);
  VL_LAUNCHED();
  return 0;
}

extern "C" int vl_cast_f32_to_bf16(const float* src, void* dst, int64_t n, vl_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VL_REQUIRE(src && dst && n >= 0, "vl_cast_f32_to_bf16: bad arguments");
  VL_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0,
             "vl_cast_f32_to_bf16: pointers must be 16B aligned");
  cast_bf16_kernel<<<sweep_grid((n + 7) / 8, 256), 256, 0, stream>>>(src, reinterpret_cast<bf16*>(dst), n);
  VL_LAUNCHED();
  return 0;
}

extern "C" int vl_gather_bf16(const float* src, const int32_t* table, void* dst, int64_t n, vl_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VL_REQUIRE(src && table && dst && n > 0 && n % 8 == 0, "vl_gather_bf16: bad arguments (n must be a multiple of 8)");
  VL_REQUIRE((reinterpret_cast<uintptr_t>(table) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0,
             "vl_gather_bf16: table and dst must be 16-byte aligned");
  gather_bf16_kernel<<<sweep_grid(n / 8, 256), 256, 0, stream>>>(src, table, reinterpret_cast<bf16*>(dst), n / 8);
  VL_LAUNCHED();
  return 0;
}

extern "C" int vl_transpose_f32(const float* src, float* dst, int32_t rows, int32_t cols, vl_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VL_REQUIRE(src && dst && rows > 0 && cols > 0, "vl_transpose_f32: bad arguments");
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  transpose_f32_kernel<<<grid, block, 0, stream>>>(src, dst, rows, cols);
  VL_LAUNCHED();
  return 0;
}

extern "C" int vl_lrn_pool_fwd_generic(const void* x, void* y, void* argmax, int32_t n, int32_t h, int32_t w, int32_t c,
                               int32_t radius, float alpha, float beta, float bias, vl_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VL_REQUIRE(x && y && argmax && c % 8 == 0 && h >= 3 && w >= 3, "vl_lrn_pool_fwd: bad arguments");
  VL_REQUIRE(radius == 2 && beta == 0.75f, "vl_lrn_pool_fwd_generic: fused path serves depth_radius 2, beta 0.75 (alexnet.py:80-89)");
  const int p = (h - 3) / 2 + 1, q = (w - 3) / 2 + 1;
  const long long total = (long long)n * p * q * (c / 8);
  lrn_pool_fwd_kernel<<<sweep_grid(total, 256), 256, 0, stream>>>(reinterpret_cast<const bf16*>(x),
                                                                  reinterpret_cast<bf16*>(y),
                                                                  reinterpret_cast<uint8_t*>(argmax), n, h, w, c, p, q,
                                                                  alpha, bias, total);
  VL_LAUNCHED();
  return 0;
}

extern "C" int vl_pool_lrn_bwd_generic(const void* x, const void* dy, const void* argmax, void* dx, float* dbias, int32_t n,
                               int32_t h, int32_t w, int32_t c, int32_t radius, float alpha, float beta, float bias,
                               vl_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VL_REQUIRE(x && dy && argmax && dx && c % 8 == 0, "vl_pool_lrn_bwd: bad arguments");
  VL_REQUIRE(radius == 2, "vl_pool_lrn_bwd: only depth_radius 2 (alexnet.py:80,121) is implemented");
  const int p = (h - 3) / 2 + 1, q = (w - 3) / 2 + 1;
  const long long total = (long long)n * h * w * (c / 8);
  pool_lrn_bwd_kernel<<<sweep_grid(total, 256), 256, dbias ? c * sizeof(float) : 0, stream>>>(
      reinterpret_cast<const bf16*>(x), reinterpret_cast<const bf16*>(dy), reinterpret_cast<const uint8_t*>(argmax),
      reinterpret_cast<bf16*>(dx), dbias, n, h, w, c, p, q, alpha, beta, bias, total);
  VL_LAUNCHED();
  return 0;
}
