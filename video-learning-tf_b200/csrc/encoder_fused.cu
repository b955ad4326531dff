// HBM-bound kernels of the AlexNet encoder (sm_100a).
//
//   vl_frames_s2d(_crop)   frames (uint8 / fp32) -> mean-subtracted bf16 space-to-depth tensor: the 11x11 stride-4
//                          SAME conv1 (alexnet.py:60-77) becomes a 3x3 stride-1 VALID convolution over 48 channels
//                          (frames_s2d_direct_kernel: register-only fast path for plain uint8 frames;
//                          frames_s2d_kernel: crop / mirror / fp32 feed)
//   vl_s2d_pack_filter     conv1 HWIO fp32 filter -> bf16 [9 taps x 64, cout] operand of that convolution
//   vl_s2d_unpack_grad     filter gradient of the 3x3x48 convolution -> HWIO gradient of the 11x11x3 filter
//   vl_lrn_pool_fwd        LRN + 3x3/2 max-pool (alexnet.py:80-98,121-139): lrn_pool_fwd_kernel4 (registers and warp
//                          shuffles only, separable pooling) for the two AlexNet geometries; lrn_pool_fwd_kernel2
//                          (strip per CTA in shared memory, run-time extents) serves every other shape
//   vl_pool_lrn_bwd        MaxPoolGrad -> LRNGrad -> ReluGrad (+ bias gradient): pool_lrn_bwd_kernel4 (2 x 2 pixel
//                          blocks, packed gather) for the two AlexNet geometries, pool_lrn_bwd_kernel2 (per pixel,
//                          run-time extents) otherwise
//
// All global accesses are 16-byte vectors over the channel axis of NHWC tensors (spelled through uint4: a copy of a
// struct of four bf16x2 compiles to four 32-bit accesses); these kernels are judged on achieved HBM GB/s (DESIGN.md
// lists the algorithmic bytes and the instruction budget of each).
#include "common.cuh"
#include "../../include/vlb200.h"

#include <atomic>
#include <cstdlib>

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
__device__ __forceinline__ float sqrt_approx(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rsqrt_approx(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float bf16_round_f(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// Channel halo of an 8-channel chunk inside a group of LPP lanes that together hold one pixel: lane l owns channels
// [8l, 8l+8); returns the two values left of the chunk (from lane l-1) and right of it (from lane l+1), zero outside.
template <int LPP>
__device__ __forceinline__ void halo2(const float (&v)[8], int l, int cpr, float (&left)[2], float (&right)[2]) {
  left[0] = __shfl_up_sync(0xffffffffu, v[6], 1, LPP);
  left[1] = __shfl_up_sync(0xffffffffu, v[7], 1, LPP);
  right[0] = __shfl_down_sync(0xffffffffu, v[0], 1, LPP);
  right[1] = __shfl_down_sync(0xffffffffu, v[1], 1, LPP);
  if (l == 0) left[0] = left[1] = 0.f;
  if (l >= cpr - 1) right[0] = right[1] = 0.f;
}

// s_j = bias + alpha * sum_{|o|<=2} x_{j+o}^2 for the 8 own channels (tf.nn.lrn, depth_radius 2)
template <int LPP>
__device__ __forceinline__ void lrn_scale8(const float (&x)[8], int l, int cpr, float alpha, float bias, float (&s)[8]) {
  float sq[12];
#pragma unroll
  for (int j = 0; j < 8; ++j) sq[2 + j] = x[j] * x[j];
  float own[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) own[j] = sq[2 + j];
  float lf[2], rt[2];
  halo2<LPP>(own, l, cpr, lf, rt);
  sq[0] = lf[0];
  sq[1] = lf[1];
  sq[10] = rt[0];
  sq[11] = rt[1];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = fmaf(alpha, ((sq[j] + sq[j + 1]) + sq[j + 2]) + (sq[j + 3] + sq[j + 4]), bias);
}

// Multi-chunk form: a lane owns NE = 8*CPL consecutive channels, LPP lanes hold one pixel (LPP*NE == c for the
// instantiated geometries, so no lane idles).  Halo = the two neighbours on each side, exchanged by shuffle.
template <int LPP, int NE>
__device__ __forceinline__ void halo2n(const float (&v)[NE], int l, float (&left)[2], float (&right)[2]) {
  left[0] = __shfl_up_sync(0xffffffffu, v[NE - 2], 1, LPP);
  left[1] = __shfl_up_sync(0xffffffffu, v[NE - 1], 1, LPP);
  right[0] = __shfl_down_sync(0xffffffffu, v[0], 1, LPP);
  right[1] = __shfl_down_sync(0xffffffffu, v[1], 1, LPP);
  if (l == 0) left[0] = left[1] = 0.f;
  if (l == LPP - 1) right[0] = right[1] = 0.f;
}
// w[j] = v[j-2] + ... + v[j+2] over the pixel's channel axis (zero outside), for the lane's NE channels
template <int LPP, int NE>
__device__ __forceinline__ void window5n(const float (&v)[NE], int l, float (&w)[NE]) {
  float lf[2], rt[2];
  halo2n<LPP, NE>(v, l, lf, rt);
  float e[NE + 4];
  e[0] = lf[0];
  e[1] = lf[1];
#pragma unroll
  for (int j = 0; j < NE; ++j) e[2 + j] = v[j];
  e[NE + 2] = rt[0];
  e[NE + 3] = rt[1];
  // 2 adds per channel (was 3 after common-subexpression elimination): for even j, q4 = e[j+1..j+4] from two pair sums at
  // odd offsets (each shared by two q4), w[j] = e[j] + q4, w[j+1] = q4 + e[j+5]; NE is even
  float pr[NE / 2 + 1];
#pragma unroll
  for (int k = 0; k <= NE / 2; ++k) pr[k] = e[2 * k + 1] + e[2 * k + 2];
#pragma unroll
  for (int i = 0; i < NE / 2; ++i) {
    const float q4 = pr[i] + pr[i + 1];
    w[2 * i] = e[2 * i] + q4;
    w[2 * i + 1] = q4 + e[2 * i + 5];
  }
}

// ------------------------------------------------------------------------------------------------
// frames -> space-to-depth bf16.  One CTA per (frame, block row), one thread per image pixel: its 3 channels of the
// S image rows are mean-subtracted, rounded to bf16 and written straight to their permuted position of the output
// row in shared memory: out[bx][(dy*S+dx)*3+c] = frame[S*by-pt+dy][S*bx-pl+dx][c]; the finished row
// (wb * S*S*3 bf16, contiguous in HBM) is then copied out with 16-byte stores.
// ------------------------------------------------------------------------------------------------
template <bool U8, int S>
__global__ void __launch_bounds__(256)
    frames_s2d_kernel(const void* __restrict__ frames_, const float* __restrict__ mean3, bf16* __restrict__ out, int h,
                      int w, int pad_top, int pad_left, int hb, int wb, int hr, int wr,
                      const int32_t* __restrict__ crops) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* orow_s = reinterpret_cast<bf16*>(smem_raw);  // [wb][S*S*3]
  constexpr int SEG = S * 3;                          // contiguous source elements per (bx, dy)
  constexpr int CBLK = S * S * 3;
  const int by = blockIdx.x % hb;
  const int nn = blockIdx.x / hb;
  float m[3] = {0.f, 0.f, 0.f};
  if (U8 && mean3 != nullptr) {
    m[0] = mean3[0];
    m[1] = mean3[1];
    m[2] = mean3[2];
  }
  // zero fill (SAME padding and rows outside the image)
  for (int i = threadIdx.x; i < wb * CBLK / 8; i += blockDim.x) reinterpret_cast<uint4*>(orow_s)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  // crop window of this frame inside the stored (hr x wr) frame and horizontal mirror (dataset_.py:444-461,498-500)
  int y0 = 0, x0 = 0, mirror = 0;
  if (crops != nullptr) {
    // clamped into the stored frame: offsets that arrive as device tensors cannot be validated on the host without a
    // synchronisation, and an out-of-range window must never read outside the frame buffer
    y0 = max(0, min(crops[nn * 3], hr - h));
    x0 = max(0, min(crops[nn * 3 + 1], wr - w));
    mirror = crops[nn * 3 + 2] != 0;
  }
  // one thread per image pixel x: its 3 channels of the S image rows go to block bx = (x+pad_left)/S, slot dx
  for (int x = threadIdx.x; x < w; x += blockDim.x) {
    const int bx = (x + pad_left) / S, dx = (x + pad_left) - bx * S;
    bf16* dstp = orow_s + bx * CBLK + dx * 3;
    const int xs = x0 + (mirror ? (w - 1 - x) : x);
#pragma unroll
    for (int dy = 0; dy < S; ++dy) {
      const int y = by * S - pad_top + dy;
      if (y < 0 || y >= h) continue;
      const long long off = (((long long)nn * hr + (y0 + y)) * wr + xs) * 3;
      float v0, v1, v2;
      if (U8) {
        const uint8_t* src = reinterpret_cast<const uint8_t*>(frames_) + off;
        v0 = (float)__ldg(src) - m[0];
        v1 = (float)__ldg(src + 1) - m[1];
        v2 = (float)__ldg(src + 2) - m[2];
      } else {
        const float* src = reinterpret_cast<const float*>(frames_) + off;
        v0 = __ldg(src);
        v1 = __ldg(src + 1);
        v2 = __ldg(src + 2);
      }
      dstp[dy * SEG] = __float2bfloat16_rn(v0);
      dstp[dy * SEG + 1] = __float2bfloat16_rn(v1);
      dstp[dy * SEG + 2] = __float2bfloat16_rn(v2);
    }
  }
  __syncthreads();
  uint4* dst = reinterpret_cast<uint4*>(out + ((long long)nn * hb + by) * wb * CBLK);
  for (int i = threadIdx.x; i < wb * CBLK / 8; i += blockDim.x) dst[i] = reinterpret_cast<const uint4*>(orow_s)[i];
}

// ------------------------------------------------------------------------------------------------
// Fast path of the staging for uint8 frames without crop / mirror and a block-aligned left padding (the bench and
// config-2 feed): with pad_left % 4 == 0 the 12 source bytes of the 4 pixels of a (block column, dy) pair are CONTIGUOUS
// in the frame row and land on 12 CONTIGUOUS bf16 of the output row.  A thread owns one output block (frame, by, bx) =
// 4 image rows x 12 source bytes -> 48 contiguous bf16 (96 bytes): four aligned 32-bit loads + funnel shifts per row
// (the frame rows are 681 bytes, i.e. unaligned), exact uint8 -> fp32 through the 2^23 mantissa trick, no shared
// memory and no barriers.  Loads of a warp walk 32 x 12 contiguous bytes of each of the four rows,
// stores cover 32 x 96 contiguous bytes.
__global__ void __launch_bounds__(128)
    frames_s2d_direct_kernel(const uint8_t* __restrict__ frames, const float* __restrict__ mean3, bf16* __restrict__ out,
                             int h, int w, int pad_top, int pad_left, int hb, int wb, long long total_bytes,
                             int n_rows /* n * hb */, int wide) {
  constexpr int S = 4, SEG = 12, CBLK = 48;
  float m[3] = {0.f, 0.f, 0.f};
  if (mean3 != nullptr) {
    m[0] = mean3[0];
    m[1] = mean3[1];
    m[2] = mean3[2];
  }
  const int row_bytes = w * 3;
  const int bx = threadIdx.x & 63;           // 64 block columns per half CTA (wb <= 64), two block rows per CTA
  const int half = threadIdx.x >> 6;
  const int xb = bx * SEG - pad_left * 3;
  const bool seg_any = bx < wb && xb + SEG > 0 && xb < row_bytes;
  const bool seg_full = xb >= 0 && xb + SEG <= row_bytes;
  const uintptr_t lo = reinterpret_cast<uintptr_t>(frames), hi = lo + (uintptr_t)total_bytes;
  for (int rowi = blockIdx.x * 2 + half; rowi < n_rows; rowi += gridDim.x * 2) {
    if (bx >= wb) continue;
    const int nn = rowi / hb;
    const int by = rowi - nn * hb;
    const long long frame_off = (long long)nn * h * w * 3;
    uint32_t wd[S][4];
    bool ok[S];
#pragma unroll
    for (int dy = 0; dy < S; ++dy) {
      const int y = by * S - pad_top + dy;
      ok[dy] = seg_any && y >= 0 && y < h;
      wd[dy][0] = wd[dy][1] = wd[dy][2] = wd[dy][3] = 0u;
      if (ok[dy]) {
        const uintptr_t a = lo + (uintptr_t)(frame_off + (long long)y * row_bytes + xb);
        const uintptr_t a4 = a & ~(uintptr_t)3;
        if (a4 >= lo && a4 + 16 <= hi) {
#pragma unroll
          for (int k = 0; k < 4; ++k) wd[dy][k] = __ldg(reinterpret_cast<const uint32_t*>(a4) + k);
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            for (int j = 0; j < 4; ++j) {
              const uintptr_t pb = a4 + 4 * k + j;
              if (pb >= lo && pb < hi) wd[dy][k] |= (uint32_t)__ldg(reinterpret_cast<const uint8_t*>(pb)) << (8 * j);
            }
        }
      }
    }
    uint32_t o[24];
#pragma unroll
    for (int dy = 0; dy < S; ++dy) {
      const int y = by * S - pad_top + dy;
      const uintptr_t a = lo + (uintptr_t)(frame_off + (long long)y * row_bytes + xb);
      const uint32_t sh = (uint32_t)(a & 3) * 8u;
      const uint32_t b3[3] = {__funnelshift_r(wd[dy][0], wd[dy][1], sh), __funnelshift_r(wd[dy][1], wd[dy][2], sh),
                              __funnelshift_r(wd[dy][2], wd[dy][3], sh)};
      float v[12];
#pragma unroll
      for (int e = 0; e < 12; ++e) {
        const uint32_t bits = __byte_perm(b3[e >> 2], 0x4B000000u, 0x7440u | (uint32_t)(e & 3));
        v[e] = (__uint_as_float(bits) - 8388608.0f) - m[e % 3];
      }
      if (!seg_full) {
#pragma unroll
        for (int e = 0; e < 12; ++e)
          if (xb + e < 0 || xb + e >= row_bytes) v[e] = 0.f;
      }
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const __nv_bfloat162 pk = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        o[dy * 6 + i] = ok[dy] ? *reinterpret_cast<const uint32_t*>(&pk) : 0u;
      }
    }
    bf16* dstp = out + ((long long)rowi * wb + bx) * CBLK;
    if (wide && (reinterpret_cast<uintptr_t>(dstp) & 31) == 0) {
      // three 32-byte stores: the lanes of a warp are 96 bytes apart, so a 16-byte store fills only half a sector
#pragma unroll
      for (int k = 0; k < 3; ++k)
        asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dstp + 16 * k), "r"(o[8 * k]),
                     "r"(o[8 * k + 1]), "r"(o[8 * k + 2]), "r"(o[8 * k + 3]), "r"(o[8 * k + 4]), "r"(o[8 * k + 5]),
                     "r"(o[8 * k + 6]), "r"(o[8 * k + 7])
                     : "memory");
    } else {
      uint4* dst = reinterpret_cast<uint4*>(dstp);
#pragma unroll
      for (int k = 0; k < 6; ++k) dst[k] = make_uint4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
    }
  }
}

// dst[(tr*kb+ts)*chunk + (dy*s+dx)*cin + c][o] = src[s*tr+dy][s*ts+dx][c][o] (0 when outside the kh x kw filter or
// in the chunk padding).  kb = ceil(kh/s) taps per axis of the space-to-depth convolution.
__global__ void s2d_pack_filter_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int kh, int kw, int cin,
                                       int cout, int s, int kb_h, int kb_w, int chunk, int transpose, long long total) {
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int o = (int)(idx % cout);
    const int row = (int)(idx / cout);
    const int tap = row / chunk;
    const int j = row - tap * chunk;
    const int tr = tap / kb_w, ts = tap - tr * kb_w;
    float v = 0.f;
    if (j < s * s * cin) {
      const int c = j % cin;
      const int dd = j / cin;
      const int dy = dd / s, dx = dd - dy * s;
      const int r = s * tr + dy, q = s * ts + dx;
      if (r < kh && q < kw) v = src[(((long long)r * kw + q) * cin + c) * cout + o];
    }
    if (transpose)
      dst[(long long)o * (kb_h * kb_w * chunk) + row] = __float2bfloat16_rn(v);
    else
      dst[idx] = __float2bfloat16_rn(v);
  }
}

// dw[r][q][c][o] = dws[(tr*kb_w+ts)*(s*s*cin) + (dy*s+dx)*cin + c][o]  with r = s*tr+dy, q = s*ts+dx
__global__ void s2d_unpack_grad_kernel(const float* __restrict__ dws, float* __restrict__ dw, int kh, int kw, int cin,
                                       int cout, int s, int kb_w, long long total) {
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int o = (int)(idx % cout);
    long long t = idx / cout;
    const int c = (int)(t % cin);
    t /= cin;
    const int q = (int)(t % kw);
    const int r = (int)(t / kw);
    const int tr = r / s, dy = r - tr * s, ts = q / s, dx = q - ts * s;
    const long long row = (long long)(tr * kb_w + ts) * (s * s * cin) + (dy * s + dx) * cin + c;
    dw[idx] = dws[row * cout + o];
  }
}

// ------------------------------------------------------------------------------------------------
// LRN + max-pool forward.  CTA = (frame, strip of `rows_out` pooled rows).  Phase 1 evaluates lrn(x) once for every
// input pixel of the strip (LPP lanes per pixel, 16 B per lane) into shared memory as bf16; phase 2 pools 3x3/2
// windows out of shared memory.  beta = 0.75 fast path: s^-0.75 = rsqrt(s) * sqrt(rsqrt(s)).
// ------------------------------------------------------------------------------------------------
template <int LPP, int CPL>
__global__ void __launch_bounds__(512)
    lrn_pool_fwd_kernel2(const bf16* __restrict__ x, bf16* __restrict__ y, uint8_t* __restrict__ arg, int h, int w, int c,
                         int p, int q, int rows_out, int strips, float alpha, float bias) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* tile = reinterpret_cast<bf16*>(smem_raw);  // [(2*rows_out+1)][w][c]
  constexpr int NE = 8 * CPL;                      // channels per lane; LPP * NE == c (host guarantees it)
  const int strip = blockIdx.x % strips;
  const int nn = blockIdx.x / strips;
  const int p0 = strip * rows_out;
  const int np = min(rows_out, p - p0);  // pooled rows of this strip
  const int in_rows = 2 * np + 1;
  const int cpr = c >> 3;
  const int l = threadIdx.x % LPP;
  const int grp = threadIdx.x / LPP;
  const int ngrp = blockDim.x / LPP;
  const bf16* xin = x + ((long long)nn * h + 2 * p0) * w * c;
  const int npix = in_rows * w;
  // phase 1: LPP lanes hold one pixel (NE channels each, CPL independent 16-byte loads in flight per lane); every
  // lane of a group takes part in the halo shuffles
  for (int pix0 = 0; pix0 < npix; pix0 += ngrp) {
    const int pix = pix0 + grp;
    const bool live = pix < npix;
    float v[NE];
    if (live) {
      Bf16x8 in[CPL];
#pragma unroll
      for (int k = 0; k < CPL; ++k) in[k] = ld8(xin + (long long)pix * c + l * NE + 8 * k);
#pragma unroll
      for (int k = 0; k < CPL; ++k) {
        float t8[8];
        unpack8(in[k], t8);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[8 * k + j] = t8[j];
      }
    } else {
#pragma unroll
      for (int j = 0; j < NE; ++j) v[j] = 0.f;
    }
    float sq[NE], ssum[NE];
#pragma unroll
    for (int j = 0; j < NE; ++j) sq[j] = v[j] * v[j];
    window5n<LPP, NE>(sq, l, ssum);
    if (live) {
#pragma unroll
      for (int k = 0; k < CPL; ++k) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float rs = rsqrt_approx(fmaf(alpha, ssum[8 * k + j], bias));
          o[j] = v[8 * k + j] * (rs * sqrt_approx(rs));
        }
        st8(tile + pix * c + l * NE + 8 * k, pack8(o));
      }
    }
  }
  __syncthreads();
  // phase 2: 3x3/2 windows out of shared memory, two channels per instruction (bf16x2 compare / max / select)
  const int nout = np * q * cpr;
  for (int idx = threadIdx.x; idx < nout; idx += blockDim.x) {
    const int ch = idx % cpr;
    const int t = idx / cpr;
    const int qq = t % q;
    const int pl = t / q;
    __nv_bfloat162 best[4];
    uint32_t bidx[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t ninf = 0xFF80FF80u;  // (-inf, -inf)
      best[i] = *reinterpret_cast<const __nv_bfloat162*>(&ninf);
      bidx[i] = 0;
    }
    const bf16* wbase = tile + ((2 * pl) * w + 2 * qq) * c + ch * 8;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
      for (int s2 = 0; s2 < 3; ++s2) {
        const Bf16x8 v = ld8(wbase + (r * w + s2) * c);
        const uint32_t code2 = (uint32_t)(r * 3 + s2) * 0x00010001u;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          // strict >: the first maximum in (h, w) scan order wins, like TF
          const uint32_t m = __hgt2_mask(v.v[i], best[i]);
          best[i] = __hmax2(best[i], v.v[i]);
          bidx[i] = (bidx[i] & ~m) | (code2 & m);
        }
      }
    }
    const long long opix = ((long long)nn * p + (p0 + pl)) * q + qq;
    Bf16x8 o;
#pragma unroll
    for (int i = 0; i < 4; ++i) o.v[i] = best[i];
    st8(y + opix * c + ch * 8, o);
    const uint32_t lo = __byte_perm(bidx[0], bidx[1], 0x6420);
    const uint32_t hi = __byte_perm(bidx[2], bidx[3], 0x6420);
    *reinterpret_cast<uint2*>(arg + opix * c + ch * 8) = make_uint2(lo, hi);
  }
}

// ------------------------------------------------------------------------------------------------
// MaxPoolGrad -> LRNGrad -> ReluGrad (+ bias gradient).
//   dn_i = sum over pooling windows whose argmax is pixel i of dy          (bf16 rounded, as the unfused path stores it)
//   dx_i = relu'(x_i) * ( dn_i * s_i^-b  -  2ab x_i * sum_{|d-i|<=2} dn_d x_d s_d^(-b-1) )
// CTAs walk image rows (no per-thread 64-bit index arithmetic); LPP lanes hold one pixel, 16 bytes per lane.
// ------------------------------------------------------------------------------------------------
// C_, H_, W_ > 0 fix the geometry at compile time (the two AlexNet instances); 0 = runtime values.
template <int LPP, int C_, int H_, int W_>
__global__ void __launch_bounds__(256, 3)
    pool_lrn_bwd_kernel2(const bf16* __restrict__ x, const bf16* __restrict__ dy, const uint8_t* __restrict__ arg,
                         bf16* __restrict__ dx, float* __restrict__ dbias, int n, int h_rt, int w_rt, int c_rt,
                         float alpha, float beta, float bias) {
  __shared__ float bsum[256];
  const int c = C_ ? C_ : c_rt, h = H_ ? H_ : h_rt, w = W_ ? W_ : w_rt;
  const int p = (h - 3) / 2 + 1, q = (w - 3) / 2 + 1;
  const int cpr = c >> 3;
  const int l = threadIdx.x % LPP;
  const int grp = threadIdx.x / LPP;
  constexpr int ngrp = 256 / LPP;
  const int c0 = l * 8;
  if (dbias != nullptr) {
    for (int i = threadIdx.x; i < c; i += blockDim.x) bsum[i] = 0.f;
    __syncthreads();
  }
  float bacc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) bacc[j] = 0.f;
  const float k2ab = 2.0f * alpha * beta;
  const int rows_total = n * h;
  for (int row = blockIdx.x; row < rows_total; row += gridDim.x) {
    const int nn = row / h;
    const int hh = row - nn * h;
    const int p_lo = max(0, (hh - 1) >> 1), p_hi = min(p - 1, hh >> 1);
    const bf16* xrow = x + (long long)row * (w * c);
    bf16* dxrow = dx + (long long)row * (w * c);
    const bf16* dyimg = dy + (long long)nn * (p * q * c);      // pooled tensors of this frame (32-bit offsets inside)
    const uint8_t* argimg = arg + (long long)nn * (p * q * c);
    for (int ww0 = 0; ww0 < w; ww0 += ngrp) {
      const int ww = ww0 + grp;
      const bool live = ww < w && l < cpr;
      float xv[8], gv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) xv[j] = gv[j] = 0.f;
      if (live) {
        const Bf16x8 xin = ld8(xrow + ww * c + c0);
        // dn: pooled gradient routed through the argmax codes, accumulated two channels per instruction in bf16
        // (exact for up to two contributions; the unfused path rounds the fp32 sum to bf16 once)
        __nv_bfloat162 g2[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) g2[i] = __floats2bfloat162_rn(0.f, 0.f);
        const int q_lo = max(0, (ww - 1) >> 1), q_hi = min(q - 1, ww >> 1);
        for (int pp = p_lo; pp <= p_hi; ++pp) {
          const int r3 = (hh - 2 * pp) * 3;  // window row 0..2 by construction of [p_lo, p_hi]
          for (int qq = q_lo; qq <= q_hi; ++qq) {
            const int o = (pp * q + qq) * c + c0;
            const uint32_t code4 = (uint32_t)(r3 + ww - 2 * qq) * 0x01010101u;
            const uint4 g = *reinterpret_cast<const uint4*>(dyimg + o);
            const uint2 a = *reinterpret_cast<const uint2*>(argimg + o);
            const uint32_t mlo = __vcmpeq4(a.x, code4), mhi = __vcmpeq4(a.y, code4);  // 0xff per matching channel
            const uint32_t gw[4] = {g.x & __byte_perm(mlo, 0, 0x1100), g.y & __byte_perm(mlo, 0, 0x3322),
                                    g.z & __byte_perm(mhi, 0, 0x1100), g.w & __byte_perm(mhi, 0, 0x3322)};
#pragma unroll
            for (int i = 0; i < 4; ++i) g2[i] = __hadd2(g2[i], *reinterpret_cast<const __nv_bfloat162*>(&gw[i]));
          }
        }
        unpack8(xin, xv);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 t = __bfloat1622float2(g2[i]);
          gv[2 * i] = t.x;
          gv[2 * i + 1] = t.y;
        }
      }
      float s[8];
      lrn_scale8<LPP>(xv, l, cpr, alpha, bias, s);
      float pw[8], tt[12];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float inv;
        if (beta == 0.75f) {
          const float rs = rsqrt_approx(s[j]);
          pw[j] = rs * sqrt_approx(rs);
          inv = rs * rs;
        } else {
          pw[j] = __powf(s[j], -beta);
          inv = __fdividef(1.0f, s[j]);
        }
        tt[2 + j] = (gv[j] * xv[j]) * (pw[j] * inv);
      }
      {
        float own[8], lf[2], rt[2];
#pragma unroll
        for (int j = 0; j < 8; ++j) own[j] = tt[2 + j];
        halo2<LPP>(own, l, cpr, lf, rt);
        tt[0] = lf[0];
        tt[1] = lf[1];
        tt[10] = rt[0];
        tt[11] = rt[1];
      }
      if (live) {
        float out[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float sum = ((tt[j] + tt[j + 1]) + tt[j + 2]) + (tt[j + 3] + tt[j + 4]);
          const float g = fmaf(gv[j], pw[j], -(k2ab * xv[j]) * sum);
          out[j] = xv[j] > 0.f ? g : 0.f;  // ReLU gradient of the producing conv
        }
        const Bf16x8 packed = pack8(out);
        st8(dxrow + ww * c + c0, packed);
        if (dbias != nullptr) {
          float rb[8];
          unpack8(packed, rb);  // the bias gradient sums the bf16 values that are stored (as vl_colsum would)
#pragma unroll
          for (int j = 0; j < 8; ++j) bacc[j] += rb[j];
        }
      }
    }
  }
  if (dbias != nullptr) {
    if (l < cpr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&bsum[c0 + j], bacc[j]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < c; i += blockDim.x) atomicAdd(dbias + i, bsum[i]);
  }
}

// ------------------------------------------------------------------------------------------------
// Fourth generation of the fused backward: uniform control flow and a packed pooled-gradient gather.
//
// A thread owns NE channels of a 2x2 PIXEL BLOCK (rows 2j, 2j+1 x columns 2b, 2b+1), LPP lanes hold the block's channel
// axis, a CTA walks (frame, row pair) units.  A 3x3/2 pooling window (p, q) covers rows 2p..2p+2 / columns 2q..2q+2,
// so the four pixels of a block see a FIXED list of windows and window positions:
//     (2j  , 2b  ): (j-1,b-1) code 8, (j-1,b) 6, (j,b-1) 2, (j,b) 0
//     (2j  , 2b+1): (j-1,b) 7, (j,b) 1           (2j+1, 2b  ): (j,b-1) 5, (j,b) 3           (2j+1, 2b+1): (j,b) 4
// i.e. every lane of every warp runs the same unrolled sequence (the per-pixel loops of the earlier kernels ran 1 or 2
// iterations depending on the column parity of the lane).  Windows outside the pooled tensor keep a clamped address and
// a code that never matches.  The gather itself works on bf16 pairs: the argmax byte b of a channel becomes the
// half-word b << 8 (one PRMT per pair; as bf16 these are distinct finite numbers), HSET2.EQ against the window
// position yields 1.0 / 0.0 and one HFMA2 accumulates dy — 2 instructions per channel pair and window instead of 3.5.
// The ReLU mask (x > 0) is applied to the gathered gradient in the packed domain; with dn = 0 every other term of the
// LRN gradient vanishes too because x = 0 there.
// ------------------------------------------------------------------------------------------------
template <int NE>
__device__ __forceinline__ void load_pairs(const bf16* __restrict__ p, uint32_t (&r)[NE / 2]) {
  if constexpr (NE % 8 == 0) {
#pragma unroll
    for (int k = 0; k < NE / 8; ++k) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(p) + k);
      r[4 * k] = u.x, r[4 * k + 1] = u.y, r[4 * k + 2] = u.z, r[4 * k + 3] = u.w;
    }
  } else {
#pragma unroll
    for (int k = 0; k < NE / 4; ++k) {
      const uint2 u = __ldg(reinterpret_cast<const uint2*>(p) + k);
      r[2 * k] = u.x, r[2 * k + 1] = u.y;
    }
  }
}
template <int NE>
__device__ __forceinline__ void store_pairs(bf16* __restrict__ p, const uint32_t (&r)[NE / 2]) {
  if constexpr (NE % 8 == 0) {
#pragma unroll
    for (int k = 0; k < NE / 8; ++k)
      reinterpret_cast<uint4*>(p)[k] = make_uint4(r[4 * k], r[4 * k + 1], r[4 * k + 2], r[4 * k + 3]);
  } else {
#pragma unroll
    for (int k = 0; k < NE / 4; ++k) reinterpret_cast<uint2*>(p)[k] = make_uint2(r[2 * k], r[2 * k + 1]);
  }
}
template <int NE>
__device__ __forceinline__ void load_bytes(const uint8_t* __restrict__ p, uint32_t (&r)[NE / 4]) {
  if constexpr (NE % 16 == 0) {
#pragma unroll
    for (int k = 0; k < NE / 16; ++k) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(p) + k);
      r[4 * k] = u.x, r[4 * k + 1] = u.y, r[4 * k + 2] = u.z, r[4 * k + 3] = u.w;
    }
  } else if constexpr (NE % 8 == 0) {
#pragma unroll
    for (int k = 0; k < NE / 8; ++k) {
      const uint2 u = __ldg(reinterpret_cast<const uint2*>(p) + k);
      r[2 * k] = u.x, r[2 * k + 1] = u.y;
    }
  } else {
#pragma unroll
    for (int k = 0; k < NE / 4; ++k) r[k] = __ldg(reinterpret_cast<const uint32_t*>(p) + k);
  }
}
__device__ __forceinline__ uint32_t heq2_one(uint32_t a, uint32_t b) {  // 1.0 / 0.0 per bf16 half
  uint32_t d;
  asm("set.eq.bf16x2.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ uint32_t hfma2_u(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ uint32_t hgt2_mask_zero(uint32_t a) {  // 0xffff per bf16 half that is > 0
  uint32_t d;
  asm("set.gt.u32.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(0u));
  return d;
}

// One pixel of the block: NWIN windows (pooled-gradient / argmax pointers already offset to the lane's channels).
template <int NE, int LPP, int NWIN>
__device__ __forceinline__ void bwd_pixel4(const bf16* __restrict__ xp, bf16* __restrict__ dxp, bool live,
                                           const bf16* const (&gp)[4], const uint8_t* const (&ap)[4],
                                           const uint32_t (&code2)[4], int l, float alpha, float bias, float k2ab,
                                           float (&bacc)[NE]) {
  uint32_t xr[NE / 2];
  if (live) {
    load_pairs<NE>(xp, xr);
  } else {
#pragma unroll
    for (int i = 0; i < NE / 2; ++i) xr[i] = 0u;
  }
  uint32_t dn2[NE / 2];
#pragma unroll
  for (int i = 0; i < NE / 2; ++i) dn2[i] = 0u;
#pragma unroll
  for (int wi = 0; wi < NWIN; ++wi) {
    uint32_t g[NE / 2], a[NE / 4];
    load_pairs<NE>(gp[wi], g);
    load_bytes<NE>(ap[wi], a);
#pragma unroll
    for (int k = 0; k < NE / 4; ++k) {
      const uint32_t h_lo = __byte_perm(a[k], 0u, 0x1404);  // (b0 << 8, b1 << 8)
      const uint32_t h_hi = __byte_perm(a[k], 0u, 0x3424);  // (b2 << 8, b3 << 8)
      dn2[2 * k] = hfma2_u(heq2_one(h_lo, code2[wi]), g[2 * k], dn2[2 * k]);
      dn2[2 * k + 1] = hfma2_u(heq2_one(h_hi, code2[wi]), g[2 * k + 1], dn2[2 * k + 1]);
    }
  }
  float xv[NE], av[NE];
#pragma unroll
  for (int i = 0; i < NE / 2; ++i) {
    const uint32_t d = dn2[i] & hgt2_mask_zero(xr[i]);  // ReLU gradient of the producing conv
    xv[2 * i] = __uint_as_float(xr[i] << 16);
    xv[2 * i + 1] = __uint_as_float(xr[i] & 0xffff0000u);
    av[2 * i] = __uint_as_float(d << 16);
    av[2 * i + 1] = __uint_as_float(d & 0xffff0000u);
  }
  float sq[NE], ssum[NE];
#pragma unroll
  for (int j = 0; j < NE; ++j) sq[j] = xv[j] * xv[j];
  window5n<LPP, NE>(sq, l, ssum);
  float tt[NE], tsum[NE];
#pragma unroll
  for (int j = 0; j < NE; ++j) {
    const float rs = rsqrt_approx(fmaf(alpha, ssum[j], bias));  // s^-1/2
    const float pw = rs * sqrt_approx(rs);                      // s^-3/4
    av[j] *= pw;                                                // dn * s^-beta
    tt[j] = av[j] * (xv[j] * (rs * rs));                        // dn * x * s^(-beta-1)
  }
  window5n<LPP, NE>(tt, l, tsum);
  uint32_t o[NE / 2];
#pragma unroll
  for (int i = 0; i < NE / 2; ++i) {
    const float d0 = fmaf(-(k2ab * xv[2 * i]), tsum[2 * i], av[2 * i]);
    const float d1 = fmaf(-(k2ab * xv[2 * i + 1]), tsum[2 * i + 1], av[2 * i + 1]);
    bacc[2 * i] += d0;
    bacc[2 * i + 1] += d1;
    const __nv_bfloat162 pk = __floats2bfloat162_rn(d0, d1);
    o[i] = *reinterpret_cast<const uint32_t*>(&pk);
  }
  if (live) store_pairs<NE>(dxp, o);
}

template <int NE, int LPP, int C_, int H_, int W_, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
    pool_lrn_bwd_kernel4(const bf16* __restrict__ x, const bf16* __restrict__ dy, const uint8_t* __restrict__ arg,
                         bf16* __restrict__ dx, float* __restrict__ dbias, int n, float alpha, float bias, int pf) {
  constexpr int P = (H_ - 3) / 2 + 1, Q = (W_ - 3) / 2 + 1;
  constexpr int JP = (H_ + 1) / 2, CP = (W_ + 1) / 2;  // row pairs per frame, column pairs per row
  static_assert(LPP * NE == C_, "lanes x channels per lane must cover the channel axis exactly");
  static_assert((THREADS / LPP) >= CP, "a CTA must hold one row of pixel blocks");
  __shared__ float bsum[C_];
  const int l = threadIdx.x % LPP;
  const int b = threadIdx.x / LPP;
  const int c0 = l * NE;
  if (dbias != nullptr) {
    for (int i = threadIdx.x; i < C_; i += THREADS) bsum[i] = 0.f;
    __syncthreads();
  }
  float bacc[NE];
#pragma unroll
  for (int j = 0; j < NE; ++j) bacc[j] = 0.f;
  const float k2ab = 2.0f * alpha * 0.75f;
  const int col0 = 2 * b, col1 = 2 * b + 1;
  const bool live0 = col0 < W_, live1 = col1 < W_;
  // pooled columns of the block: b-1 (window column 2) and b (window columns 0 / 1)
  const bool qa_ok = b >= 1 && b - 1 < Q, qb_ok = b < Q;
  const int qa = min(max(b - 1, 0), Q - 1), qb = min(b, Q - 1);
  constexpr uint32_t DEAD = 0xFF00FF00u;
#define VL_CODE2(k) ((uint32_t)(k) * 0x01000100u)
  const int units = n * JP;
  for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
    const int nn = unit / JP;
    const int j = unit - nn * JP;
    const int row0 = 2 * j, row1 = 2 * j + 1;
    const bool pa_ok = j >= 1 && j - 1 < P, pb_ok = j < P;  // pooled rows j-1 (window row 2) and j (rows 0 / 1)
    const int pa = min(max(j - 1, 0), P - 1), pb = min(j, P - 1);
    const long long pooled = (long long)nn * (P * Q * C_) + c0;
    const int o_aa = (pa * Q + qa) * C_, o_ab = (pa * Q + qb) * C_, o_ba = (pb * Q + qa) * C_, o_bb = (pb * Q + qb) * C_;
    const bf16* g_aa = dy + pooled + o_aa;
    const bf16* g_ab = dy + pooled + o_ab;
    const bf16* g_ba = dy + pooled + o_ba;
    const bf16* g_bb = dy + pooled + o_bb;
    const uint8_t* a_aa = arg + pooled + o_aa;
    const uint8_t* a_ab = arg + pooled + o_ab;
    const uint8_t* a_ba = arg + pooled + o_ba;
    const uint8_t* a_bb = arg + pooled + o_bb;
    const long long r0 = ((long long)nn * H_ + row0) * (W_ * C_) + c0;
    if (pf && unit + pf * (int)gridDim.x < units) {
      // pull the lines of this CTA's NEXT unit into L2 (no registers held): the kernel's top stall is the latency of its
      // global loads (long scoreboard, profiles/r02_ncu_lrn_v4_summary.csv)
      const int u2 = unit + pf * (int)gridDim.x;
      const int n2 = u2 / JP, j2 = u2 - n2 * JP;
      const long long r2 = ((long long)n2 * H_ + 2 * j2) * (W_ * C_) + c0;
      const int pb2 = min(j2, P - 1), pa2 = min(max(j2 - 1, 0), P - 1);
      const long long pool2 = (long long)n2 * (P * Q * C_) + c0;
      auto pfl = [](const void* ptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr)); };
      if (live0) {
        pfl(x + r2 + col0 * C_);
        if (2 * j2 + 1 < H_) pfl(x + r2 + (W_ * C_) + col0 * C_);
      }
      if (live1) {
        pfl(x + r2 + col1 * C_);
        if (2 * j2 + 1 < H_) pfl(x + r2 + (W_ * C_) + col1 * C_);
      }
      pfl(dy + pool2 + (pb2 * Q + qb) * C_);
      pfl(arg + pool2 + (pb2 * Q + qb) * C_);
      pfl(dy + pool2 + (pa2 * Q + qb) * C_);
      pfl(arg + pool2 + (pa2 * Q + qb) * C_);
    }
    {  // (row0, col0): four windows
      const bf16* const gp[4] = {g_aa, g_ab, g_ba, g_bb};
      const uint8_t* const ap[4] = {a_aa, a_ab, a_ba, a_bb};
      const uint32_t cd[4] = {pa_ok && qa_ok ? VL_CODE2(8) : DEAD, pa_ok && qb_ok ? VL_CODE2(6) : DEAD,
                              pb_ok && qa_ok ? VL_CODE2(2) : DEAD, pb_ok && qb_ok ? VL_CODE2(0) : DEAD};
      bwd_pixel4<NE, LPP, 4>(x + r0 + col0 * C_, dx + r0 + col0 * C_, live0, gp, ap, cd, l, alpha, bias, k2ab, bacc);
    }
    {  // (row0, col1): two windows
      const bf16* const gp[4] = {g_ab, g_bb, g_bb, g_bb};
      const uint8_t* const ap[4] = {a_ab, a_bb, a_bb, a_bb};
      const uint32_t cd[4] = {pa_ok && qb_ok ? VL_CODE2(7) : DEAD, pb_ok && qb_ok ? VL_CODE2(1) : DEAD, DEAD, DEAD};
      bwd_pixel4<NE, LPP, 2>(x + r0 + col1 * C_, dx + r0 + col1 * C_, live1, gp, ap, cd, l, alpha, bias, k2ab, bacc);
    }
    if (row1 < H_) {  // uniform per CTA
      const long long r1 = r0 + (W_ * C_);
      {  // (row1, col0): two windows
        const bf16* const gp[4] = {g_ba, g_bb, g_bb, g_bb};
        const uint8_t* const ap[4] = {a_ba, a_bb, a_bb, a_bb};
        const uint32_t cd[4] = {pb_ok && qa_ok ? VL_CODE2(5) : DEAD, pb_ok && qb_ok ? VL_CODE2(3) : DEAD, DEAD, DEAD};
        bwd_pixel4<NE, LPP, 2>(x + r1 + col0 * C_, dx + r1 + col0 * C_, live0, gp, ap, cd, l, alpha, bias, k2ab, bacc);
      }
      {  // (row1, col1): one window
        const bf16* const gp[4] = {g_bb, g_bb, g_bb, g_bb};
        const uint8_t* const ap[4] = {a_bb, a_bb, a_bb, a_bb};
        const uint32_t cd[4] = {pb_ok && qb_ok ? VL_CODE2(4) : DEAD, DEAD, DEAD, DEAD};
        bwd_pixel4<NE, LPP, 1>(x + r1 + col1 * C_, dx + r1 + col1 * C_, live1, gp, ap, cd, l, alpha, bias, k2ab, bacc);
      }
    }
  }
#undef VL_CODE2
  if (dbias != nullptr) {
    // dead lanes (blocks beyond the last column pair) accumulated exact zeros: x = 0 -> dn masked -> dx = 0
#pragma unroll
    for (int j = 0; j < NE; ++j) atomicAdd(&bsum[c0 + j], bacc[j]);
    __syncthreads();
    for (int i = threadIdx.x; i < C_; i += THREADS) atomicAdd(dbias + i, bsum[i]);
  }
}


// ------------------------------------------------------------------------------------------------
// Fourth generation of the fused forward: registers and warp shuffles only - no shared memory, no CTA barrier (the top
// stall of the row-ring kernel), every lrn(x) evaluated (almost) once, separable pooling.
//
// A WARP walks down the image rows of a strip of input columns.  A group of LPP lanes holds one pixel's channel slice
// (NE channels per lane, LPP * NE per slice, SLICES slices per pixel: the channel halo of a slice's outer lanes is read
// straight from global memory, the inner halos travel by shuffle); group u of the warp owns the input column pair
// (2q, 2q+1) of pooled column q = U * cb + u and receives column 2q+2 - the first column of group u+1, already normalised
// and packed - by one shuffle per bf16 pair.  The last group (u = G-1) only feeds its left neighbour, so U = G-1 of the
// G groups produce output (for the two AlexNet geometries: Q = 28 = 4 x 7 and Q = 13 <= 2 x 7).  Per input row the
// three-column maximum and its column code are formed once (horizontal pass); the vertical pass merges the rows 2p, 2p+1,
// 2p+2 of pooled row p and row 2p+2 is carried over as row 0 of pooled row p+1.  Strict '>' in (row, column) scan order
// keeps TF's first-maximum rule.  The next row's pixels are fetched before the current row is processed.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t gt_bf16x2(uint32_t a, uint32_t b) {
  uint32_t m;
  asm("set.gt.u32.bf16x2 %0, %1, %2;" : "=r"(m) : "r"(a), "r"(b));
  return m;
}
__device__ __forceinline__ uint32_t max_bf16x2(uint32_t a, uint32_t b) {
  uint32_t m;
  asm("max.bf16x2 %0, %1, %2;" : "=r"(m) : "r"(a), "r"(b));
  return m;
}

template <int NE, int LPP, int SLICES, int C_, int H_, int W_>
__global__ void __launch_bounds__(128, 4)
    lrn_pool_fwd_kernel4(const bf16* __restrict__ x, bf16* __restrict__ y, uint8_t* __restrict__ arg, int n, int seg_rows,
                         int segs, float alpha, float bias) {
  constexpr int P = (H_ - 3) / 2 + 1, Q = (W_ - 3) / 2 + 1;
  constexpr int G = 32 / LPP, U = G - 1;
  constexpr int WPR = (Q + U - 1) / U;  // column blocks (warps) per image row
  constexpr int CS = LPP * NE;          // channels of a slice
  constexpr int NP = NE / 2;            // bf16 pairs per lane
  constexpr int ROW = W_ * C_;
  static_assert(CS * SLICES == C_ && NE % 8 == 0, "slices x lanes x channels per lane must cover the channel axis");
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int u = lane / LPP, l = lane % LPP;
  const long long items = (long long)n * segs * SLICES * WPR;
  for (long long item = (long long)blockIdx.x * 4 + wib; item < items; item += (long long)gridDim.x * 4) {
    long long rest = item;
    const int cb = (int)(rest % WPR);
    rest /= WPR;
    const int sl = (int)(rest % SLICES);
    rest /= SLICES;
    const int sg = (int)(rest % segs);
    const int nn = (int)(rest / segs);
    const int p0 = sg * seg_rows, p1 = min(P, p0 + seg_rows);
    if (p0 >= p1) continue;
    const int q = U * cb + u;  // pooled column of this group
    const int col0 = 2 * q;    // owned input columns col0, col0 + 1
    const bool ok0 = col0 < W_, ok1 = col0 + 1 < W_;
    const bool out_ok = u < U && q < Q;
    const int ch0 = sl * CS + l * NE;
    const bool halo_l = SLICES > 1 && l == 0 && sl > 0;
    const bool halo_r = SLICES > 1 && l == LPP - 1 && sl < SLICES - 1;
    const bf16* xcol = x + (long long)nn * (H_ * ROW) + (long long)col0 * C_ + ch0;

    // raw pixels of one input row: the two owned columns (+ the packed channel halo pairs of the slice's outer lanes)
    struct RowRegs {
      uint32_t px[2][NP];
      uint32_t hl[2], hr[2];
    };
    auto fetch = [&](int r, RowRegs& t) {
      const bf16* rp = xcol + (long long)r * ROW;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const bool ok = c == 0 ? ok0 : ok1;
        if (ok) {
          load_pairs<NE>(rp + c * C_, t.px[c]);
        } else {
#pragma unroll
          for (int i = 0; i < NP; ++i) t.px[c][i] = 0u;
        }
        t.hl[c] = (halo_l && ok) ? __ldg(reinterpret_cast<const uint32_t*>(rp + c * C_ - 2)) : 0u;
        t.hr[c] = (halo_r && ok) ? __ldg(reinterpret_cast<const uint32_t*>(rp + c * C_ + NE)) : 0u;
      }
    };
    // lrn of one pixel slice -> packed bf16 pairs
    auto normalise = [&](const uint32_t (&xr)[NP], uint32_t hl, uint32_t hr, uint32_t (&o)[NP]) {
      float v[NE], e[NE + 4];
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        v[2 * i] = __uint_as_float(xr[i] << 16);
        v[2 * i + 1] = __uint_as_float(xr[i] & 0xffff0000u);
      }
#pragma unroll
      for (int j = 0; j < NE; ++j) e[2 + j] = v[j] * v[j];
      float lf0 = __shfl_up_sync(0xffffffffu, e[NE], 1, LPP);
      float lf1 = __shfl_up_sync(0xffffffffu, e[NE + 1], 1, LPP);
      float rt0 = __shfl_down_sync(0xffffffffu, e[2], 1, LPP);
      float rt1 = __shfl_down_sync(0xffffffffu, e[3], 1, LPP);
      if (l == 0) {
        const float a = __uint_as_float(hl << 16), b = __uint_as_float(hl & 0xffff0000u);
        lf0 = a * a;
        lf1 = b * b;
      }
      if (l == LPP - 1) {
        const float a = __uint_as_float(hr << 16), b = __uint_as_float(hr & 0xffff0000u);
        rt0 = a * a;
        rt1 = b * b;
      }
      e[0] = lf0, e[1] = lf1, e[NE + 2] = rt0, e[NE + 3] = rt1;
      // window sums with 2 adds per channel and no cancellation: for even j, q4 = e[j+1..j+4] (two pair sums at odd
      // offsets, each shared by two q4), w[j] = e[j] + q4, w[j+1] = q4 + e[j+5]
      float pr[NP + 1];
#pragma unroll
      for (int k = 0; k <= NP; ++k) pr[k] = e[2 * k + 1] + e[2 * k + 2];
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        const float q4 = pr[i] + pr[i + 1];
        const float w0 = e[2 * i] + q4, w1 = q4 + e[2 * i + 5];
        const float r0s = rsqrt_approx(fmaf(alpha, w0, bias));
        const float r1s = rsqrt_approx(fmaf(alpha, w1, bias));
        const __nv_bfloat162 pk =
            __floats2bfloat162_rn(v[2 * i] * (r0s * sqrt_approx(r0s)), v[2 * i + 1] * (r1s * sqrt_approx(r1s)));
        o[i] = *reinterpret_cast<const uint32_t*>(&pk);
      }
    };
    // one input row: normalise the two owned columns, take column 2q+2 from the next group, three-column maximum + code
    auto hrow = [&](const RowRegs& t, uint32_t (&hb)[NP], uint32_t (&hc)[NP]) {
      uint32_t n0[NP], n1[NP];
      normalise(t.px[0], t.hl[0], t.hr[0], n0);
      normalise(t.px[1], t.hl[1], t.hr[1], n1);
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        const uint32_t n2 = __shfl_down_sync(0xffffffffu, n0[i], LPP);
        const uint32_t m1 = gt_bf16x2(n1[i], n0[i]);
        uint32_t b = max_bf16x2(n0[i], n1[i]);
        uint32_t c = m1 & 0x00010001u;
        const uint32_t m2 = gt_bf16x2(n2, b);
        b = max_bf16x2(b, n2);
        c = (c & ~m2) | (0x00020002u & m2);
        hb[i] = b, hc[i] = c;
      }
    };

    // two row buffers used alternately (no register copies): at the top of iteration p, ra holds input row 2p+1
    RowRegs ra, rb;
    uint32_t ab[NP], ac[NP], hb[NP], hc[NP];
    fetch(2 * p0, rb);
    fetch(2 * p0 + 1, ra);
    hrow(rb, ab, ac);  // window row 0 of pooled row p0
    for (int p = p0; p < p1; ++p) {
      // window row 1
      fetch(2 * p + 2, rb);
      hrow(ra, hb, hc);
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        const uint32_t m = gt_bf16x2(hb[i], ab[i]);
        ab[i] = max_bf16x2(ab[i], hb[i]);
        ac[i] = (ac[i] & ~m) | ((hc[i] + 0x00030003u) & m);
      }
      // window row 2 (= window row 0 of pooled row p + 1)
      if (p + 1 < p1) fetch(2 * p + 3, ra);
      hrow(rb, hb, hc);
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        const uint32_t m = gt_bf16x2(hb[i], ab[i]);
        ab[i] = max_bf16x2(ab[i], hb[i]);
        ac[i] = (ac[i] & ~m) | ((hc[i] + 0x00060006u) & m);
      }
      if (out_ok) {
        const long long opix = (((long long)nn * P + p) * Q + q) * C_ + ch0;
        store_pairs<NE>(y + opix, ab);
#pragma unroll
        for (int k = 0; k < NE / 8; ++k) {
          const uint32_t lo = __byte_perm(ac[4 * k], ac[4 * k + 1], 0x6420);
          const uint32_t hi = __byte_perm(ac[4 * k + 2], ac[4 * k + 3], 0x6420);
          *reinterpret_cast<uint2*>(arg + opix + 8 * k) = make_uint2(lo, hi);
        }
      }
#pragma unroll
      for (int i = 0; i < NP; ++i) ab[i] = hb[i], ac[i] = hc[i];
    }
  }
}

}  // namespace

#define VL_LAUNCHED()                  \
  do {                                 \
    vl::g_launches.fetch_add(1);       \
    VL_CHECK_CUDA(cudaGetLastError()); \
  } while (0)

extern "C" int vl_frames_s2d_crop(const void* frames, int32_t is_u8, const float* mean3, void* out, int32_t n,
                                  int32_t hr, int32_t wr, const int32_t* crops, int32_t h, int32_t w, int32_t s,
                                  int32_t pad_top, int32_t pad_left, int32_t hb, int32_t wb, vl_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VL_REQUIRE(frames && out && hb > 0 && wb > 0, "vl_frames_s2d: bad arguments");
  VL_REQUIRE(s == 4, "vl_frames_s2d: only stride 4 (conv1, alexnet.py:76) is instantiated");
  VL_REQUIRE((wb * s * s * 3) % 8 == 0, "vl_frames_s2d: wb*s*s*3 must be a multiple of 8");
  VL_REQUIRE((w + pad_left) * 3 <= wb * s * 3, "vl_frames_s2d: image row does not fit the block row");
  VL_REQUIRE(hr >= h && wr >= w, "vl_frames_s2d: stored frame %dx%d smaller than the network input %dx%d", hr, wr, h, w);
  VL_REQUIRE(crops != nullptr || (hr == h && wr == w), "vl_frames_s2d: crop offsets are required when the stored frame is larger");
  const size_t smem = (size_t)wb * s * s * 3 * sizeof(bf16);
  VL_REQUIRE(smem <= 48 * 1024, "vl_frames_s2d: image row too wide (%zu bytes of shared memory)", smem);
  if (is_u8 && crops == nullptr && pad_left % 4 == 0 && pad_top >= 0 && wb <= 64) {
    const int n_rows = n * hb;
    const int want = (n_rows + 1) / 2;
    const int grid_d = want < vl::num_sms() * 16 ? want : vl::num_sms() * 16;
    frames_s2d_direct_kernel<<<grid_d, 128, 0, stream>>>(reinterpret_cast<const uint8_t*>(frames), mean3,
                                                        reinterpret_cast<bf16*>(out), h, w, pad_top, pad_left, hb, wb,
                                                        (long long)n * h * w * 3, n_rows,
                                                        getenv("VL_S2D_WIDE") ? atoi(getenv("VL_S2D_WIDE")) : 1);
    VL_LAUNCHED();
    return 0;
  }
  const int grid = n * hb;
  if (is_u8)
    frames_s2d_kernel<true, 4><<<grid, 256, smem, stream>>>(frames, mean3, reinterpret_cast<bf16*>(out), h, w, pad_top,
                                                            pad_left, hb, wb, hr, wr, crops);
  else
    frames_s2d_kernel<false, 4><<<grid, 256, smem, stream>>>(frames, mean3, reinterpret_cast<bf16*>(out), h, w, pad_top,
                                                             pad_left, hb, wb, hr, wr, crops);
  VL_LAUNCHED();
  return 0;
}

extern "C" int vl_frames_s2d(const void* frames, int32_t is_u8, const float* mean3, void* out, int32_t n, int32_t h,
                             int32_t w, int32_t s, int32_t pad_top, int32_t pad_left, int32_t hb, int32_t wb,
                             vl_stream_t stream_) {
  return vl_frames_s2d_crop(frames, is_u8, mean3, out, n, h, w, nullptr, h, w, s, pad_top, pad_left, hb, wb, stream_);
}

extern "C" int vl_s2d_pack_filter(const float* src, void* dst, int32_t kh, int32_t kw, int32_t cin, int32_t cout,
                                  int32_t s, int32_t chunk, int32_t transpose, vl_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VL_REQUIRE(src && dst && s >= 1 && chunk >= s * s * cin, "vl_s2d_pack_filter: bad arguments");
  const int kb_h = (kh + s - 1) / s, kb_w = (kw + s - 1) / s;
  const long long total = (long long)kb_h * kb_w * chunk * cout;
  s2d_pack_filter_kernel<<<(int)((total + 255) / 256), 256, 0, stream>>>(src, reinterpret_cast<bf16*>(dst), kh, kw, cin,
                                                                         cout, s, kb_h, kb_w, chunk, transpose, total);
  VL_LAUNCHED();
  return 0;
}

extern "C" int vl_s2d_unpack_grad(const float* dws, float* dw, int32_t kh, int32_t kw, int32_t cin, int32_t cout,
                                  int32_t s, vl_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VL_REQUIRE(dws && dw && s >= 1, "vl_s2d_unpack_grad: bad arguments");
  const int kb_w = (kw + s - 1) / s;
  const long long total = (long long)kh * kw * cin * cout;
  s2d_unpack_grad_kernel<<<(int)((total + 255) / 256), 256, 0, stream>>>(dws, dw, kh, kw, cin, cout, s, kb_w, total);
  VL_LAUNCHED();
  return 0;
}

// Generic (any c % 8 == 0, radius <= 4) versions live in encoder_mem.cu.
extern "C" int vl_lrn_pool_fwd_generic(const void* x, void* y, void* argmax, int32_t n, int32_t h, int32_t w, int32_t c,
                                       int32_t radius, float alpha, float beta, float bias, vl_stream_t stream_);
extern "C" int vl_pool_lrn_bwd_generic(const void* x, const void* dy, const void* argmax, void* dx, float* dbias,
                                       int32_t n, int32_t h, int32_t w, int32_t c, int32_t radius, float alpha,
                                       float beta, float bias, vl_stream_t stream_);

extern "C" int vl_lrn_pool_fwd(const void* x, void* y, void* argmax, int32_t n, int32_t h, int32_t w, int32_t c,
                               int32_t radius, float alpha, float beta, float bias, vl_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VL_REQUIRE(x && y && argmax && c % 8 == 0 && h >= 3 && w >= 3, "vl_lrn_pool_fwd: bad arguments");
  VL_REQUIRE(radius == 2 && beta == 0.75f, "vl_lrn_pool_fwd: fused path serves depth_radius 2, beta 0.75 (alexnet.py:80-89)");
  const int p = (h - 3) / 2 + 1, q = (w - 3) / 2 + 1;
  const size_t row_bytes = (size_t)w * c * sizeof(bf16);
  // strip height: small strips -> more CTAs (warps) per SM to hide the load latency, at the price of re-evaluating
  // the LRN of the rows shared by neighbouring strips
  const size_t strip_limit = (getenv("VL_LRN_STRIP_KB") ? atoi(getenv("VL_LRN_STRIP_KB")) : 100) * 1024;
  int rows_out = 0;
  for (int r = 1; r <= p; ++r)
    if ((size_t)(2 * r + 1) * row_bytes <= strip_limit) rows_out = r;
  // third generation (row ring, compile-time geometry): the two AlexNet instances
  // fourth generation (registers + shuffles, no shared memory): the two AlexNet instances
  if (!getenv("VL_LRN_FWD_V2") && ((c == 96 && h == 57 && w == 57) || (c == 256 && h == 28 && w == 28))) {
    const int segs = c == 96 ? 4 : 2;  // row segments per image: one re-normalised input row per extra segment
    const int seg_rows = (p + segs - 1) / segs;
    const int wpr = c == 96 ? 4 : 2, slices = c == 96 ? 1 : 4;
    const long long items = (long long)n * segs * slices * wpr;
    const long long want = (items + 3) / 4;
    const long long cap = (long long)vl::num_sms() * (getenv("VL_LRN_FWD_CTAS") ? atoi(getenv("VL_LRN_FWD_CTAS")) : 16);
    const int g = (int)(want < cap ? want : cap);
    // (measured, profiles/r02_lrn_fwd4.txt: 5 / 6 CTAs per SM only by spilling - slower; a grid of 16 CTAs per SM lets
    // the block scheduler even out the tail)
    if (c == 96)
      lrn_pool_fwd_kernel4<24, 4, 1, 96, 57, 57><<<g, 128, 0, stream>>>(
          reinterpret_cast<const bf16*>(x), reinterpret_cast<bf16*>(y), reinterpret_cast<uint8_t*>(argmax), n, seg_rows, segs,
          alpha, bias);
    else
      lrn_pool_fwd_kernel4<16, 4, 4, 256, 28, 28><<<g, 128, 0, stream>>>(
          reinterpret_cast<const bf16*>(x), reinterpret_cast<bf16*>(y), reinterpret_cast<uint8_t*>(argmax), n, seg_rows, segs,
          alpha, bias);
    VL_LAUNCHED();
    return 0;
  }
  if (c > 256 || rows_out == 0)
    return vl_lrn_pool_fwd_generic(x, y, argmax, n, h, w, c, radius, alpha, beta, bias, stream_);
  const int strips = (p + rows_out - 1) / rows_out;
  rows_out = (p + strips - 1) / strips;  // balance the strips
  const size_t smem = (size_t)(2 * rows_out + 1) * row_bytes;
  const int grid = n * strips;
#define VL_FWD_LAUNCH(LPP_, CPL_)                                                                                       \
  do {                                                                                                                 \
    static bool attr = false;                                                                                          \
    if (!attr) {                                                                                                       \
      VL_CHECK_CUDA(cudaFuncSetAttribute(lrn_pool_fwd_kernel2<LPP_, CPL_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         112 * 1024));                                                                 \
      attr = true;                                                                                                     \
    }                                                                                                                  \
    lrn_pool_fwd_kernel2<LPP_, CPL_><<<grid, 512, smem, stream>>>(                                                      \
        reinterpret_cast<const bf16*>(x), reinterpret_cast<bf16*>(y), reinterpret_cast<uint8_t*>(argmax), h, w, c, p, \
        q, rows_out, strips, alpha, bias);                                                                             \
  } while (0)
  if (c == 96)
    VL_FWD_LAUNCH(4, 3);
  else if (c == 256)
    VL_FWD_LAUNCH(32, 1);
  else if (c == 128)
    VL_FWD_LAUNCH(16, 1);
  else if (c == 64)
    VL_FWD_LAUNCH(8, 1);
  else
    return vl_lrn_pool_fwd_generic(x, y, argmax, n, h, w, c, radius, alpha, beta, bias, stream_);
#undef VL_FWD_LAUNCH
  VL_LAUNCHED();
  return 0;
}

extern "C" int vl_pool_lrn_bwd(const void* x, const void* dy, const void* argmax, void* dx, float* dbias, int32_t n,
                               int32_t h, int32_t w, int32_t c, int32_t radius, float alpha, float beta, float bias,
                               vl_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VL_REQUIRE(x && dy && argmax && dx && c % 8 == 0, "vl_pool_lrn_bwd: bad arguments");
  VL_REQUIRE(radius == 2, "vl_pool_lrn_bwd: only depth_radius 2 (alexnet.py:80,121) is implemented");
  if (c > 256) return vl_pool_lrn_bwd_generic(x, dy, argmax, dx, dbias, n, h, w, c, radius, alpha, beta, bias, stream_);
  const int p = (h - 3) / 2 + 1, q = (w - 3) / 2 + 1;
  const int lpp = c <= 128 ? 16 : 32;
  long long blocks = (long long)n * h;
  const long long cap = (long long)vl::num_sms() * 6;
  if (blocks > cap) blocks = cap;
#define VL_BWD_ARGS                                                                                                   \
  reinterpret_cast<const bf16*>(x), reinterpret_cast<const bf16*>(dy), reinterpret_cast<const uint8_t*>(argmax),    \
      reinterpret_cast<bf16*>(dx), dbias, n, h, w, c, alpha, beta, bias
  (void)p;
  (void)q;
  // fourth generation (2x2 pixel blocks, uniform control flow): the two AlexNet geometries, beta = 0.75.  Four CTAs
  // per SM in the grid; 96 channels: 256 threads, two CTAs resident; 256 channels: 448 threads at 72 registers, two
  // CTAs resident (243 us against 293 us with one CTA of 128 registers)
  const long long gmul = 4;
  const int pf = getenv("VL_LRN_BWD_PF") ? atoi(getenv("VL_LRN_BWD_PF")) : 1;
  if (beta == 0.75f && c == 96 && h == 57 && w == 57) {
    const long long units = (long long)n * 29;
    const long long g = units < (long long)vl::num_sms() * gmul ? units : (long long)vl::num_sms() * gmul;
    pool_lrn_bwd_kernel4<12, 8, 96, 57, 57, 256, 2><<<(int)g, 256, 0, stream>>>(
        reinterpret_cast<const bf16*>(x), reinterpret_cast<const bf16*>(dy), reinterpret_cast<const uint8_t*>(argmax),
        reinterpret_cast<bf16*>(dx), dbias, n, alpha, bias, pf);
    VL_LAUNCHED();
    return 0;
  }
  if (beta == 0.75f && c == 256 && h == 28 && w == 28) {
    const long long units = (long long)n * 14;
    const long long g = units < (long long)vl::num_sms() * gmul ? units : (long long)vl::num_sms() * gmul;
    pool_lrn_bwd_kernel4<8, 32, 256, 28, 28, 448, 2><<<(int)g, 448, 0, stream>>>(
        reinterpret_cast<const bf16*>(x), reinterpret_cast<const bf16*>(dy), reinterpret_cast<const uint8_t*>(argmax),
        reinterpret_cast<bf16*>(dx), dbias, n, alpha, bias, pf);
    VL_LAUNCHED();
    return 0;
  }
  // every other geometry (c <= 256, c % 8 == 0): the per-pixel kernel with run-time extents
  if (lpp == 16)
    pool_lrn_bwd_kernel2<16, 0, 0, 0><<<(int)blocks, 256, 0, stream>>>(VL_BWD_ARGS);
  else
    pool_lrn_bwd_kernel2<32, 0, 0, 0><<<(int)blocks, 256, 0, stream>>>(VL_BWD_ARGS);
#undef VL_BWD_ARGS
  VL_LAUNCHED();
  return 0;
}
