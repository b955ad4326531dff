// Read-time image resampling of the reference on the device (sm_100a): `scipy.misc.imresize(image, shape)` with its
// default bilinear filter (dataset_.py:238,484,491; serialize.py:425) = Pillow's ImagingResample for 8-bit channels:
// two separable passes (horizontal first), fixed-point coefficients (22 fractional bits) computed on the host exactly
// as Pillow does, every pass rounded to uint8.  Integer arithmetic -> bit-exact against PIL
// (tests/golden/resize_bilinear_golden.npz).  HBM-bound byte work: one thread per output pixel (3 channels), the taps
// of a pixel walk contiguous bytes (horizontal) or a row stride (vertical); frames stay uint8 until the staging
// kernel (vl_frames_s2d_crop) crops / mirrors / mean-subtracts them.
#include "common.cuh"
#include "../../include/vlb200.h"

#include <atomic>

namespace vl {
extern std::atomic<long long> g_launches;
}

namespace {

constexpr int PRECISION_BITS = 32 - 8 - 2;

__device__ __forceinline__ uint8_t clip8(int v) {
  v >>= PRECISION_BITS;
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// out[n][y][xx][c] = clip8(2^21 + sum_k in[n][y][xmin(xx) + k][c] * coeff[xx][k]),  c < CH
template <int CH>
__global__ void resize_h_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, long long rows, int w_in,
                                int w_out, const int32_t* __restrict__ bounds, const int32_t* __restrict__ coeffs,
                                int ksize) {
  const long long total = rows * w_out;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int xx = (int)(idx % w_out);
    const long long row = idx / w_out;
    const int xmin = __ldg(bounds + 2 * xx), cnt = __ldg(bounds + 2 * xx + 1);
    const uint8_t* src = in + (row * w_in + xmin) * CH;
    const int32_t* kk = coeffs + (long long)xx * ksize;
    int acc[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) acc[c] = 1 << (PRECISION_BITS - 1);
    for (int k = 0; k < cnt; ++k) {
      const int w = __ldg(kk + k);
#pragma unroll
      for (int c = 0; c < CH; ++c) acc[c] += (int)__ldg(src + k * CH + c) * w;
    }
    uint8_t* dst = out + idx * CH;
#pragma unroll
    for (int c = 0; c < CH; ++c) dst[c] = clip8(acc[c]);
  }
}

// out[n][yy][x][c] = clip8(2^21 + sum_k in[n][ymin(yy) + k][x][c] * coeff[yy][k]); a thread owns one byte column
__global__ void resize_v_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int n, int h_in, int h_out,
                                int row_bytes, const int32_t* __restrict__ bounds, const int32_t* __restrict__ coeffs,
                                int ksize) {
  const long long total = (long long)n * h_out * row_bytes;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int xb = (int)(idx % row_bytes);
    const long long t = idx / row_bytes;
    const int yy = (int)(t % h_out);
    const long long img = t / h_out;
    const int ymin = __ldg(bounds + 2 * yy), cnt = __ldg(bounds + 2 * yy + 1);
    const uint8_t* src = in + (img * h_in + ymin) * row_bytes + xb;
    const int32_t* kk = coeffs + (long long)yy * ksize;
    int acc = 1 << (PRECISION_BITS - 1);
    for (int k = 0; k < cnt; ++k) acc += (int)__ldg(src + (long long)k * row_bytes) * __ldg(kk + k);
    out[idx] = clip8(acc);
  }
}

int grid_for(long long work) {
  long long g = (work + 255) / 256;
  const long long cap = (long long)vl::num_sms() * 16;
  return (int)(g > cap ? cap : (g < 1 ? 1 : g));
}

}  // namespace

extern "C" int vl_resize_bilinear_u8(const void* in, void* out, void* tmp, int32_t n, int32_t h_in, int32_t w_in,
                                     int32_t h_out, int32_t w_out, int32_t channels, const int32_t* bounds_w,
                                     const int32_t* coeffs_w, int32_t ksize_w, const int32_t* bounds_h,
                                     const int32_t* coeffs_h, int32_t ksize_h, vl_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VL_REQUIRE(in && out && n > 0 && h_in > 0 && w_in > 0 && h_out > 0 && w_out > 0, "vl_resize_bilinear_u8: bad arguments");
  VL_REQUIRE(channels == 3 || channels == 1, "vl_resize_bilinear_u8: 1 or 3 channels (got %d)", channels);
  const bool need_w = w_out != w_in, need_h = h_out != h_in;
  VL_REQUIRE(!need_w || (bounds_w && coeffs_w && ksize_w > 0), "vl_resize_bilinear_u8: horizontal coefficients missing");
  VL_REQUIRE(!need_h || (bounds_h && coeffs_h && ksize_h > 0), "vl_resize_bilinear_u8: vertical coefficients missing");
  VL_REQUIRE(!(need_w && need_h) || tmp, "vl_resize_bilinear_u8: a two-pass resize needs the [n][h_in][w_out][c] scratch");
  if (!need_w && !need_h) {
    VL_CHECK_CUDA(cudaMemcpyAsync(out, in, (size_t)n * h_in * w_in * channels, cudaMemcpyDeviceToDevice, stream));
    return 0;
  }
  const uint8_t* src = reinterpret_cast<const uint8_t*>(in);
  if (need_w) {
    uint8_t* dst = reinterpret_cast<uint8_t*>(need_h ? tmp : out);
    const long long rows = (long long)n * h_in;
    if (channels == 3)
      resize_h_kernel<3><<<grid_for(rows * w_out), 256, 0, stream>>>(src, dst, rows, w_in, w_out, bounds_w, coeffs_w, ksize_w);
    else
      resize_h_kernel<1><<<grid_for(rows * w_out), 256, 0, stream>>>(src, dst, rows, w_in, w_out, bounds_w, coeffs_w, ksize_w);
    vl::g_launches.fetch_add(1);
    VL_CHECK_CUDA(cudaGetLastError());
    src = dst;
  }
  if (need_h) {
    const int row_bytes = w_out * channels;
    resize_v_kernel<<<grid_for((long long)n * h_out * row_bytes), 256, 0, stream>>>(
        src, reinterpret_cast<uint8_t*>(out), n, h_in, h_out, row_bytes, bounds_h, coeffs_h, ksize_h);
    vl::g_launches.fetch_add(1);
    VL_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}
