// Development probe (not on the product path): per-SM fill rate of TMA tiled vs im2col loads as a function of box
// rows and ring depth.  One thread per CTA issues `iters` loads into a ring of `stages` slots and waits for slot
// reuse; the kernel reports elapsed SM cycles per CTA.
#include "common.cuh"
#include "../../include/vlb200.h"

namespace {
using namespace vl::ptx;

__global__ void __launch_bounds__(64, 1)
    tma_bench_kernel(const __grid_constant__ CUtensorMap tm, int im2col, int rows, int stages, int iters, int pq, int q,
                     int n_img, long long* out_cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint8_t* tiles = smem + 1024;
  const int bytes = rows * 128;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(&bars[s], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t0 = clock64();
    // coordinates advance with adds/compares only: the issuing thread must not be the bottleneck
    int n = (blockIdx.x * 7) % n_img, p = 0, qq = 0, tr = 0, ts = 0;
    int row = (blockIdx.x * 1031 * rows) % (n_img * pq - rows);
    int s = 0;
    uint32_t par = 0;
    for (int i = 0; i < iters; ++i) {
      if (i >= stages) mbar_wait(&bars[s], par ^ 1u);
      mbar_expect_tx(&bars[s], bytes);
      if (im2col) {
        tma_load_im2col_4d(tiles + s * bytes, &tm, &bars[s], 0, qq - 1, p - 1, n, (uint16_t)ts, (uint16_t)tr);
        if (++ts == 3) { ts = 0; if (++tr == 3) { tr = 0; if (++n >= n_img - 2) n = 0; } }
      } else {
        tma_load_2d(tiles + s * bytes, &tm, &bars[s], 0, row);
        row += rows;
        if (row >= n_img * pq - rows) row = 0;
      }
      if (++s == stages) { s = 0; par ^= 1u; }
    }
    (void)q;
    for (int k = 0; k < stages; ++k) {  // drain: the last `stages` loads
      mbar_wait(&bars[s], par ^ 1u);
      if (++s == stages) { s = 0; par ^= 1u; }
    }
    out_cycles[blockIdx.x] = clock64() - t0;
  }
}
}  // namespace

// exported for tests/bringup/tma_bench.py only
extern "C" int vl_debug_tma_bench(const void* base, int32_t im2col, int32_t n_img, int32_t h, int32_t w, int32_t c,
                                  int32_t rows, int32_t stages, int32_t iters, int32_t grid, long long* out_cycles,
                                  vl_stream_t stream_);

#include <cuda.h>
namespace vlb_dbg {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
}

extern "C" int vl_debug_tma_bench(const void* base, int32_t im2col, int32_t n_img, int32_t h, int32_t w, int32_t c,
                                  int32_t rows, int32_t stages, int32_t iters, int32_t grid, long long* out_cycles,
                                  vl_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  cudaDriverEntryPointQueryResult qr;
  void* fn = nullptr;
  CUtensorMap tm;
  if (im2col) {
    VL_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &qr));
    cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n_img};
    cuuint64_t strides[3] = {(cuuint64_t)c * 2, (cuuint64_t)c * w * 2, (cuuint64_t)c * w * h * 2};
    int lower[2] = {-1, -1}, upper[2] = {-1, -1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = reinterpret_cast<vlb_dbg::EncodeIm2colFn>(fn)(
        &tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, lower, upper, 64,
        (cuuint32_t)rows, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VL_REQUIRE(r == CUDA_SUCCESS, "im2col encode failed %d", (int)r);
  } else {
    VL_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
    cuuint64_t dims[2] = {(cuuint64_t)c, (cuuint64_t)n_img * h * w};
    cuuint64_t strides[1] = {(cuuint64_t)c * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = reinterpret_cast<vlb_dbg::EncodeTiledFn>(fn)(
        &tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VL_REQUIRE(r == CUDA_SUCCESS, "tiled encode failed %d", (int)r);
  }
  const int smem = 2048 + stages * rows * 128;
  VL_REQUIRE(smem <= 232448, "too much smem");
  VL_CHECK_CUDA(cudaFuncSetAttribute(tma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
  tma_bench_kernel<<<grid, 64, smem, stream>>>(tm, im2col, rows, stages, iters, h * w, w, n_img, out_cycles);
  VL_CHECK_CUDA(cudaGetLastError());
  return 0;
}
