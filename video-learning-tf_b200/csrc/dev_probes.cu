// Development probes (NOT on the product path, not declared in include/vlb200.h): micro-benchmarks that settled design
// questions on real hardware -- TMA fill rates, the cost of the mbarrier hand-shake and of tcgen05.mma vs tile shape,
// and whether swizzled operands may start at arbitrary row offsets.  Driven by tests/bringup/{tma_bench,sync_bench,shift_mma}.py.
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

// ------------------------------------------------------------------------------------------------
// Development probe 2: cost per iteration of the producer/consumer mbarrier handshake of the contraction kernel,
// without any data movement.  variant bits: 1 = consumer releases slots with tcgen05.commit (else mbarrier.arrive),
// 2 = tcgen05.fence::after_thread_sync after each full-wait, 4 = consumer issues 4 UMMAs (128 x bn x 16) per iteration.
// ------------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(64, 1)
    sync_bench_kernel(int variant, int stages, int iters, int bn, int bm, long long* out_cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty_bar = full_bar + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 256);
  const uint32_t tiles = smem_u32(smem + 1024);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t full_u32 = smem_u32(full_bar), empty_u32 = smem_u32(empty_bar);
  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < iters; ++i) {
      mbar_wait_u32(empty_u32 + stage * 8, phase ^ 1u);
      if (elect_one()) mbar_expect_tx_u32(full_u32 + stage * 8, 0);
      __syncwarp();
      if (++stage == stages) { stage = 0; phase ^= 1u; }
    }
  } else {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | (((uint32_t)bm >> 4) << 24);
    const uint64_t hi = (static_cast<uint64_t>((1024u >> 4) | (1u << 14) | (2u << 29)) << 32) | (1ull << 16);
    int stage = 0;
    uint32_t phase = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      mbar_wait_u32(full_u32 + stage * 8, phase);
      if (variant & 2) tc_fence_after();
      const uint64_t adesc = hi | (((tiles + stage * 49152) >> 4) & 0x3FFFu);
      const uint64_t bdesc = hi | (((tiles + stage * 49152 + 16384) >> 4) & 0x3FFFu);
      if (elect_one()) {
        if (variant & 4) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, adesc + k * 2, bdesc + k * 2, idesc, (uint32_t)(i | k));
        }
        if (variant & 1) umma_commit_u32(empty_u32 + stage * 8);
        else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty_u32 + stage * 8) : "memory");
      }
      __syncwarp();
      if (++stage == stages) { stage = 0; phase ^= 1u; }
    }
    // drain: wait until the last commit has landed (producer side of the last used slot)
    long long t1 = clock64();
    if (threadIdx.x == 32) out_cycles[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  // let outstanding MMAs / commits finish before TMEM is released
  if (warp == 1) {
    __nanosleep(20000);
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}
}  // namespace

extern "C" int vl_debug_sync_bench(int32_t variant, int32_t stages, int32_t iters, int32_t bn, int32_t bm, int32_t grid,
                                   long long* out_cycles, vl_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VL_REQUIRE(stages >= 1 && stages <= 4, "stages 1..4");
  const int smem = 2048 + stages * 49152;
  VL_CHECK_CUDA(cudaFuncSetAttribute(sync_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
  sync_bench_kernel<<<grid, 64, smem, stream>>>(variant, stages, iters, bn, bm, out_cycles);
  VL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Development probe 3: can a K-major SWIZZLE_128B operand be addressed at an arbitrary ROW offset (start address
// = tile + off * 128 B, descriptor "base offset" = (start >> 7) & 7)?  D[128][n] = W[128][64] * X[off + j][64]^T.
// mode bit 0: set the base-offset field; bit 1: shift the A operand instead of B (then D[i][j] = X[off+i] . W[j]).
// ------------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(128, 1)
    shift_mma_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX, int off, int n,
                     int mode, float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  uint64_t* done = bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 64);
  uint8_t* sW = smem + 1024;            // 128 rows x 128 B
  uint8_t* sX = smem + 1024 + 16384;    // 512 rows x 128 B
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, 16384 + 65536);
    tma_load_2d(sW, &tmW, bar, 0, 0);
    tma_load_2d(sX, &tmX, bar, 0, 0);
    tma_load_2d(sX + 32768, &tmX, bar, 0, 256);
    mbar_wait(bar, 0);
    tc_fence_after();
    const bool shift_a = mode & 2;
    const uint32_t a_addr = shift_a ? smem_u32(sX) + off * 128 : smem_u32(sW);
    const uint32_t b_addr = shift_a ? smem_u32(sW) : smem_u32(sX) + off * 128;
    auto desc = [&](uint32_t addr) {
      uint64_t d = (static_cast<uint64_t>((1024u >> 4) | (1u << 14) | (2u << 29)) << 32) | (1ull << 16) |
                   ((addr >> 4) & 0x3FFFu);
      if (mode & 1) d |= static_cast<uint64_t>((addr >> 7) & 7u) << 49;
      return d;
    };
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
    if (mode & 4) {
      // mode bit 2: N-major B whose 64-wide N atoms OVERLAP: atom s starts one pixel row (128 B) after atom s-1
      // (leading byte offset 128 B instead of a separate 8 KB block per atom), i.e. the three taps of a filter row
      // addressed through ONE descriptor over the same staged rows:  D[m][(s, c)] = sum_k W[m][k] * X[off + k + s][c],
      // k < 64 pixel rows, n = 192.  K advances by 16 rows = 2048 B per UMMA.
      const uint32_t idesc_mn = idesc | (1u << 16);
      auto desc_mn = [&](uint32_t addr) {
        return (static_cast<uint64_t>((1024u >> 4) | (1u << 14) | (2u << 29)) << 32) |
               (static_cast<uint64_t>(128u >> 4) << 16) | ((addr >> 4) & 0x3FFFu);
      };
      const uint32_t xb = smem_u32(sX) + off * 128;
      for (int k = 0; k < 4; ++k)
        umma_bf16(tmem_base, desc(smem_u32(sW) + k * 32), desc_mn(xb + k * 2048), idesc_mn, (uint32_t)k);
    } else
    // K-steps advance by 32 B inside the swizzle row; the base offset refers to the row phase of the start address
    for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, desc(a_addr + k * 32), desc(b_addr + k * 32), idesc, (uint32_t)k);
    umma_commit(done);
    mbar_wait(done, 0);
    tc_fence_after();
  }
  __syncthreads();
  tc_fence_after();
  // 4 warps read their 32-lane quadrant
  const uint32_t taddr = tmem_base + (uint32_t(warp * 32) << 16);
  for (int c0 = 0; c0 < n; c0 += 16) {
    uint32_t v[16];
    tmem_ld_x16(taddr + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) out[(warp * 32 + (threadIdx.x & 31)) * n + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}
}  // namespace

extern "C" int vl_debug_shift_mma(const void* w, const void* x, int32_t x_rows, int32_t off, int32_t n, int32_t mode,
                                  float* out, vl_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  cudaDriverEntryPointQueryResult qr;
  void* fn = nullptr;
  VL_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
  auto enc = reinterpret_cast<vlb_dbg::EncodeTiledFn>(fn);
  CUtensorMap tmW, tmX;
  cuuint32_t estr[2] = {1, 1};
  {
    cuuint64_t dims[2] = {64, 128};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, 128};
    VL_REQUIRE(enc(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS, "encode W failed");
  }
  {
    cuuint64_t dims[2] = {64, (cuuint64_t)x_rows};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, 256};
    VL_REQUIRE(enc(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(x), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS, "encode X failed");
  }
  VL_CHECK_CUDA(cudaFuncSetAttribute(shift_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  shift_mma_kernel<<<1, 128, 1024 + 1024 + 16384 + 65536, stream>>>(tmW, tmX, off, n, mode, out);
  VL_CHECK_CUDA(cudaGetLastError());
  return 0;
}
