// Tensor-core contraction core for the LRCN hot path (sm_100a only).
//
// One persistent, warp-specialised kernel: TMA (tiled or im2col mode) stages bf16 operand tiles in
// 128B-swizzled shared memory, one elected lane issues tcgen05.mma (UMMA 128 x BN x 16) into
// double-buffered TMEM accumulators, four epilogue warps drain TMEM with tcgen05.ld and apply
// bias / ReLU / ReLU-gradient mask before vectorised stores (or split-K red.add for filter gradients).
// A CTA tile is 128*msub rows x BN columns: with msub = 2 (BN <= 128) two 128-row sub-tiles share every
// B load and every pipeline hand-shake, which is what the narrow-N convolutions (conv1, conv2, conv5) need.
//
// It replaces the TensorFlow ops of the reference's hot path:
//   tf.nn.conv2d (+ split/concat groups)   models/alexnet/alexnet.py:15-31
//   tf.nn.relu_layer / xw_plus_b           models/alexnet/alexnet.py:228,248,275 ; tf_util.py:56
//   BasicLSTMCell input projection         models/lstm/lstm.py:9-20,141
//   and the conv2d / matmul gradients TF derives for train.py:210.
#include "common.cuh"
#include "../../include/vlb200.h"

#include <atomic>
#include <cstdlib>

namespace vl {
extern std::atomic<long long> g_launches;
}

namespace {

using namespace vl::ptx;
typedef __nv_bfloat16 bf16;

constexpr int BM = 128;                     // UMMA M (cta_group::1)
constexpr int BK = 64;                      // bf16 elements per k-block = one 128B swizzle row
constexpr int A_SUB_BYTES = BM * BK * 2;    // 16 KB per 128-row sub-tile and k-block
constexpr int EPI_GROUPS = 2;               // epilogue groups of four warps (one warp per TMEM lane quadrant and group)
constexpr int PROD2_WARP = 2 + EPI_GROUPS * 4;   // second TMA producer (odd k-blocks), after the epilogue warps
constexpr int NUM_THREADS = 64 + EPI_GROUPS * 128 + 32;  // warp0 TMA, warp1 MMA (+TMEM alloc), epilogue warps, warp PROD2 TMA
constexpr int TMEM_COLS = 512;              // 2 accumulator stages x 256 fp32 columns
constexpr int ACC_STRIDE_COLS = 256;
constexpr int MAX_STAGES = 8;
constexpr int MAX_MSUB = 2;
constexpr int SMEM_LIMIT = 232448;          // 227 KB opt-in maximum per CTA
constexpr int BAR_REGION = 2048;            // mbarriers + TMEM slot (256 B) + bias slice (1 KB)

// Division by a runtime-invariant divisor without the ~40-instruction software divide: q = (mulhi(n, m) + n) >> s
// (round-up method; exact for 0 <= n < 2^31).
struct FastDiv {
  uint32_t m, s, d;
};
__device__ __forceinline__ int fd_div(int n, const FastDiv& f) {
  return (int)((__umulhi((uint32_t)n, f.m) + (uint32_t)n) >> f.s);
}
__device__ __forceinline__ void fd_divmod(int n, const FastDiv& f, int& q, int& r) {
  q = fd_div(n, f);
  r = n - q * (int)f.d;
}

struct KParams {
  int M, N;  // valid extents per group
  int groups, num_m_blk, num_n_blk, split_k, kb_total, kb_per_split, total_tiles;
  int BN, msub;
  int a_goff, b_goff, c_goff;
  int b_tap_stride, b_row_goff, b_tap_inner;
  int taps, cchunks, kw, flip;
  int P, Q, PQ, stride_h, stride_w, lower_h, lower_w, cin_g;
  FastDiv fd_nblk, fd_mblk, fd_groups, fd_cchunks, fd_kw, fd_PQ, fd_Q;
  int d2s_sh, d2s_sw, d2s_c, d2s_h, d2s_w;  // depth-to-space epilogue (d2s_c > 0), see vl_gemm_desc
  int row_shift;  // transposed-im2col B through ONE tiled box per filter row (see vl_gemm_desc.row_shift)
  FastDiv fd_P;
  FastDiv fd_d2s_c, fd_d2s_sw;
  void* C;
  int c_ld, c_dtype, c_atomic, relu;
  const float* bias;
  const bf16* mask;
  int mask_ld;
  int num_stages, b_stage_bytes, stage_bytes;
  uint32_t idesc;
  uint32_t a_desc_hi, b_desc_hi;      // upper 32 bits of the smem descriptors (SBO, version, swizzle)
  uint32_t a_lbo_enc, b_lbo_enc;      // encoded leading byte offset (bits 16..29 of the low word)
  uint32_t a_kstep_enc, b_kstep_enc;  // encoded start-address advance per UMMA_K (=16 elements)
  int mma_cchunks;  // k-blocks per tap as seen by the MMA issuer (tail detection)
  int ksteps_tail;  // UMMA K-steps (of 16) that carry data in the last channel chunk of a tap / last dense k-block
  int mn_step_rows, mn_step_rem;  // transposed im2col: 64 pixels = mn_step_rows image rows + mn_step_rem pixels
  int dbg;  // development probes (VL_GEMM_DBG): 1 = no MMA, 2 = no A loads, 4 = no B loads, 8 = no stores
  int producers;  // 1: warp 0 stages every k-block; 2: warp 0 the even, warp PROD2_WARP the odd k-blocks (see the producer)
};

__device__ __forceinline__ void tma_load_tiled_4d_u32(uint32_t dst, const CUtensorMap* m, uint32_t bar, int32_t c0,
                                                      int32_t c1, int32_t c2, int32_t c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      :
      : "r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

struct TileCoord {
  int m_blk, n_blk, g, kb_begin, kb_end;
};

__device__ __forceinline__ TileCoord decode_tile(const KParams& p, int tile) {
  TileCoord t;
  int r, split;
  fd_divmod(tile, p.fd_nblk, r, t.n_blk);
  fd_divmod(r, p.fd_mblk, r, t.m_blk);
  fd_divmod(r, p.fd_groups, split, t.g);
  t.kb_begin = split * p.kb_per_split;
  t.kb_end = min(p.kb_total, t.kb_begin + p.kb_per_split);
  return t;
}

// The producer and the MMA issuer are WARP-UNIFORM loops: all 32 lanes run the loop control on values derived
// from blockIdx / kernel parameters and one elected lane issues the TMA / tcgen05 instructions.  A single-lane
// (`lane == 0`) formulation was measured at ~450 cycles of issue overhead per k-block (R2UR round trips and
// per-instruction ELECT loops), which capped every shape below 45 % of the tensor pipe
// (profiles/r01_gemm_probe.txt).
template <int AM, int BMODE>
__global__ void __launch_bounds__(NUM_THREADS, 1)
    umma_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ KParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  // 128B swizzle atoms need 1024B-aligned tiles.  The offset is applied to the __shared__ array itself (not through an
  // integer round trip) so that the compiler keeps the shared address space: LDS/STS instead of generic LD/ST.
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* tiles = smem + BAR_REGION;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* tmem_full = empty_bar + MAX_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* sbias = reinterpret_cast<float*>(smem + 256);  // [256] bias slice of the current tile

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform role index
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < p.num_stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 4 * EPI_GROUPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_base_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  const uint32_t tiles_u32 = smem_u32(tiles);
  const uint32_t full_u32 = smem_u32(full_bar);
  const uint32_t empty_u32 = smem_u32(empty_bar);
  const int num_stages = p.num_stages;
  const int stage_bytes = p.stage_bytes;
  const int total_tiles = p.total_tiles;
  const int msub = p.msub;
  const int BN = p.BN;
  const uint32_t a_tile_bytes = (uint32_t)msub * A_SUB_BYTES;

  if (warp == 0 || (warp == PROD2_WARP && p.producers == 2)) {
    // ===================== TMA producer(s) =====================
    // A producer iteration costs ~90 dependent single-warp instructions (barrier wait, R2UR moves of the box
    // coordinates, one UTMALDG per box, walk advance): ~390 cycles, which is as long as the four UMMAs of a short k-block
    // (filter gradients: 128 x 192..256 x 64 per k-block; profiles/r02_bwd_stage_decomposition.txt).  With two producer
    // warps every k-block is staged by exactly one of them (running k-block counter parity); both walk the full
    // contraction state, only the owner waits for the slot and issues the loads.
    const int prod_id = warp == 0 ? 0 : 1;
    const bool two_prod = p.producers == 2;
    int kcount = 0;  // running k-block counter over all tiles of this CTA
    // The k-block loop stays free of integer divisions: the (tap, channel-chunk) and pixel coordinates advance
    // with adds and compares only; the per-tile set-up uses multiply-shift division.
    const int cchunks = p.cchunks, kw = p.kw, taps = p.taps;
    const bool walk_taps = (AM == VL_A_IM2COL_K) || taps > 1;
    const bool ld_a = !(p.dbg & 2), ld_b = !(p.dbg & 4);
    const int mn_step_rows = p.mn_step_rows, mn_step_rem = p.mn_step_rem;
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(p, tile);
      const int m0 = t.m_blk * BM * msub;
      const int n0 = t.n_blk * BN;
      const int a_c0 = t.g * p.a_goff;
      const int b_c0 = t.g * p.b_goff;
      const int b_r0 = t.g * p.b_row_goff + n0;
      // contraction walk state (per tile, then incremental)
      int tap = 0, cc = 0, tr = 0, ts = 0;
      if (walk_taps) {
        fd_divmod(t.kb_begin, p.fd_cchunks, tap, cc);
        fd_divmod(tap, p.fd_kw, tr, ts);
      } else {
        cc = t.kb_begin;
      }
      // im2col K-major: base pixel of the first row of each 128-row sub-tile
      int sub_wx[MAX_MSUB] = {0, 0}, sub_wy[MAX_MSUB] = {0, 0}, sub_n[MAX_MSUB] = {0, 0};
      // transposed im2col: current pixel of the k-block, and the (channel, tap) of up to 2*msub 64-row chunks
      int pn = 0, pp = 0, pq = 0;
      int mn_c[2 * MAX_MSUB] = {0, 0, 0, 0}, mn_tr[2 * MAX_MSUB] = {0, 0, 0, 0}, mn_ts[2 * MAX_MSUB] = {0, 0, 0, 0};
      int a_valid = 2 * msub;
      if (AM == VL_A_IM2COL_K) {
#pragma unroll
        for (int s = 0; s < MAX_MSUB; ++s) {
          if (s < msub) {
            int rem, ppp, pqq;
            fd_divmod(m0 + s * BM, p.fd_PQ, sub_n[s], rem);
            fd_divmod(rem, p.fd_Q, ppp, pqq);
            sub_wx[s] = pqq * p.stride_w + p.lower_w;
            sub_wy[s] = ppp * p.stride_h + p.lower_h;
          }
        }
      } else if (AM == VL_A_IM2COL_MN) {
        int rem;
        fd_divmod(t.kb_begin * BK, p.fd_PQ, pn, rem);
        fd_divmod(rem, p.fd_Q, pp, pq);
        const int chunks = taps * cchunks;
        a_valid = min(2 * msub, chunks - t.m_blk * 2 * msub);
#pragma unroll
        for (int j = 0; j < 2 * MAX_MSUB; ++j) {
          if (j < a_valid) {
            int tp, ch;
            fd_divmod(t.m_blk * 2 * msub + j, p.fd_cchunks, tp, ch);
            mn_c[j] = a_c0 + ch * BK;
            fd_divmod(tp, p.fd_kw, mn_tr[j], mn_ts[j]);
          }
        }
      }
      int b_valid = 0;  // transposed-im2col B: 64-column chunks (tap, channel chunk) of this n-block
      if (BMODE == VL_B_IM2COL_MN && p.row_shift) {
        // one k-block = one OUTPUT ROW (q <= 64 pixels, zero padded to 64 by the out-of-bounds fill of the A box);
        // the n-block is one filter row: its kw taps read the same staged input row shifted by 0..kw-1 pixels
        fd_divmod(t.kb_begin, p.fd_P, pn, pp);
        pq = 0;
        b_valid = BN >> 6;
      } else if (BMODE == VL_B_IM2COL_MN) {
        int rem;
        fd_divmod(t.kb_begin * BK, p.fd_PQ, pn, rem);
        fd_divmod(rem, p.fd_Q, pp, pq);
        const int per_blk = BN >> 6;
        b_valid = min(per_blk, taps * cchunks - t.n_blk * per_blk);
#pragma unroll
        for (int j = 0; j < 2 * MAX_MSUB; ++j) {
          if (j < b_valid) {
            int tp, ch;
            fd_divmod(t.n_blk * per_blk + j, p.fd_cchunks, tp, ch);
            mn_c[j] = b_c0 + ch * BK;
            fd_divmod(tp, p.fd_kw, mn_tr[j], mn_ts[j]);
          }
        }
      }
      const uint32_t a_bytes = !ld_a ? 0u : (AM == VL_A_IM2COL_MN ? 8192u * a_valid : a_tile_bytes);
      const uint32_t b_bytes = !ld_b ? 0u
                               : (BMODE == VL_B_IM2COL_MN ? (p.row_shift ? (uint32_t)(64 + p.kw - 1) * 128u : 8192u * b_valid)
                                                          : (uint32_t)p.b_stage_bytes);
      const uint32_t bytes = b_bytes + a_bytes;
      for (int kb = t.kb_begin; kb < t.kb_end; ++kb, ++kcount) {
        const bool mine = !two_prod || (kcount & 1) == prod_id;
        if (mine) mbar_wait_u32(empty_u32 + stage * 8, phase ^ 1u);
        const uint32_t sA = tiles_u32 + stage * stage_bytes;
        const uint32_t sB = sA + a_tile_bytes;
        const uint32_t fb = full_u32 + stage * 8;
        if (mine && elect_one()) {
          mbar_expect_tx_u32(fb, bytes);
          // ---- A ----
          if (ld_a) {
            if (AM == VL_A_TILED_K) {
#pragma unroll
              for (int s = 0; s < MAX_MSUB; ++s)
                if (s < msub) tma_load_2d_u32(sA + s * A_SUB_BYTES, &tmA, fb, a_c0 + kb * BK, m0 + s * BM);
            } else if (AM == VL_A_TILED_MN && BMODE == VL_B_IM2COL_MN && p.row_shift) {
              // dy^T of output row (pn, pp): 4-D box (64 channels, 64 pixels, 1 row); pixels >= q are zero filled
#pragma unroll
              for (int j = 0; j < 2 * MAX_MSUB; ++j)
                if (j < 2 * msub) tma_load_tiled_4d_u32(sA + j * 8192, &tmA, fb, a_c0 + m0 + j * 64, 0, pp, pn);
            } else if (AM == VL_A_TILED_MN) {
#pragma unroll
              for (int j = 0; j < 2 * MAX_MSUB; ++j)
                if (j < 2 * msub) tma_load_2d_u32(sA + j * 8192, &tmA, fb, a_c0 + m0 + j * 64, kb * BK);
            } else if (AM == VL_A_IM2COL_K) {
#pragma unroll
              for (int s = 0; s < MAX_MSUB; ++s)
                if (s < msub)
                  tma_load_im2col_4d_u32(sA + s * A_SUB_BYTES, &tmA, fb, a_c0 + cc * BK, sub_wx[s], sub_wy[s], sub_n[s],
                                         (uint16_t)ts, (uint16_t)tr);
            } else {  // VL_A_IM2COL_MN: this k-block's 64 pixels start at (pn, pp, pq)
              const int bx = pq * p.stride_w + p.lower_w, by = pp * p.stride_h + p.lower_h;
#pragma unroll
              for (int j = 0; j < 2 * MAX_MSUB; ++j)
                if (j < a_valid)
                  tma_load_im2col_4d_u32(sA + j * 8192, &tmA, fb, mn_c[j], bx, by, pn, (uint16_t)mn_ts[j],
                                         (uint16_t)mn_tr[j]);
            }
          }
          // ---- B ----
          if (ld_b) {
            if (BMODE == VL_B_TILED_K) {
              const int tapb = p.flip ? (taps - 1 - tap) : tap;
              tma_load_2d_u32(sB, &tmB, fb, b_c0 + tap * p.b_tap_inner + cc * BK, b_r0 + tapb * p.b_tap_stride);
            } else if (BMODE == VL_B_IM2COL_MN && p.row_shift) {
              // input row (pp * stride + lower_h + filter row) of image pn, 64 + kw - 1 pixels from lower_w: ONE tiled box
              tma_load_tiled_4d_u32(sB, &tmB, fb, b_c0, p.lower_w, pp * p.stride_h + p.lower_h + t.n_blk, pn);
            } else if (BMODE == VL_B_IM2COL_MN) {  // this k-block's 64 pixels start at (pn, pp, pq)
              const int bx = pq * p.stride_w + p.lower_w, by = pp * p.stride_h + p.lower_h;
#pragma unroll
              for (int j = 0; j < 2 * MAX_MSUB; ++j)
                if (j < b_valid)
                  tma_load_im2col_4d_u32(sB + j * 8192, &tmB, fb, mn_c[j], bx, by, pn, (uint16_t)mn_ts[j],
                                         (uint16_t)mn_tr[j]);
            } else {
              for (int j = 0; j < BN; j += 64) tma_load_2d_u32(sB + j * 128, &tmB, fb, b_c0 + n0 + j, kb * BK);
            }
          }
        }
        __syncwarp();
        // ---- advance the contraction walk ----
        if (BMODE == VL_B_IM2COL_MN && p.row_shift) {
          if (++pp == p.P) {
            pp = 0;
            ++pn;
          }
        } else if (AM == VL_A_IM2COL_MN || BMODE == VL_B_IM2COL_MN) {
          pq += mn_step_rem;  // advance 64 pixels = mn_step_rows rows + mn_step_rem pixels
          pp += mn_step_rows;
          if (pq >= p.Q) {
            pq -= p.Q;
            ++pp;
          }
          while (pp >= p.P) {
            pp -= p.P;
            ++pn;
          }
        }
        if (++cc == cchunks && walk_taps) {
          cc = 0;
          ++tap;
          if (++ts == kw) {
            ts = 0;
            ++tr;
          }
        }
        if (++stage == num_stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one elected lane; warp-uniform control) =====================
    const uint32_t idesc_full = p.idesc;
    const uint64_t a_hi = (static_cast<uint64_t>(p.a_desc_hi) << 32) | (static_cast<uint64_t>(p.a_lbo_enc) << 16);
    const uint64_t b_hi = (static_cast<uint64_t>(p.b_desc_hi) << 32) | (static_cast<uint64_t>(p.b_lbo_enc) << 16);
    const uint32_t a_kstep = p.a_kstep_enc, b_kstep = p.b_kstep_enc;
    const bool do_mma = !(p.dbg & 1);
    const int cchunks_mma = p.mma_cchunks, ksteps_tail = p.ksteps_tail;
    int stage = 0;
    uint32_t phase = 0;
    int local = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
      const TileCoord t = decode_tile(p, tile);
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      mbar_wait_u32(smem_u32(&tmem_empty[acc]), acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * ACC_STRIDE_COLS;
      uint32_t accumulate = 0;
      // transposed-im2col B: the last n-block may hold fewer (tap, channel chunk) columns than BN (conv2: 25 taps =
      // 6 x 4 + 1); its UMMAs are issued with N = 64 x the valid chunks instead of multiplying unloaded shared memory
      uint32_t idesc = idesc_full;
      if (BMODE == VL_B_IM2COL_MN && !p.row_shift) {
        const int per_blk = BN >> 6;
        const int valid = min(per_blk, p.taps * p.cchunks - t.n_blk * per_blk);
        idesc = (idesc_full & ~(0x3Fu << 17)) | ((uint32_t)(valid * 64 >> 3) << 17);
      }
      int cc = t.kb_begin % cchunks_mma;
      for (int kb = t.kb_begin; kb < t.kb_end; ++kb) {
        const bool tail = (++cc == cchunks_mma);
        if (tail) cc = 0;
        mbar_wait_u32(full_u32 + stage * 8, phase);
        tc_fence_after();
        const uint32_t sA = tiles_u32 + stage * stage_bytes;
        const uint64_t adesc = a_hi | ((sA >> 4) & 0x3FFFu);
        const uint64_t bdesc = b_hi | (((sA + a_tile_bytes) >> 4) & 0x3FFFu);
        if (elect_one()) {
          if (do_mma) {
            if (!tail || ksteps_tail == BK / 16) {
#pragma unroll
              for (int s = 0; s < MAX_MSUB; ++s) {
                if (s < msub) {
#pragma unroll
                  for (int k = 0; k < BK / 16; ++k)
                    umma_bf16(tmem_d + s * BN, adesc + s * (A_SUB_BYTES >> 4) + k * a_kstep, bdesc + k * b_kstep, idesc,
                              accumulate | (uint32_t)k);
                }
              }
            } else {
              // zero-padded K-steps of the last channel chunk (cin_g = 48 -> 3 of 4) are not issued at all
              for (int s = 0; s < msub; ++s)
                for (int k = 0; k < ksteps_tail; ++k)
                  umma_bf16(tmem_d + s * BN, adesc + s * (A_SUB_BYTES >> 4) + k * a_kstep, bdesc + k * b_kstep, idesc,
                            accumulate | (uint32_t)k);
            }
          }
          umma_commit_u32(empty_u32 + stage * 8);  // frees the smem slot once the MMAs have read it
        }
        __syncwarp();
        accumulate = 1;
        if (++stage == num_stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
      if (elect_one()) umma_commit_u32(smem_u32(&tmem_full[acc]));  // accumulator complete -> epilogue
      __syncwarp();
    }
  } else if (warp < PROD2_WARP) {
    // ===================== epilogue warps (TMEM -> registers -> HBM) =====================
    // Two groups of four warps share a tile: warp % 4 selects the TMEM lane quadrant, (warp - 2) / 4 the group; group g
    // drains the 16-column chunks g, g + 2, g + 4, ...  A single warp sustains ~650 cycles per chunk, so the short-K
    // products (fc filter gradients: 16 k-blocks per 128 x 256 fp32 tile; conv5; the LSTM / head matrices) were bound by
    // four warps draining 16 chunks per sub-tile while the tensor pipe waited for the accumulator stage.
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const int grp = (warp - 2) >> 2;
    const int epi_tid = threadIdx.x - 64;
    int local = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
      const TileCoord t = decode_tile(p, tile);
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      const int m0 = t.m_blk * BM * msub;
      const int n0 = t.n_blk * BN;
      const int row_in_tile = quad * 32 + lane;
      const int gcol0 = t.g * p.c_goff + n0;  // global column of tile column 0
      // stage this tile's bias slice in shared memory once (instead of one global load per element)
      if (p.bias != nullptr) {
        asm volatile("bar.sync 1, %0;" ::"n"(EPI_GROUPS * 128) : "memory");  // previous tile's readers are done
        // split-K (atomic fp32 output): only the split that owns the first k-block adds the bias
        for (int j = epi_tid; j < BN; j += EPI_GROUPS * 128)
          sbias[j] = (n0 + j < p.N && t.kb_begin == 0) ? __ldg(p.bias + gcol0 + j) : 0.f;
        asm volatile("bar.sync 1, %0;" ::"n"(EPI_GROUPS * 128) : "memory");
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const bool vec_ok = (p.c_ld % 8 == 0) && (gcol0 % 8 == 0) && !p.c_atomic &&
                          ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) &&
                          (p.mask == nullptr || p.mask_ld % 8 == 0);
      const bool vec4_ok = (p.c_ld % 4 == 0) && (gcol0 % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);  // 16-byte aligned fp32 quads (vector red.add)
      for (int sub = 0; sub < msub; ++sub) {
        long long grow;
        bool row_ok;
        if (AM == VL_A_IM2COL_MN) {
          const int mc = (t.m_blk * msub + sub) * 2 + (row_in_tile >> 6);
          int tap, cc;
          fd_divmod(mc, p.fd_cchunks, tap, cc);
          const int ci = cc * 64 + (row_in_tile & 63);
          row_ok = (mc < p.taps * p.cchunks) && (ci < p.cin_g);
          grow = (long long)tap * p.cin_g + ci;
        } else {
          grow = m0 + sub * BM + row_in_tile;
          row_ok = grow < p.M;
        }
        const uint32_t taddr = tmem_base + (uint32_t(quad * 32) << 16) + acc * ACC_STRIDE_COLS + sub * BN;
        // depth-to-space epilogue: row (n, Y, X) -> base pixel (sh*Y, sw*X) of the NHWC output
        long long d2s_base = 0;
        int d2s_y = 0, d2s_x = 0;
        if (p.d2s_c > 0 && row_ok) {
          int nn, rem, yy, xx;
          fd_divmod((int)grow, p.fd_PQ, nn, rem);
          fd_divmod(rem, p.fd_Q, yy, xx);
          d2s_y = yy * p.d2s_sh;
          d2s_x = xx * p.d2s_sw;
          d2s_base = (((long long)nn * p.d2s_h + d2s_y) * p.d2s_w + d2s_x) * p.c_ld + t.g * p.c_goff;
        }

        // ReLU-gradient mask of a 16-column chunk, fetched two chunks ahead of its use (the row-strided 32-byte
        // reads of the epilogue threads are latency bound when issued at the point of use)
        const bool mask_vec = p.mask != nullptr && row_ok && vec_ok && !(p.dbg & 8);
        auto fetch_mask = [&](int c0, uint4 (&m)[2]) {
          if (mask_vec && c0 < BN && n0 + c0 + 16 <= p.N) {
            const bf16* mrow = p.mask + grow * p.mask_ld + gcol0 + c0;
            if ((reinterpret_cast<uintptr_t>(mrow) & 31) == 0) {  // one 32-byte (full sector) read per thread
              asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                           : "=r"(m[0].x), "=r"(m[0].y), "=r"(m[0].z), "=r"(m[0].w), "=r"(m[1].x), "=r"(m[1].y), "=r"(m[1].z),
                             "=r"(m[1].w)
                           : "l"(mrow));
            } else {
              m[0] = __ldg(reinterpret_cast<const uint4*>(mrow));
              m[1] = __ldg(reinterpret_cast<const uint4*>(mrow + 8));
            }
          }
        };
        auto process = [&](const uint32_t (&v)[16], int c0, const uint4 (&mpre)[2]) {
          if (!row_ok || (p.dbg & 8)) return;
          if (BMODE == VL_B_IM2COL_MN) {
            // filter gradient with the filter taps on the N side: C[(tap, ci)][g * c_goff + row]; consecutive lanes
            // (rows = output channels) hit consecutive addresses, so the red.adds of a warp coalesce
            int tap, cc;
            fd_divmod((n0 + c0) >> 6, p.fd_cchunks, tap, cc);
            const int ci0 = cc * 64 + (c0 & 63);
            const int ncols = (tap < p.taps) ? min(16, p.cin_g - ci0) : 0;
            float* outp = reinterpret_cast<float*>(p.C) + ((long long)tap * p.cin_g + ci0) * p.c_ld + t.g * p.c_goff + grow;
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (j < ncols) {
                if (p.c_atomic)
                  atomicAdd(outp + (long long)j * p.c_ld, __uint_as_float(v[j]));
                else
                  outp[(long long)j * p.c_ld] = __uint_as_float(v[j]);
              }
            return;
          }
          const int ncols = min(16, p.N - (n0 + c0));
          if (ncols <= 0) return;
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
          if (p.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 bv = *reinterpret_cast<const float4*>(sbias + c0 + 4 * j);
              f[4 * j] += bv.x;
              f[4 * j + 1] += bv.y;
              f[4 * j + 2] += bv.z;
              f[4 * j + 3] += bv.w;
            }
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.0f);
          }
          // bf16 vector path: the mask is applied to the packed pairs below (one HSET2 + one LOP3 per pair)
          const bool mask_packed = p.mask != nullptr && vec_ok && ncols == 16 && p.c_dtype == VL_DT_BF16;
          if (p.mask != nullptr && !mask_packed) {
            const bf16* mrow = p.mask + grow * p.mask_ld + gcol0 + c0;
            if (vec_ok && ncols == 16) {
              uint4 m0v = mpre[0];
              uint4 m1v = mpre[1];
              const bf16* mv0 = reinterpret_cast<const bf16*>(&m0v);
              const bf16* mv1 = reinterpret_cast<const bf16*>(&m1v);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                if (!(__bfloat162float(mv0[j]) > 0.0f)) f[j] = 0.0f;
                if (!(__bfloat162float(mv1[j]) > 0.0f)) f[8 + j] = 0.0f;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (j < ncols && !(__bfloat162float(mrow[j]) > 0.0f)) f[j] = 0.0f;
            }
          }
          long long off = grow * p.c_ld + gcol0 + c0;
          if (p.d2s_c > 0) {
            int seg, within, ddy, ddx;
            fd_divmod(n0 + c0, p.fd_d2s_c, seg, within);
            fd_divmod(seg, p.fd_d2s_sw, ddy, ddx);
            if (d2s_y + ddy >= p.d2s_h || d2s_x + ddx >= p.d2s_w) return;
            off = d2s_base + (long long)(ddy * p.d2s_w + ddx) * p.c_ld + within;
          }
          if (p.c_dtype == VL_DT_BF16) {
            bf16* out = reinterpret_cast<bf16*>(p.C) + off;
            if (vec_ok && ncols == 16) {
              uint32_t w[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
                w[j] = *reinterpret_cast<uint32_t*>(&h);
              }
              if (mask_packed) {
                const uint32_t mw[8] = {mpre[0].x, mpre[0].y, mpre[0].z, mpre[0].w,
                                        mpre[1].x, mpre[1].y, mpre[1].z, mpre[1].w};
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  uint32_t keep;  // 0xffff per half where mask > 0 (tf ReluGrad)
                  asm("set.gt.u32.bf16x2 %0, %1, %2;" : "=r"(keep) : "r"(mw[j]), "r"(0u));
                  w[j] &= keep;
                }
              }
              if ((reinterpret_cast<uintptr_t>(out) & 31) == 0) {  // one full 32-byte sector per thread
                asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(out), "r"(w[0]), "r"(w[1]),
                             "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                             : "memory");
              } else {
                *reinterpret_cast<uint4*>(out) = make_uint4(w[0], w[1], w[2], w[3]);
                *reinterpret_cast<uint4*>(out + 8) = make_uint4(w[4], w[5], w[6], w[7]);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (j < ncols) out[j] = __float2bfloat16_rn(f[j]);
            }
          } else {
            float* out = reinterpret_cast<float*>(p.C) + off;
            if (p.c_atomic) {
              if (vec4_ok && ncols == 16) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(out + 4 * j), "f"(f[4 * j]),
                               "f"(f[4 * j + 1]), "f"(f[4 * j + 2]), "f"(f[4 * j + 3])
                               : "memory");
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  if (j < ncols) atomicAdd(out + j, f[j]);
              }
            } else if (vec_ok && ncols == 16) {
              if ((reinterpret_cast<uintptr_t>(out) & 31) == 0) {
                // 32-byte stores: a thread's 64 bytes leave as two FULL sectors instead of four half-sector writes (the
                // rows of a warp are c_ld floats apart, so nothing else fills the other half of a sector)
#pragma unroll
                for (int j = 0; j < 2; ++j)
                  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(out + 8 * j), "f"(f[8 * j]),
                               "f"(f[8 * j + 1]), "f"(f[8 * j + 2]), "f"(f[8 * j + 3]), "f"(f[8 * j + 4]), "f"(f[8 * j + 5]),
                               "f"(f[8 * j + 6]), "f"(f[8 * j + 7])
                               : "memory");
              } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  *reinterpret_cast<float4*>(out + 4 * j) = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (j < ncols) out[j] = f[j];
            }
          }
        };

        // software pipelined TMEM drain over this group's chunks (stride 16 * EPI_GROUPS columns): the load of the
        // group's next chunk is in flight while the current one is processed; ReLU-gradient masks one chunk ahead
        constexpr int CSTEP = 16 * EPI_GROUPS;
        uint32_t va[16], vb[16];
        uint4 ma[2], mb[2];
        ma[0] = ma[1] = mb[0] = mb[1] = make_uint4(0, 0, 0, 0);
        const int cfirst = grp * 16;
        if (cfirst < BN) {
          fetch_mask(cfirst, ma);
          tmem_ld_x16(taddr + cfirst, va);
        }
        for (int c0 = cfirst; c0 < BN; c0 += 2 * CSTEP) {
          const bool has_b = c0 + CSTEP < BN;
          if (has_b) fetch_mask(c0 + CSTEP, mb);
          tmem_ld_wait();
          if (has_b) tmem_ld_x16(taddr + c0 + CSTEP, vb);
          process(va, c0, ma);
          if (has_b) {
            const bool has_a = c0 + 2 * CSTEP < BN;
            if (has_a) fetch_mask(c0 + 2 * CSTEP, ma);
            tmem_ld_wait();
            if (has_a) tmem_ld_x16(taddr + c0 + 2 * CSTEP, va);
            process(vb, c0 + CSTEP, mb);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// Host side: tensor-map encoding through the driver entry points (no link-time libcuda dependency).
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn g_encode_tiled = nullptr;
EncodeIm2colFn g_encode_im2col = nullptr;

int load_driver_fns() {
  if (g_encode_tiled && g_encode_im2col) return 0;
  cudaDriverEntryPointQueryResult q;
  void* fn = nullptr;
  VL_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  VL_REQUIRE(fn != nullptr && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available");
  g_encode_tiled = reinterpret_cast<EncodeTiledFn>(fn);
  fn = nullptr;
  VL_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &q));
  VL_REQUIRE(fn != nullptr && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeIm2col not available");
  g_encode_im2col = reinterpret_cast<EncodeIm2colFn>(fn);
  return 0;
}

// 2D bf16 row-major tensor [outer][inner] with row pitch `ld` elements, 128B-swizzled boxes.
int make_tiled_map(CUtensorMap* m, const void* base, long long inner, long long outer, long long ld, int box_inner,
                   int box_outer) {
  VL_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base pointer must be 16B aligned");
  VL_REQUIRE((ld * 2) % 16 == 0, "TMA row pitch must be a multiple of 16 bytes (ld=%lld)", ld);
  VL_REQUIRE(box_inner * 2 <= 128 && box_outer <= 256 && box_outer >= 1, "bad TMA box %dx%d", box_inner, box_outer);
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode_tiled(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box,
                              estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VL_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d): inner=%lld outer=%lld ld=%lld box=%dx%d", (int)r,
             inner, outer, ld, box_inner, box_outer);
  return 0;
}

// NHWC bf16 tensor seen as (C, W, H, N); im2col boxes of `pixels` x 64 channels, 128B swizzle.
int make_im2col_map(CUtensorMap* m, const void* base, const vl_conv_geom& g, int pixels, int c_dim) {
  VL_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base pointer must be 16B aligned");
  VL_REQUIRE(g.c % 8 == 0, "im2col TMA needs a channel count that is a multiple of 8 (c=%d)", g.c);
  cuuint64_t dims[4] = {(cuuint64_t)c_dim, (cuuint64_t)g.w, (cuuint64_t)g.h, (cuuint64_t)g.n};
  cuuint64_t strides[3] = {(cuuint64_t)g.c * 2, (cuuint64_t)g.c * g.w * 2, (cuuint64_t)g.c * g.w * g.h * 2};
  // TF SAME/VALID geometry: base pixel of output (p,q) is (p*stride - pad_top, q*stride - pad_left); the upper
  // corner bounds the last base pixel so that exactly P x Q positions are traversed per image.
  int lower[2] = {-g.pad_left, -g.pad_top};
  int upper[2] = {(g.q - 1) * g.stride_w - g.pad_left - (g.w - 1), (g.p - 1) * g.stride_h - g.pad_top - (g.h - 1)};
  cuuint32_t estr[4] = {1, (cuuint32_t)g.stride_w, (cuuint32_t)g.stride_h, 1};
  CUresult r = g_encode_im2col(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, lower,
                               upper, 64, (cuuint32_t)pixels, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VL_REQUIRE(r == CUDA_SUCCESS,
             "cuTensorMapEncodeIm2col failed (%d): nhwc=%dx%dx%dx%d lower=(%d,%d) upper=(%d,%d) stride=(%d,%d)", (int)r,
             g.n, g.h, g.w, g.c, lower[0], lower[1], upper[0], upper[1], g.stride_w, g.stride_h);
  return 0;
}

// NHWC bf16 tensor seen as (C, W, H, N) with channel pitch `c_pitch`: tiled boxes of 64 channels x `box_w` pixels of
// one image row, 128B swizzle; channels >= c_valid and pixels outside [0, w) are zero filled.
int make_row_map(CUtensorMap* m, const void* base, long long c_valid, long long c_pitch, int w, int h, int n, int box_w) {
  VL_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (c_pitch * 2) % 16 == 0, "row TMA map: 16B alignment");
  VL_REQUIRE(box_w >= 1 && box_w <= 256, "row TMA map: box of %d pixels", box_w);
  cuuint64_t dims[4] = {(cuuint64_t)c_valid, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)c_pitch * 2, (cuuint64_t)c_pitch * w * 2, (cuuint64_t)c_pitch * w * h * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)box_w, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = g_encode_tiled(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VL_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (row map) failed (%d)", (int)r);
  return 0;
}

int ceil_div(int a, int b) { return (a + b - 1) / b; }

FastDiv make_fastdiv(int d) {
  FastDiv f;
  if (d < 1) d = 1;
  f.d = (uint32_t)d;
  uint32_t l = 0;
  while ((1u << l) < (uint32_t)d) ++l;
  f.s = l;
  f.m = (uint32_t)(((1ull << 32) * ((1ull << l) - (uint64_t)d)) / (uint64_t)d + 1ull);
  return f;
}

}  // namespace

extern "C" int vl_gemm(const vl_gemm_desc* d, const void* a, const void* b, void* c, const float* bias,
                       const void* relu_mask, vl_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VL_REQUIRE(d && a && b && c, "vl_gemm: null argument");
  VL_REQUIRE(d->m > 0 && d->n > 0 && d->k > 0 && d->groups >= 1, "vl_gemm: bad extents m=%d n=%d k=%d groups=%d",
             d->m, d->n, d->k, d->groups);
  if (load_driver_fns() != 0) return -1;

  KParams p;
  memset(&p, 0, sizeof(p));
  const bool a_im2col = d->a_mode == VL_A_IM2COL_K || d->a_mode == VL_A_IM2COL_MN;
  const bool a_mn = d->a_mode == VL_A_TILED_MN || d->a_mode == VL_A_IM2COL_MN;
  const bool b_mn = d->b_mode == VL_B_TILED_MN || d->b_mode == VL_B_IM2COL_MN;
  const bool b_im2col = d->b_mode == VL_B_IM2COL_MN;
  const vl_conv_geom& cg = d->conv;

  p.groups = d->groups;
  p.a_goff = d->a_goff;
  p.b_goff = d->b_goff;
  p.c_goff = d->c_goff;
  p.b_tap_stride = d->b_tap_stride;
  p.b_row_goff = d->b_row_goff;
  p.b_tap_inner = d->b_tap_inner;
  p.taps = 1;
  p.kw = 1;
  p.flip = 0;

  // ---- block_n ----
  int BN = d->block_n;
  if (BN == 0) {
    int nb = ceil_div(d->n, 256);
    int per = ceil_div(d->n, nb);
    int gran = b_mn ? 64 : 16;
    BN = ceil_div(per, gran) * gran;
  }
  VL_REQUIRE(BN >= 16 && BN <= 256 && BN % 16 == 0, "vl_gemm: block_n %d invalid", BN);
  VL_REQUIRE(!b_mn || BN % 64 == 0, "vl_gemm: N-major B needs block_n %% 64 == 0 (got %d)", BN);
  p.BN = BN;
  p.N = d->n;

  // ---- contraction decomposition ----
  if (a_im2col || b_im2col) {
    VL_REQUIRE(cg.kh > 0 && cg.kw > 0 && cg.p > 0 && cg.q > 0 && cg.cin_g > 0, "vl_gemm: bad conv geometry");
    p.taps = cg.kh * cg.kw;
    p.kw = cg.kw;
    p.cchunks = ceil_div(cg.cin_g, 64);
    p.flip = cg.flip_taps;
    p.P = cg.p;
    p.Q = cg.q;
    p.PQ = cg.p * cg.q;
    p.stride_h = cg.stride_h;
    p.stride_w = cg.stride_w;
    p.lower_h = -cg.pad_top;
    p.lower_w = -cg.pad_left;
    p.cin_g = cg.cin_g;
  }
  int m_rows;  // rows of the (per group) output as tiled by 128-row sub-tiles
  if (d->a_mode == VL_A_IM2COL_K) {
    p.M = cg.n * cg.p * cg.q;
    VL_REQUIRE(d->m == p.M, "vl_gemm: m (%d) must equal n*p*q (%d) for im2col A", d->m, p.M);
    p.kb_total = p.taps * p.cchunks;
    m_rows = p.M;
  } else if (d->a_mode == VL_A_IM2COL_MN) {
    // M axis = (tap, 64-channel chunk); K axis = output pixels.
    p.M = p.taps * cg.cin_g;
    VL_REQUIRE(d->k == cg.n * cg.p * cg.q, "vl_gemm: k (%d) must equal n*p*q for transposed im2col A", d->k);
    p.kb_total = ceil_div(d->k, BK);
    m_rows = p.taps * p.cchunks * 64;
  } else {
    p.M = d->m;
    p.kb_total = ceil_div(d->k, BK);
    m_rows = p.M;
    VL_REQUIRE(!(d->b_mode == VL_B_TILED_K && (d->b_tap_stride > 0 || d->b_tap_inner > 0)),
               "vl_gemm: b_tap_stride / b_tap_inner need an im2col A operand");
  }
  if (b_im2col) {
    VL_REQUIRE(d->a_mode == VL_A_TILED_MN, "vl_gemm: transposed im2col B needs an M-major tiled A (dy^T)");
    VL_REQUIRE(d->k == cg.n * cg.p * cg.q, "vl_gemm: k (%d) must equal n*p*q for transposed im2col B", d->k);
    VL_REQUIRE(d->n == p.taps * p.cchunks * 64, "vl_gemm: n must be taps * ceil(cin_g/64) * 64 for transposed im2col B");
    VL_REQUIRE(d->c_dtype == VL_DT_F32 && BN <= 64 * 2 * MAX_MSUB, "vl_gemm: transposed im2col B: fp32 output, block_n <= 256");
    if (d->row_shift) {
      // one k-block per output row; an n-block = the kw taps of one filter row over ONE staged input row
      VL_REQUIRE(cg.q <= 64 && cg.cin_g <= 64 && cg.stride_w == 1 && BN == cg.kw * 64 && d->groups == 1,
                 "vl_gemm: row_shift needs q <= 64, cin_g <= 64, stride_w 1, block_n == kw*64, one group (q=%d cin_g=%d "
                 "block_n=%d)", cg.q, cg.cin_g, BN);
      p.row_shift = 1;
      p.kb_total = cg.n * cg.p;
    }
  }
  // two 128-row sub-tiles per CTA tile when the accumulators fit (2 * BN <= 256 TMEM columns per stage) and there
  // are enough rows to keep every SM busy with the halved tile count
  p.msub = 1;
  if (d->msub == 2 || (d->msub == 0 && 2 * BN <= ACC_STRIDE_COLS && ceil_div(m_rows, BM) >= 4)) p.msub = 2;
  VL_REQUIRE(p.msub * BN <= ACC_STRIDE_COLS, "vl_gemm: msub %d x block_n %d exceeds the TMEM accumulator stage", p.msub,
             BN);
  p.num_m_blk = ceil_div(m_rows, BM * p.msub);
  const int conv_cchunks = p.cchunks;
  if (!a_im2col) p.cchunks = p.kb_total;  // dense k-block walk: tap = 0, cc = kb
  if (b_im2col) p.cchunks = conv_cchunks;  // the chunk -> (tap, channel chunk) decode of the B loads / epilogue
  p.ksteps_tail = BK / 16;
  p.mma_cchunks = p.cchunks;
  if (d->a_mode == VL_A_IM2COL_K)
    p.ksteps_tail = ceil_div(cg.cin_g - (p.cchunks - 1) * 64, 16);
  else if (!a_im2col)
    p.ksteps_tail = ceil_div(d->k - (p.kb_total - 1) * BK, 16);
  if (b_im2col) p.ksteps_tail = BK / 16;
  if (d->a_mode == VL_A_IM2COL_MN || b_im2col) {
    p.mma_cchunks = 1 << 30;  // the contraction runs over pixels: no per-tap tail
    p.mn_step_rows = BK / cg.q;
    p.mn_step_rem = BK % cg.q;
  }
  // B_TILED_K decodes (tap, cc) from kb with p.cchunks; for transposed-im2col A the B operand is always N-major.
  VL_REQUIRE(!(d->a_mode == VL_A_IM2COL_MN && d->b_mode == VL_B_TILED_K),
             "vl_gemm: transposed im2col A requires an N-major B");
  p.num_n_blk = ceil_div(d->n, BN);
  const int out_tiles = p.num_m_blk * p.num_n_blk * p.groups;
  p.split_k = d->split_k;
  if (p.split_k == 0) {
    // automatic split-K for the atomic (filter gradient) epilogue: about two waves of CTAs, >= 4 k-blocks each
    p.split_k = 1;
    if (d->c_atomic && d->c_dtype == VL_DT_F32) {
      int want = (2 * vl::num_sms()) / (out_tiles > 0 ? out_tiles : 1);
      int cap = p.kb_total >= 8 ? p.kb_total / 4 : 1;
      p.split_k = want < 1 ? 1 : (want > cap ? cap : want);
    }
  }
  if (p.split_k < 1) p.split_k = 1;
  if (p.split_k > p.kb_total) p.split_k = p.kb_total;
  p.kb_per_split = ceil_div(p.kb_total, p.split_k);
  p.split_k = ceil_div(p.kb_total, p.kb_per_split);  // no empty splits
  VL_REQUIRE(p.split_k == 1 || (d->c_atomic && d->c_dtype == VL_DT_F32), "vl_gemm: split_k needs fp32 atomic output");
  p.total_tiles = out_tiles * p.split_k;
  p.fd_nblk = make_fastdiv(p.num_n_blk);
  p.fd_mblk = make_fastdiv(p.num_m_blk);
  p.fd_groups = make_fastdiv(p.groups);
  p.fd_cchunks = make_fastdiv(p.cchunks);
  p.fd_kw = make_fastdiv(p.kw);
  p.fd_PQ = make_fastdiv(p.PQ);
  p.fd_Q = make_fastdiv(p.Q);
  p.fd_P = make_fastdiv(p.P);
  p.d2s_c = d->d2s_c;
  if (p.d2s_c > 0) {
    VL_REQUIRE(d->a_mode == VL_A_IM2COL_K && d->b_mode == VL_B_TILED_K && !d->c_atomic && bias == nullptr &&
                   relu_mask == nullptr,
               "vl_gemm: the depth-to-space epilogue serves plain im2col convolutions only");
    VL_REQUIRE(d->d2s_c % 16 == 0 && d->d2s_sh >= 1 && d->d2s_sw >= 1 && d->n == d->d2s_sh * d->d2s_sw * d->d2s_c,
               "vl_gemm: depth-to-space needs n == sh*sw*c and c %% 16 == 0 (n=%d sh=%d sw=%d c=%d)", d->n, d->d2s_sh,
               d->d2s_sw, d->d2s_c);
    VL_REQUIRE(d->d2s_h > 0 && d->d2s_w > 0 && d->c_goff % 8 == 0, "vl_gemm: bad depth-to-space output extent");
    p.d2s_sh = d->d2s_sh;
    p.d2s_sw = d->d2s_sw;
    p.d2s_h = d->d2s_h;
    p.d2s_w = d->d2s_w;
    p.fd_d2s_c = make_fastdiv(p.d2s_c);
    p.fd_d2s_sw = make_fastdiv(p.d2s_sw);
  }

  {
    const char* e = getenv("VL_GEMM_DBG");
    p.dbg = e ? atoi(e) : 0;
    const char* e2 = getenv("VL_GEMM_PRODUCERS");
    p.producers = e2 ? atoi(e2) : 2;
    if (p.producers != 1) p.producers = 2;
  }
  // ---- epilogue ----
  p.C = c;
  p.c_ld = d->c_ld;
  p.c_dtype = d->c_dtype;
  // C must be zeroed by the caller for the atomic epilogue, so a single split may store (vectorised) instead
  p.c_atomic = d->c_atomic && p.split_k > 1;
  p.relu = d->relu;
  p.bias = bias;
  p.mask = reinterpret_cast<const bf16*>(relu_mask);
  p.mask_ld = d->mask_ld;
  VL_REQUIRE(!d->c_atomic || d->c_dtype == VL_DT_F32, "vl_gemm: atomic output must be fp32");

  // ---- smem pipeline ----
  p.b_stage_bytes = BN * 128;
  if (p.row_shift) p.b_stage_bytes = (64 + cg.kw - 1) * 128;  // one staged input row serves the kw taps: smaller stages, deeper ring
  p.stage_bytes = p.msub * A_SUB_BYTES + ((p.b_stage_bytes + 1023) / 1024) * 1024;
  // shared memory left free for CTAs of OTHER kernels on the same SM (vl_set_smem_reserve): the issue-bound LRN / pool
  // gradient kernels need 2 KB per CTA and can then run next to a persistent contraction CTA instead of after it
  int stages = (SMEM_LIMIT - vl::smem_reserve() - 1024 - BAR_REGION) / p.stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  VL_REQUIRE(stages >= 2, "vl_gemm: not enough shared memory for 2 stages");
  p.num_stages = stages;
  const int smem_bytes = 1024 + BAR_REGION + stages * p.stage_bytes;

  // ---- descriptors ----
  // instruction descriptor: D=f32, A=B=bf16, majors, N>>3 at bit 17, M>>4 at bit 24
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
            ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
  // smem descriptor high word: SBO = 1024B (8 rows x 128B) at bits 32..45, version 1 at bit 46, SWIZZLE_128B (2) at 61
  const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
  p.a_desc_hi = hi;
  p.b_desc_hi = hi;
  p.a_lbo_enc = a_mn ? (8192u >> 4) : 1u;
  p.b_lbo_enc = b_mn ? (8192u >> 4) : 1u;
  // row_shift: the N atoms of a filter row OVERLAP, atom s starts one pixel row (128 B) after atom s-1
  // (verified on hardware: profiles/r01_shift_mma_v2.txt, mode 4)
  if (p.row_shift) p.b_lbo_enc = 128u >> 4;
  p.a_kstep_enc = a_mn ? (2048u >> 4) : (32u >> 4);
  p.b_kstep_enc = b_mn ? (2048u >> 4) : (32u >> 4);

  // ---- tensor maps ----
  CUtensorMap tmA, tmB;
  if (d->a_mode == VL_A_TILED_K) {
    long long inner = (long long)d->a_goff * (d->groups - 1) + d->k;
    if (make_tiled_map(&tmA, a, inner, d->m, d->a_ld, 64, BM) != 0) return -1;
  } else if (d->a_mode == VL_A_TILED_MN && p.row_shift) {
    if (make_row_map(&tmA, a, d->m, d->a_ld, cg.q, cg.p, cg.n, 64) != 0) return -1;
  } else if (d->a_mode == VL_A_TILED_MN) {
    long long inner = (long long)d->a_goff * (d->groups - 1) + d->m;
    if (make_tiled_map(&tmA, a, inner, d->k, d->a_ld, 64, 64) != 0) return -1;
  } else {
    if (make_im2col_map(&tmA, a, cg, d->a_mode == VL_A_IM2COL_K ? BM : 64, cg.c) != 0) return -1;
  }
  if (d->b_mode == VL_B_TILED_K) {
    // rows = n (x taps for conv data-gradients, x groups for K-major forward filters), inner = contraction
    long long inner, outer;
    if (d->a_mode == VL_A_IM2COL_K) {
      const long long per_tap = d->b_tap_inner > 0 ? (long long)p.taps * d->b_tap_inner : cg.cin_g;
      inner = (long long)d->b_goff * (d->groups - 1) + per_tap;
      outer = (long long)d->b_row_goff * (d->groups - 1) +
              (d->b_tap_stride > 0 ? (long long)p.taps * d->b_tap_stride : d->n);
    } else {
      inner = (long long)d->b_goff * (d->groups - 1) + d->k;
      outer = (long long)d->b_row_goff * (d->groups - 1) + d->n;
    }
    if (make_tiled_map(&tmB, b, inner, outer, d->b_ld, 64, BN) != 0) return -1;
  } else if (b_im2col && p.row_shift) {
    if (make_row_map(&tmB, b, cg.c, cg.c, cg.w, cg.h, cg.n, 64 + cg.kw - 1) != 0) return -1;
  } else if (b_im2col) {
    if (make_im2col_map(&tmB, b, cg, 64, cg.c) != 0) return -1;
  } else {
    long long inner = (long long)d->b_goff * (d->groups - 1) + d->n;
    long long outer = (d->a_mode == VL_A_IM2COL_K) ? (long long)p.kb_total * 64 : d->k;
    if (make_tiled_map(&tmB, b, inner, outer, d->b_ld, 64, 64) != 0) return -1;
  }

  const int grid = p.total_tiles < vl::num_sms() ? p.total_tiles : vl::num_sms();
#define VL_LAUNCH_GEMM(AMODE, BMODE_)                                                                            \
  do {                                                                                                          \
    static bool attr_set = false;                                                                               \
    if (!attr_set) {                                                                                            \
      VL_CHECK_CUDA(cudaFuncSetAttribute(umma_gemm_kernel<AMODE, BMODE_>,                                       \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));             \
      attr_set = true;                                                                                          \
    }                                                                                                           \
    umma_gemm_kernel<AMODE, BMODE_><<<grid, NUM_THREADS, smem_bytes, stream>>>(tmA, tmB, p);                    \
  } while (0)
  const int combo = b_im2col ? 100 : d->a_mode * 2 + d->b_mode;
  switch (combo) {
    case VL_A_TILED_K * 2 + VL_B_TILED_K: VL_LAUNCH_GEMM(VL_A_TILED_K, VL_B_TILED_K); break;
    case VL_A_TILED_K * 2 + VL_B_TILED_MN: VL_LAUNCH_GEMM(VL_A_TILED_K, VL_B_TILED_MN); break;
    case VL_A_TILED_MN * 2 + VL_B_TILED_K: VL_LAUNCH_GEMM(VL_A_TILED_MN, VL_B_TILED_K); break;
    case VL_A_TILED_MN * 2 + VL_B_TILED_MN: VL_LAUNCH_GEMM(VL_A_TILED_MN, VL_B_TILED_MN); break;
    case VL_A_IM2COL_K * 2 + VL_B_TILED_K: VL_LAUNCH_GEMM(VL_A_IM2COL_K, VL_B_TILED_K); break;
    case VL_A_IM2COL_K * 2 + VL_B_TILED_MN: VL_LAUNCH_GEMM(VL_A_IM2COL_K, VL_B_TILED_MN); break;
    case VL_A_IM2COL_MN * 2 + VL_B_TILED_MN: VL_LAUNCH_GEMM(VL_A_IM2COL_MN, VL_B_TILED_MN); break;
    case 100: VL_LAUNCH_GEMM(VL_A_TILED_MN, VL_B_IM2COL_MN); break;
    default: VL_REQUIRE(false, "vl_gemm: unsupported operand mode combination a=%d b=%d", d->a_mode, d->b_mode);
  }
#undef VL_LAUNCH_GEMM
  vl::g_launches.fetch_add(1);
  VL_CHECK_CUDA(cudaGetLastError());
  return 0;
}
