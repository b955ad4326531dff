// Tap-shifted stride-1 convolution on the tensor cores (sm_100a): the narrow-channel layers of the AlexNet encoder.
//
//   conv1 (as 3x3 VALID over the space-to-depth input, 48 -> 96)     models/alexnet/alexnet.py:60-77
//   conv2 (5x5 SAME, 2 groups of 48 -> 128)                          models/alexnet/alexnet.py:100-118
//   conv2 data gradient (5x5 "full" correlation of dy with the flipped filter, 2 groups of 128 -> 48)
//
// The im2col formulation of gemm_umma.cu re-reads every input pixel once per filter tap through im2col-mode TMA
// (9x / 25x, measured 42-58 B/clk/SM) and puts the pixels on the M side, where a 128 x N x 16 UMMA costs ~92 cycles
// however small N (= output channels) is.  Here
//   * ONE tiled 4-D TMA box per (tile, 64-channel chunk) brings the input rows of the tile -- with the SAME padding
//     materialised by TMA's out-of-bounds zero fill -- into shared memory as a flat [rows][Wp] x 128 B array;
//   * every filter tap (r, s) addresses that array through the SAME shared-memory descriptor shifted by
//     (r * Wp + s) * 128 bytes (the 128B swizzle phase comes from the absolute address, so any row shift is legal:
//     profiles/r01_shift_mma.txt);
//   * operands are swapped: output channels (<= 128) sit on the M side, R x Wp flat positions (<= 256) on the N side,
//     so each UMMA is 128 x N x 16 at N/2 cycles; the accumulator tile is D^T[channel][position] in TMEM;
//   * positions whose column falls into the padding are computed and dropped by the epilogue (conv2: 28 of 32).
// Weights stream through their own ring (one 128 x 64 K-major box per tap and chunk).
#include "common.cuh"
#include "../../include/vlb200.h"

#include <atomic>
#include <cstdlib>
#include <cstring>

namespace vl {
extern std::atomic<long long> g_launches;
}

namespace {

using namespace vl::ptx;
typedef __nv_bfloat16 bf16;

constexpr int EPI_GROUPS = 4;                           // epilogue groups of four warps (one per TMEM lane quadrant); p.epi_groups of them work
constexpr int NUM_THREADS = 64 + EPI_GROUPS * 128;      // warp0 TMA, warp1 MMA (+TMEM alloc), then the epilogue warps
constexpr int TMEM_COLS = 512;    // 2 accumulator stages x 256 fp32 columns
constexpr int ACC_STRIDE_COLS = 256;
constexpr int MAX_W_STAGES = 8;
constexpr int MAX_X_STAGES = 2;
constexpr int SMEM_LIMIT = 232448;
constexpr int BAR_REGION = 1024;

struct FastDiv {
  uint32_t m, s, d;
};
__device__ __forceinline__ void fd_divmod(int n, const FastDiv& f, int& q, int& r) {
  q = (int)((__umulhi((uint32_t)n, f.m) + (uint32_t)n) >> f.s);
  r = n - q * (int)f.d;
}

struct FParams {
  int n_img, Ho, Wo, Wp, R, npos;     // output extent, padded row width, output rows per tile, positions per tile (N)
  int row_tiles, groups, m_blks, total_tiles;
  int M;                               // output channels per group
  int cin_g, cchunks, taps, kw, flip;
  int pad_top, pad_left;
  int a_goff;                          // per-group channel offset into the input tensor
  int w_row_goff;                      // per-group row offset into the K-major filter
  int c_goff, c_ld;
  int x_stage_bytes, x_box_bytes, x_stages, w_stages;
  int tpg;                             // filter taps per weight-ring stage (one hand-shake per group of taps)
  int w_tap_bytes;                     // bytes of one (tap, chunk) filter box: rows x 128 B
  int w_region;                        // bytes of the filter area (ring or resident copy)
  int resident;                        // 1: the whole filter stays in shared memory (loaded once per CTA)
  FastDiv fd_mblk, fd_groups, fd_rt, fd_kw;
  void* C;
  int relu;
  const float* bias;
  uint32_t idesc;
  int dbg;
  int x_loads, x_load_rows;            // the input tile arrives as x_loads TMA boxes of x_load_rows rows each
  int epi_groups;                      // epilogue groups that drain the accumulator (<= EPI_GROUPS)
  int stage_bufs;                      // transposition buffers per group: 2 (double buffered) or 1 (shared memory is tight)
  int tma_store;                       // 1: the epilogue leaves through TMA tensor stores (tmY)
  int direct;                          // 1: the epilogue stores 2-byte elements straight from registers (default)
};

// 16 positions x 128 channels of the epilogue stage -> out[n][h][w .. w+15][c .. c+127]; elements outside the tensor
// (channels beyond c_ld, columns outside [0, Wo), rows beyond Ho) are clipped by the TMA unit.
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, uint32_t src, int32_t c0, int32_t c1, int32_t c2,
                                             int32_t c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

__device__ __forceinline__ void tma_load_4d_u32(uint32_t dst, const CUtensorMap* m, uint32_t bar, int32_t c0, int32_t c1,
                                                int32_t c2, int32_t c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      :
      : "r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

struct FTile {
  int m_blk, g, rt, img;
};
__device__ __forceinline__ FTile decode(const FParams& p, int tile) {
  FTile t;
  int r;
  fd_divmod(tile, p.fd_mblk, r, t.m_blk);
  fd_divmod(r, p.fd_groups, r, t.g);
  fd_divmod(r, p.fd_rt, t.img, t.rt);
  return t;
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
    conv_flat_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                     const __grid_constant__ CUtensorMap tmY, const __grid_constant__ FParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  // the 1024-byte alignment is applied to the __shared__ array itself so that the compiler keeps the shared address
  // space (LDS/STS instead of generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* w_full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* w_empty = w_full + MAX_W_STAGES;
  uint64_t* x_full = w_empty + MAX_W_STAGES;
  uint64_t* x_empty = x_full + MAX_X_STAGES;
  uint64_t* tmem_full = x_empty + MAX_X_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint8_t* x_tiles = smem + BAR_REGION;
  uint8_t* w_tiles = x_tiles + p.x_stages * p.x_stage_bytes;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmW);
    for (int s = 0; s < p.w_stages; ++s) {
      mbar_init(&w_full[s], 1);
      mbar_init(&w_empty[s], 1);
    }
    for (int s = 0; s < p.x_stages; ++s) {
      mbar_init(&x_full[s], 1);
      mbar_init(&x_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 4 * p.epi_groups);
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
  const uint32_t x_u32 = smem_u32(x_tiles), w_u32 = smem_u32(w_tiles);
  const uint32_t wf_u32 = smem_u32(w_full), we_u32 = smem_u32(w_empty);
  const uint32_t xf_u32 = smem_u32(x_full), xe_u32 = smem_u32(x_empty);
  const int total_tiles = p.total_tiles;
  const int taps = p.taps, cchunks = p.cchunks;
  const int w_stages = p.w_stages, x_stages = p.x_stages;

  if (warp == 0) {
    // ===================== TMA producer =====================
    const bool ld_w = !(p.dbg & 4);
    int ws = 0, xs = 0;
    uint32_t wph = 0, xph = 0;
    if (p.resident) {
      // the whole filter (taps x chunks boxes) is loaded once and stays resident for every tile of this CTA
      if (elect_one()) {
        mbar_expect_tx_u32(wf_u32, ld_w ? (uint32_t)(taps * cchunks * p.w_tap_bytes) : 0u);
        if (ld_w)
          for (int i = 0; i < taps * cchunks; ++i) {  // box i = (chunk i / taps, tap i % taps [flipped])
            const int tp = i % taps, tapw = p.flip ? (taps - 1 - tp) : tp;
            tma_load_2d_u32(w_u32 + i * p.w_tap_bytes, &tmW, wf_u32, (tapw * cchunks + i / taps) * 64, 0);
          }
      }
      __syncwarp();
    }
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const FTile t = decode(p, tile);
      const int row0 = t.rt * p.R - p.pad_top;  // first input row of the tile
      const int w_row = t.g * p.w_row_goff + t.m_blk * 128;
      for (int cc = 0; cc < cchunks; ++cc) {
        mbar_wait_u32(xe_u32 + xs * 8, xph ^ 1u);
        if (elect_one()) {
          mbar_expect_tx_u32(xf_u32 + xs * 8, (p.dbg & 2) ? 0u : (uint32_t)p.x_box_bytes);
          if (!(p.dbg & 2))
            for (int j = 0; j < p.x_loads; ++j)
              tma_load_4d_u32(x_u32 + xs * p.x_stage_bytes + j * p.x_load_rows * p.Wp * 128, &tmX, xf_u32 + xs * 8,
                              t.g * p.a_goff + cc * 64, -p.pad_left, row0 + j * p.x_load_rows, t.img);
        }
        __syncwarp();
        if (++xs == x_stages) {
          xs = 0;
          xph ^= 1u;
        }
        // filter column of tap t and chunk cc: (t * cchunks + cc) * 64, walked forwards or (flipped) backwards
        int kcoord = (p.flip ? (taps - 1) * cchunks + cc : cc) * 64;
        const int kstep = (p.flip ? -cchunks : cchunks) * 64;
        for (int tap0 = 0; tap0 < taps && !p.resident; tap0 += p.tpg) {
          const int nt = min(p.tpg, taps - tap0);
          mbar_wait_u32(we_u32 + ws * 8, wph ^ 1u);
          if (elect_one()) {
            mbar_expect_tx_u32(wf_u32 + ws * 8, ld_w ? (uint32_t)(nt * p.w_tap_bytes) : 0u);
            if (ld_w) {
              uint32_t dst = w_u32 + ws * p.tpg * p.w_tap_bytes;
              int kc = kcoord;
              for (int j = 0; j < nt; ++j, dst += p.w_tap_bytes, kc += kstep)
                tma_load_2d_u32(dst, &tmW, wf_u32 + ws * 8, kc, w_row);
            }
          }
          kcoord += nt * kstep;
          __syncwarp();
          if (++ws == w_stages) {
            ws = 0;
            wph ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // One hand-shake per GROUP of taps; inside a group the elected lane issues every UMMA back to back with
    // descriptors advanced by adds only (a single warp retires one dependent instruction every ~5 cycles, so the
    // instruction count of this loop is what the tensor pipe waits for between groups).
    const uint32_t idesc = p.idesc;
    // K-major, SWIZZLE_128B: SBO = 1024 B (8 rows), LBO unused (1), descriptor version 1
    const uint64_t hi = (static_cast<uint64_t>((1024u >> 4) | (1u << 14) | (2u << 29)) << 32) | (1ull << 16);
    const int kw = p.kw, tpg = p.tpg;
    const uint32_t row_skip = (uint32_t)(p.Wp - kw) * 8u;  // extra 16-byte units from the end of a filter row to the next
    const bool do_mma = !(p.dbg & 1);
    int ws = 0, xs = 0;
    uint32_t wph = 0, xph = 0;
    int local = 0;
    const int tpg_eff = p.resident ? taps : tpg;  // resident filter: one group = all taps, no weight hand-shake
    if (p.resident) mbar_wait_u32(wf_u32, 0);
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      mbar_wait_u32(smem_u32(&tmem_empty[acc]), acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * ACC_STRIDE_COLS;
      uint32_t accumulate = 0;
      for (int cc = 0; cc < cchunks; ++cc) {
        mbar_wait_u32(xf_u32 + xs * 8, xph);
        const int ksteps = min(4, (p.cin_g - cc * 64 + 15) >> 4);  // zero-padded K-steps are never issued
        uint32_t b_lo = ((x_u32 + xs * p.x_stage_bytes) >> 4) & 0x3FFFu;  // descriptor start of tap (0,0), 16-byte units
        int ts = 0;
        for (int tap0 = 0; tap0 < taps; tap0 += tpg_eff) {
          const int nt = min(tpg_eff, taps - tap0);
          if (!p.resident) mbar_wait_u32(wf_u32 + ws * 8, wph);
          tc_fence_after();
          uint32_t a_lo = ((w_u32 + (p.resident ? cc * taps : ws * tpg) * p.w_tap_bytes) >> 4) & 0x3FFFu;
          if (elect_one()) {
            if (do_mma) {
              for (int j = 0; j < nt; ++j) {
                const uint64_t adesc = hi | a_lo, bdesc = hi | b_lo;  // the tap = a row shift of the same input tile
                if (ksteps == 3) {
                  umma_bf16(tmem_d, adesc, bdesc, idesc, accumulate);
                  umma_bf16(tmem_d, adesc + 2, bdesc + 2, idesc, 1u);
                  umma_bf16(tmem_d, adesc + 4, bdesc + 4, idesc, 1u);
                } else if (ksteps == 4) {
                  umma_bf16(tmem_d, adesc, bdesc, idesc, accumulate);
                  umma_bf16(tmem_d, adesc + 2, bdesc + 2, idesc, 1u);
                  umma_bf16(tmem_d, adesc + 4, bdesc + 4, idesc, 1u);
                  umma_bf16(tmem_d, adesc + 6, bdesc + 6, idesc, 1u);
                } else {
                  for (int k = 0; k < ksteps; ++k)
                    umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, accumulate | (uint32_t)k);
                }
                accumulate = 1;
                a_lo += (uint32_t)(p.w_tap_bytes >> 4);
                if (!(p.dbg & 128)) {  // probe bit 128: every tap reads the un-shifted (atom-aligned) tile
                  b_lo += 8;
                  if (++ts == kw) {
                    ts = 0;
                    b_lo += row_skip;
                  }
                }
              }
            }
            if (!p.resident) umma_commit_u32(we_u32 + ws * 8);
            if (tap0 + nt == taps) umma_commit_u32(xe_u32 + xs * 8);  // the input tile is free after its last tap
          }
          // (accumulate, a_lo, b_lo, ts advance in the elected lane only: elect.sync picks the same lane every time)
          __syncwarp();
          if (++ws == w_stages) {
            ws = 0;
            wph ^= 1u;
          }
        }
        if (++xs == x_stages) {
          xs = 0;
          xph ^= 1u;
        }
      }
      if (elect_one()) umma_commit_u32(smem_u32(&tmem_full[acc]));
      __syncwarp();
    }
  } else if (((warp - 2) >> 2) < p.epi_groups) {
    // ===================== epilogue: D^T[channel lane][flat position] -> NHWC =====================
    // Each thread owns one output channel (TMEM lane).  16 positions at a time are transposed through a small
    // double-buffered shared-memory stage ([position][channel] bf16) so that global stores are 16-byte vectors over
    // the channel axis (one contiguous 2*mc-byte run per position) instead of 2-byte scatters.
    // epi_groups x 4 epilogue warps: warp % 4 selects the TMEM lane quadrant, (warp - 2) / 4 the group; group g
    // drains the 16-position chunks g, g + epi_groups, ... (a single warp sustains only ~650 cycles per chunk, so a
    // short-K tile - conv1: 27 UMMAs = 3.2 k cycles for 15 chunks - needs four groups to stay under its main loop).
    const int quad = warp & 3;
    const int grp = (warp - 2) >> 2;
    const int etid = (threadIdx.x - 64) & 127;  // 0..127 inside the group
    const int m_local = quad * 32 + lane;
    bf16* out = reinterpret_cast<bf16*>(p.C);
    const int ngrp = p.epi_groups;
    const int bufs = p.stage_bufs;
    bf16* stage = reinterpret_cast<bf16*>(w_tiles + p.w_region) + grp * (bufs * 16 * 128);  // [bufs][16][128] per group
    int local = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
      const FTile t = decode(p, tile);
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      const int mc = min(128, p.M - t.m_blk * 128);  // channels of this m-block (multiple of 8)
      const int cpr = mc >> 3;                       // 16-byte chunks per position
      const bool m_ok = m_local < mc;
      const float bv = (m_ok && p.bias != nullptr) ? __ldg(p.bias + t.g * p.c_goff + t.m_blk * 128 + m_local) : 0.f;
      const int orow0 = t.rt * p.R;
      const int rows_here = min(p.R, p.Ho - orow0);
      bf16* obase = out + ((long long)t.img * p.Ho + orow0) * p.Wo * p.c_ld + t.g * p.c_goff + t.m_blk * 128;
      // the (position, chunk) items this thread copies out of the stage: item = etid and etid + 128
      int it_pos[2], it_ch[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int item = etid + u * 128;
        it_pos[u] = item / cpr;
        it_ch[u] = item - it_pos[u] * cpr;
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(quad * 32) << 16) + acc * ACC_STRIDE_COLS;
      // (row, col) of the first position of this group's current 16-position chunk
      int row_c = 0, col_c = grp * 16;
      while (col_c >= p.Wp) {
        col_c -= p.Wp;
        ++row_c;
      }
      // one 16-position chunk: registers -> stage[buf] -> global (static register indexing: no local memory)
      auto emit = [&](const uint32_t(&v)[16], int buf) {
        if (p.direct) {
          // Direct epilogue: lane = channel, so the 32 lanes of a warp write 64 contiguous bytes per position; no
          // shared-memory transposition, no barrier between the four warps of a group - every epilogue warp runs on
          // its own, and the short-K layers (conv1: 27 UMMAs per tile) are bound by exactly this code.
          if (m_ok && !(p.dbg & 8)) {
            int col = col_c, row = row_c;
            int off = (row_c * p.Wo + col_c) * p.c_ld + m_local;  // tile-local element offset (fits 32 bits)
            const int wrap_fix = (p.Wo - p.Wp) * p.c_ld;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float f = __uint_as_float(v[j]) + bv;
              if (p.relu) f = fmaxf(f, 0.f);
              if (col < p.Wo && row < rows_here) obase[off] = __float2bfloat16_rn(f);
              ++col;
              off += p.c_ld;
              if (col == p.Wp) {
                col = 0;
                ++row;
                off += wrap_fix;
              }
            }
          }
          col_c += 16 * ngrp;
          while (col_c >= p.Wp) {
            col_c -= p.Wp;
            ++row_c;
          }
          return;
        }
        bf16* sb = stage + (bufs == 2 ? buf : 0) * (16 * 128);
        if (m_ok) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float f = __uint_as_float(v[j]) + bv;
            if (p.relu) f = fmaxf(f, 0.f);
            sb[j * 128 + m_local] = __float2bfloat16_rn(f);
          }
        }
        if (p.tma_store) {
          // the stage leaves through the TMA unit: no per-thread addressing, junk columns / channels clipped by the
          // tensor map.  Thread 0 of the group first waits until the store issued two chunks ago (same buffer as the
          // NEXT chunk) has read its data, so that passing this barrier also frees that buffer.
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          if (etid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory");
          if (etid == 0 && !(p.dbg & 8)) {
            const uint32_t src = smem_u32(sb);
            const int cch = t.g * p.c_goff + t.m_blk * 128;
            if (row_c < rows_here && !(p.dbg & 32)) tma_store_4d(&tmY, src, cch, col_c, orow0 + row_c, t.img);
            if (col_c + 16 > p.Wp && row_c + 1 < rows_here && !(p.dbg & 16))
              tma_store_4d(&tmY, src, cch, col_c - p.Wp, orow0 + row_c + 1, t.img);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          col_c += 16 * ngrp;
          while (col_c >= p.Wp) {
            col_c -= p.Wp;
            ++row_c;
          }
          return;
        }
        asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory");
        if (!(p.dbg & 8)) {
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            if (it_pos[u] < 16) {
              int col = col_c + it_pos[u], row = row_c;
              while (col >= p.Wp) {
                col -= p.Wp;
                ++row;
              }
              if (col < p.Wo && row < rows_here) {
                const uint4 val = *reinterpret_cast<const uint4*>(sb + it_pos[u] * 128 + it_ch[u] * 8);
                *reinterpret_cast<uint4*>(obase + (long long)(row * p.Wo + col) * p.c_ld + it_ch[u] * 8) = val;
              }
            }
          }
        }
        if (bufs == 1) asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory");  // single stage: reads done before the next chunk's writes
        col_c += 16 * ngrp;  // the other groups handle the chunks in between
        while (col_c >= p.Wp) {
          col_c -= p.Wp;
          ++row_c;
        }
      };
      uint32_t va[16], vb[16];
      const int first = grp * 16;
      if (first < p.npos && !(p.dbg & 64)) tmem_ld_x16(taddr + first, va);
      const int STEP = 16 * ngrp;
      for (int c0 = first; c0 < p.npos && !(p.dbg & 64); c0 += 2 * STEP) {
        tmem_ld_wait();
        const bool has_b = c0 + STEP < p.npos;
        if (has_b) tmem_ld_x16(taddr + c0 + STEP, vb);
        emit(va, 0);
        if (has_b) {
          tmem_ld_wait();
          if (c0 + 2 * STEP < p.npos) tmem_ld_x16(taddr + c0 + 2 * STEP, va);
          emit(vb, 1);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
  }

  if (p.tma_store && warp >= 2 && ((threadIdx.x - 64) & 127) == 0)
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // the stage must outlive the last stores
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// dst[g*cin_g + ci][tap*kpad + co] = src[tap*cin_g + ci][g*cout_g + co] (0 for co >= cout_g): the K-major operand of
// the data gradient (rows = input channels, contraction = (tap, output channel of the group)).
__global__ void pack_dgrad_kmajor_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int taps, int cin_g,
                                         int cout_g, int groups, int kpad, long long total) {
  const int ld = taps * kpad;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int row = (int)(idx / ld);
    const int k = (int)(idx - (long long)row * ld);
    const int tap = k / kpad, co = k - tap * kpad;
    const int g = row / cin_g, ci = row - g * cin_g;
    float v = 0.f;
    if (co < cout_g) v = src[((long long)tap * cin_g + ci) * (groups * cout_g) + g * cout_g + co];
    dst[idx] = __float2bfloat16_rn(v);
  }
}


// K-major filter of the depth-to-space data gradient (include/vlb200.h: vl_pack_dgrad_d2s):
// dst[g*sh*sw*cin_g + (dy*sw+dx)*cin_g + c][(ty*kw2+tx)*kpad + k] = src[kh-1+dy-ty][kw-1+dx-tx][c][g*cout_g + k]
// (0 outside the filter or for k >= cout_g); consecutive threads walk k, i.e. the contiguous axis of both tensors.
__global__ void pack_dgrad_d2s_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int kh, int kw, int cin_g,
                                      int cout_g, int groups, int sh, int sw, int kpad, long long total) {
  const int kh2 = kh + sh - 1, kw2 = kw + sw - 1;
  const int ld = kh2 * kw2 * kpad;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int row = (int)(idx / ld);
    const int col = (int)(idx - (long long)row * ld);
    const int tap = col / kpad, k = col - tap * kpad;
    const int ty = tap / kw2, tx = tap - ty * kw2;
    const int per_g = sh * sw * cin_g;
    const int g = row / per_g;
    const int rr = row - g * per_g;
    const int seg = rr / cin_g, c = rr - seg * cin_g;
    const int dy = seg / sw, dx = seg - dy * sw;
    const int r = kh - 1 + dy - ty, q = kw - 1 + dx - tx;
    float v = 0.f;
    if (k < cout_g && r >= 0 && r < kh && q >= 0 && q < kw)
      v = src[(((long long)r * kw + q) * cin_g + c) * (groups * cout_g) + g * cout_g + k];
    dst[idx] = __float2bfloat16_rn(v);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;

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

extern "C" int vl_conv_flat(const vl_conv_flat_desc* d, const void* x, const void* w_kmajor, const float* bias, void* out,
                            vl_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VL_REQUIRE(d && x && w_kmajor && out, "vl_conv_flat: null argument");
  VL_REQUIRE(d->n > 0 && d->h > 0 && d->w > 0 && d->c > 0 && d->kh > 0 && d->kw > 0 && d->groups >= 1,
             "vl_conv_flat: bad geometry");
  VL_REQUIRE(d->c % 8 == 0 && d->w_ld % 8 == 0 && d->c_ld % 8 == 0 && d->cout_g % 8 == 0 &&
                 (reinterpret_cast<uintptr_t>(out) & 15) == 0,
             "vl_conv_flat: channel counts / pitches must be multiples of 8 and out 16-byte aligned");
  if (!g_encode) {
    cudaDriverEntryPointQueryResult q;
    void* fn = nullptr;
    VL_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    VL_REQUIRE(fn != nullptr && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available");
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  FParams p;
  memset(&p, 0, sizeof(p));
  const int pad_bottom = d->pad_bottom, pad_right = d->pad_right;
  p.n_img = d->n;
  p.Wp = d->w + d->pad_left + pad_right;
  p.Ho = d->h + d->pad_top + pad_bottom - d->kh + 1;
  p.Wo = p.Wp - d->kw + 1;
  VL_REQUIRE(p.Ho > 0 && p.Wo > 0 && p.Wp <= 256, "vl_conv_flat: output extent %dx%d, padded width %d", p.Ho, p.Wo, p.Wp);
  p.R = 256 / p.Wp;
  if (p.R > p.Ho) p.R = p.Ho;
  VL_REQUIRE(p.R >= 1, "vl_conv_flat: padded row wider than 256 positions");
  // balance the row tiles of an image (conv2: 28 rows -> 4 x 7 rather than 3 x 8 + 4)
  p.row_tiles = (p.Ho + p.R - 1) / p.R;
  p.R = (p.Ho + p.row_tiles - 1) / p.row_tiles;
  p.npos = ((p.R * p.Wp + 15) / 16) * 16;
  VL_REQUIRE(p.npos <= 256, "vl_conv_flat: %d positions per tile", p.npos);
  p.groups = d->groups;
  p.M = d->cout_g;
  p.m_blks = (d->cout_g + 127) / 128;
  p.total_tiles = p.n_img * p.row_tiles * p.groups * p.m_blks;
  p.cin_g = d->cin_g;
  p.cchunks = (d->cin_g + 63) / 64;
  p.taps = d->kh * d->kw;
  p.kw = d->kw;
  p.flip = d->flip_taps;
  p.pad_top = d->pad_top;
  p.pad_left = d->pad_left;
  p.a_goff = d->cin_g;
  p.w_row_goff = d->cout_g;
  p.c_goff = d->cout_g;
  p.c_ld = d->c_ld;
  p.C = out;
  p.relu = d->relu;
  p.bias = bias;
  p.fd_mblk = make_fastdiv(p.m_blks);
  p.fd_groups = make_fastdiv(p.groups);
  p.fd_rt = make_fastdiv(p.row_tiles);
  p.fd_kw = make_fastdiv(p.kw);
  {
    const char* e = getenv("VL_GEMM_DBG");
    p.dbg = e ? atoi(e) : 0;
  }
  // input rows per tile: R + kh - 1; the last taps of the (dropped) padding columns read a few positions past the
  // box, which stay inside the next stage / the slack at the end of the allocation
  const int box_rows = p.R + d->kh - 1;
  VL_REQUIRE(box_rows <= 256, "vl_conv_flat: %d input rows per tile", box_rows);
  p.x_box_bytes = box_rows * p.Wp * 128;
  p.x_stage_bytes = ((p.x_box_bytes + 1023) / 1024) * 1024;
  p.x_stages = MAX_X_STAGES;
  // several smaller boxes per tile keep more requests of the TMA unit in flight (VL_FLAT_XSPLIT = boxes per tile)
  p.x_loads = 1;
  if (getenv("VL_FLAT_XSPLIT")) p.x_loads = atoi(getenv("VL_FLAT_XSPLIT"));
  if (p.x_loads < 1 || box_rows % p.x_loads != 0) p.x_loads = 1;
  p.x_load_rows = box_rows / p.x_loads;
  // Epilogue modes.  Default: 16 positions x 128 channels are transposed through a shared-memory stage and leave
  // through ONE TMA tensor store per chunk when the padded row width is a multiple of 16 (conv2: chunks never straddle
  // rows; a store with a negative start column faults), else through per-thread 16-byte stores.  Measured at 1024
  // frames (conv1 / conv2 forward, us): no epilogue at all 277 / 380, stage + 16-byte stores 377 / 424, stage + TMA
  // store 342 (second store of straddling chunks skipped, i.e. incomplete) / 414, VL_FLAT_EPI=direct (2-byte stores
  // straight from registers, lane = channel) 536 / 416: the 64-byte partial-line writes of the direct form are what
  // the transposition avoids, and conv1 is bound by its main loop (277 us against a 163 us UMMA floor) before
  // anything else.
  const char* epi_env = getenv("VL_FLAT_EPI");
  p.direct = (epi_env && !strcmp(epi_env, "direct")) ? 1 : 0;
  // epilogue groups: a short contraction (conv1: 27 UMMAs per 15-chunk tile) is bound by the accumulator drain, which
  // scales with the number of groups; long contractions keep two groups and more shared memory for the filter ring
  const int umma_per_tile = d->kh * d->kw * ((d->cin_g + 15) / 16);
  p.epi_groups = umma_per_tile <= 48 ? 4 : 2;
  if (getenv("VL_FLAT_GROUPS")) p.epi_groups = atoi(getenv("VL_FLAT_GROUPS"));
  if (p.epi_groups < 1) p.epi_groups = 1;
  if (p.epi_groups > EPI_GROUPS) p.epi_groups = EPI_GROUPS;
  p.stage_bufs = 2;
  int STAGE_BYTES = p.direct ? 0 : p.epi_groups * p.stage_bufs * 16 * 128 * 2;  // transposition stages per group
  // (vl::smem_reserve(): shared memory left to CTAs of other kernels on the same SM, see vl_set_smem_reserve)
  const int smem_limit = SMEM_LIMIT - vl::smem_reserve();
  int w_avail = smem_limit - 1024 - BAR_REGION - STAGE_BYTES - 1024 - p.x_stages * p.x_stage_bytes;
  // filter box: only the rows that exist (8-row swizzle atoms); the UMMA reads 128 rows, the surplus lanes are
  // never stored
  int w_box_rows = d->w_rows < 128 ? ((d->w_rows + 7) / 8) * 8 : 128;
  if (p.m_blks > 1 || p.groups > 1) w_box_rows = 128;
  p.w_tap_bytes = w_box_rows * 128;
  // resident filter when it fits (conv1: 9 taps x 96 rows = 108 KB) and is the same for every tile
  const int resident_bytes = p.taps * p.cchunks * p.w_tap_bytes + (128 - w_box_rows) * 128;
  if (!p.direct && p.m_blks == 1 && p.groups == 1 && resident_bytes > w_avail &&
      resident_bytes <= w_avail + STAGE_BYTES / 2) {
    // the resident filter fits only next to single-buffered transposition stages (conv1 with four groups)
    p.stage_bufs = 1;
    STAGE_BYTES /= 2;
    w_avail += STAGE_BYTES;
  }
  p.resident = (p.m_blks == 1 && p.groups == 1 && resident_bytes <= w_avail && !getenv("VL_FLAT_NO_RESIDENT")) ? 1 : 0;
  // taps per weight stage: a whole filter row if two such stages fit (one mbarrier hand-shake costs ~350 cycles, a
  // 128 x 240 x 16 UMMA 120), else as many taps as still allow double buffering
  p.tpg = d->kw;
  if (getenv("VL_FLAT_TPG")) p.tpg = atoi(getenv("VL_FLAT_TPG"));
  while (p.tpg > 1 && w_avail / (p.tpg * p.w_tap_bytes) < 2) --p.tpg;
  int w_stages = w_avail / (p.tpg * p.w_tap_bytes);
  if (w_stages > MAX_W_STAGES) w_stages = MAX_W_STAGES;
  int w_region = w_stages * p.tpg * p.w_tap_bytes;
  if (p.resident) {
    w_stages = 1;
    w_region = ((p.taps * p.cchunks * p.w_tap_bytes + (128 - w_box_rows) * 128 + 1023) / 1024) * 1024;
  }
  VL_REQUIRE(w_stages >= 2 || p.resident, "vl_conv_flat: not enough shared memory (input tile %d B)", p.x_stage_bytes);
  p.w_stages = w_stages;
  p.w_region = w_region;
  const int smem_bytes = 1024 + BAR_REGION + p.x_stages * p.x_stage_bytes + w_region + STAGE_BYTES + 1024;
  // instruction descriptor: D=f32, A=B=bf16 K-major, N>>3 at bit 17, M>>4 at bit 24
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.npos >> 3) << 17) | ((128u >> 4) << 24);

  CUtensorMap tmX, tmW, tmY;
  // epilogue through TMA stores: a 128-channel box must not spill into the channels of another group / m-block
  // (a 16-position chunk may touch at most two output rows: Wp >= 16)
  // (TMA stores fault on a negative start coordinate, measured: chunks must not straddle rows -> Wp % 16 == 0)
  p.tma_store = ((d->cout_g % 128 == 0) || d->groups == 1) && (d->c_ld % 8 == 0) && p.Wp % 16 == 0 &&
                        ((reinterpret_cast<uintptr_t>(out) & 15) == 0) && !p.direct && !(epi_env && !strcmp(epi_env, "stage"))
                    ? 1
                    : 0;
  if (p.stage_bufs == 1) p.tma_store = 0;  // the TMA-store form relies on the double buffer
  {
    cuuint64_t dims[4] = {(cuuint64_t)d->c_ld, (cuuint64_t)p.Wo, (cuuint64_t)p.Ho, (cuuint64_t)d->n};
    cuuint64_t strides[3] = {(cuuint64_t)d->c_ld * 2, (cuuint64_t)d->c_ld * p.Wo * 2,
                             (cuuint64_t)d->c_ld * p.Wo * p.Ho * 2};
    cuuint32_t box[4] = {128, 16, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = g_encode(&tmY, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, out, dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      p.tma_store = 0;  // e.g. an exotic pitch: the per-thread store path serves it
      memset(&tmY, 0, sizeof(tmY));
    }
  }
  {
    cuuint64_t dims[4] = {(cuuint64_t)d->c, (cuuint64_t)d->w, (cuuint64_t)d->h, (cuuint64_t)d->n};
    cuuint64_t strides[3] = {(cuuint64_t)d->c * 2, (cuuint64_t)d->c * d->w * 2, (cuuint64_t)d->c * d->w * d->h * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)p.Wp, (cuuint32_t)p.x_load_rows, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = g_encode(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VL_REQUIRE(r == CUDA_SUCCESS, "vl_conv_flat: input tensor map failed (%d)", (int)r);
  }
  {
    const long long k_total = (long long)p.taps * p.cchunks * 64;
    cuuint64_t dims[2] = {(cuuint64_t)k_total, (cuuint64_t)d->w_rows};
    cuuint64_t strides[1] = {(cuuint64_t)d->w_ld * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)w_box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w_kmajor), dims, strides, box,
                          estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VL_REQUIRE(r == CUDA_SUCCESS, "vl_conv_flat: filter tensor map failed (%d)", (int)r);
  }
  static bool attr_set = false;
  if (!attr_set) {
    VL_CHECK_CUDA(cudaFuncSetAttribute(conv_flat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    attr_set = true;
  }
  const int grid = p.total_tiles < vl::num_sms() ? p.total_tiles : vl::num_sms();
  conv_flat_kernel<<<grid, NUM_THREADS, smem_bytes, stream>>>(tmX, tmW, tmY, p);
  vl::g_launches.fetch_add(1);
  VL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int vl_pack_dgrad_kmajor(const float* src, void* dst, int32_t taps, int32_t cin_g, int32_t cout_g,
                                    int32_t groups, vl_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VL_REQUIRE(src && dst && taps > 0 && cin_g > 0 && cout_g > 0 && groups > 0, "vl_pack_dgrad_kmajor: bad arguments");
  const int kpad = ((cout_g + 63) / 64) * 64;
  const long long total = (long long)groups * cin_g * taps * kpad;
  pack_dgrad_kmajor_kernel<<<(int)((total + 255) / 256), 256, 0, stream>>>(src, reinterpret_cast<bf16*>(dst), taps,
                                                                           cin_g, cout_g, groups, kpad, total);
  vl::g_launches.fetch_add(1);
  VL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int vl_pack_dgrad_d2s(const float* src, void* dst, int32_t kh, int32_t kw, int32_t cin_g, int32_t cout_g,
                                 int32_t groups, int32_t sh, int32_t sw, vl_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VL_REQUIRE(src && dst && kh > 0 && kw > 0 && cin_g > 0 && cout_g > 0 && groups > 0 && sh >= 1 && sw >= 1,
             "vl_pack_dgrad_d2s: bad arguments");
  const int kpad = ((cout_g + 63) / 64) * 64;
  const long long total = (long long)groups * sh * sw * cin_g * (kh + sh - 1) * (kw + sw - 1) * kpad;
  long long blocks = (total + 255) / 256;
  if (blocks > 65535LL * 16) blocks = 65535LL * 16;
  pack_dgrad_d2s_kernel<<<(int)blocks, 256, 0, stream>>>(src, reinterpret_cast<bf16*>(dst), kh, kw, cin_g, cout_g, groups,
                                                         sh, sw, kpad, total);
  vl::g_launches.fetch_add(1);
  VL_CHECK_CUDA(cudaGetLastError());
  return 0;
}
