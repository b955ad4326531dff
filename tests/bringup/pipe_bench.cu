// Throughput probe of the SM pipes the HBM-bound LRN / pool kernels lean on (sm_100a): warp-instructions per clock
// per SM for FFMA / FMUL / FADD (3-register forms), packed FFMA2, MUFU, HFMA2.BF16, HSET2, PRMT, LOP3 and two mixes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_bench pipe_bench.cu ; run: ./pipe_bench
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

#define ITERS 4096
#define CHAINS 8

#define KERNEL(NAME, DECL, BODY, SINK)                                         \
  __global__ void __launch_bounds__(256) NAME(float* out, float seed) {        \
    DECL;                                                                      \
    _Pragma("unroll 1") for (int it = 0; it < ITERS; ++it) {                   \
      _Pragma("unroll") for (int c = 0; c < CHAINS; ++c) { BODY; }             \
    }                                                                          \
    float acc = 0.f;                                                           \
    _Pragma("unroll") for (int c = 0; c < CHAINS; ++c) acc += SINK;            \
    if (acc == 123.456f) out[threadIdx.x] = acc;                               \
  }

#define FDECL float v[CHAINS]; float a = seed, b = seed * 0.5f; for (int c = 0; c < CHAINS; ++c) v[c] = seed + c + threadIdx.x
#define UDECL uint32_t v[CHAINS]; uint32_t a = __float_as_uint(seed), b = a * 3u; for (int c = 0; c < CHAINS; ++c) v[c] = a + c + threadIdx.x
#define DDECL unsigned long long v[CHAINS]; unsigned long long a = __float_as_uint(seed) * 0x100000001ull, b = a * 3u; for (int c = 0; c < CHAINS; ++c) v[c] = a + c + threadIdx.x

KERNEL(k_ffma, FDECL, asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(v[c]) : "f"(a), "f"(b)), v[c])
KERNEL(k_fmul, FDECL, asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(v[c]) : "f"(a)), v[c])
KERNEL(k_fadd, FDECL, asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(v[c]) : "f"(a)), v[c])
KERNEL(k_ffma2, DDECL, asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[c]) : "l"(a), "l"(b)), (float)(v[c] & 0xff))
KERNEL(k_fmul2, DDECL, asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(v[c]) : "l"(a)), (float)(v[c] & 0xff))
KERNEL(k_fadd2, DDECL, asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(v[c]) : "l"(a)), (float)(v[c] & 0xff))
KERNEL(k_rsq, FDECL, asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(v[c])), v[c])
KERNEL(k_sqrt, FDECL, asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(v[c])), v[c])
KERNEL(k_hfma2, UDECL, asm volatile("fma.rn.bf16x2 %0, %0, %1, %2;" : "+r"(v[c]) : "r"(a), "r"(b)), (float)v[c])
KERNEL(k_hset2, UDECL, asm volatile("set.eq.bf16x2.bf16x2 %0, %0, %1;" : "+r"(v[c]) : "r"(a)), (float)v[c])
KERNEL(k_hmax2, UDECL, asm volatile("max.bf16x2 %0, %0, %1;" : "+r"(v[c]) : "r"(a)), (float)v[c])
KERNEL(k_prmt, UDECL, asm volatile("prmt.b32 %0, %0, %1, 0x3120;" : "+r"(v[c]) : "r"(a)), (float)v[c])
KERNEL(k_lop3, UDECL, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[c]) : "r"(a), "r"(b)), (float)v[c])
KERNEL(k_shl, UDECL, asm volatile("shl.b32 %0, %0, 16;" : "+r"(v[c])), (float)v[c])
KERNEL(k_cvtpack, UDECL, asm volatile("{.reg .f32 t; mov.b32 t, %0; cvt.rn.bf16x2.f32 %0, t, t;}" : "+r"(v[c])), (float)v[c])
// mixes: one FFMA + one LOP3 per chain (do fma and alu pipes overlap?), FFMA + MUFU, FFMA2 + LOP3
KERNEL(k_mix_ffma_lop, FDECL; uint32_t u[CHAINS]; for (int c = 0; c < CHAINS; ++c) u[c] = threadIdx.x + c,
       asm volatile("fma.rn.f32 %0, %0, %2, %3;\n\tlop3.b32 %1, %1, %4, %4, 0x96;" : "+f"(v[c]), "+r"(u[c]) : "f"(a), "f"(b), "r"(__float_as_uint(a))),
       v[c] + (float)u[c])
KERNEL(k_mix_ffma_rsq, FDECL; float u[CHAINS]; for (int c = 0; c < CHAINS; ++c) u[c] = threadIdx.x + c + 1.f,
       asm volatile("fma.rn.f32 %0, %0, %2, %3;\n\trsqrt.approx.ftz.f32 %1, %1;" : "+f"(v[c]), "+f"(u[c]) : "f"(a), "f"(b)),
       v[c] + u[c])
KERNEL(k_mix_4ffma_rsq, FDECL; float u[CHAINS]; for (int c = 0; c < CHAINS; ++c) u[c] = threadIdx.x + c + 1.f,
       asm volatile("fma.rn.f32 %0, %0, %2, %3;\n\tfma.rn.f32 %0, %0, %2, %3;\n\tfma.rn.f32 %0, %0, %2, %3;\n\tfma.rn.f32 %0, %0, %2, %3;\n\trsqrt.approx.ftz.f32 %1, %1;" : "+f"(v[c]), "+f"(u[c]) : "f"(a), "f"(b)),
       v[c] + u[c])
KERNEL(k_mix_ffma2_lop, DDECL; uint32_t u[CHAINS]; for (int c = 0; c < CHAINS; ++c) u[c] = threadIdx.x + c,
       asm volatile("fma.rn.f32x2 %0, %0, %2, %3;\n\tlop3.b32 %1, %1, %4, %4, 0x96;" : "+l"(v[c]), "+r"(u[c]) : "l"(a), "l"(b), "r"((uint32_t)a)),
       (float)(v[c] & 0xff) + (float)u[c])
KERNEL(k_mix_ffma_hfma2, FDECL; uint32_t u[CHAINS]; for (int c = 0; c < CHAINS; ++c) u[c] = threadIdx.x + c,
       asm volatile("fma.rn.f32 %0, %0, %2, %3;\n\tfma.rn.bf16x2 %1, %1, %4, %4;" : "+f"(v[c]), "+r"(u[c]) : "f"(a), "f"(b), "r"(__float_as_uint(a))),
       v[c] + (float)u[c])

template <typename K>
void run(const char* name, K kern, int per_iter, float* out, int sms, double ghz) {
  const int ctas = sms * 8;
  kern<<<ctas, 256>>>(out, 1.5f);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  kern<<<ctas, 256>>>(out, 1.5f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double warp_insts = (double)ctas * 8 * ITERS * CHAINS * per_iter;
  const double clk = ms * 1e-3 * ghz * 1e9;
  printf("%-18s %8.3f ms  %6.3f warp-inst/clk/SM  (%.2f per SMSP)\n", name, ms, warp_insts / clk / sms, warp_insts / clk / sms / 4);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double ghz = khz * 1e-6;
  printf("%s, %d SMs, %.3f GHz (max); rates assume the max clock\n", p.name, p.multiProcessorCount, ghz);
  float* out;
  cudaMalloc(&out, 4096);
  const int s = p.multiProcessorCount;
  run("FFMA", k_ffma, 1, out, s, ghz);
  run("FMUL", k_fmul, 1, out, s, ghz);
  run("FADD", k_fadd, 1, out, s, ghz);
  run("FFMA2 (f32x2)", k_ffma2, 1, out, s, ghz);
  run("FMUL2 (f32x2)", k_fmul2, 1, out, s, ghz);
  run("FADD2 (f32x2)", k_fadd2, 1, out, s, ghz);
  run("MUFU.RSQ", k_rsq, 1, out, s, ghz);
  run("MUFU.SQRT", k_sqrt, 1, out, s, ghz);
  run("HFMA2.BF16", k_hfma2, 1, out, s, ghz);
  run("HSET2.BF16 eq", k_hset2, 1, out, s, ghz);
  run("HMNMX2.BF16", k_hmax2, 1, out, s, ghz);
  run("PRMT", k_prmt, 1, out, s, ghz);
  run("LOP3", k_lop3, 1, out, s, ghz);
  run("SHL", k_shl, 1, out, s, ghz);
  run("F2FP.BF16 pack", k_cvtpack, 1, out, s, ghz);
  run("FFMA+LOP3", k_mix_ffma_lop, 2, out, s, ghz);
  run("FFMA+MUFU", k_mix_ffma_rsq, 2, out, s, ghz);
  run("4FFMA+MUFU", k_mix_4ffma_rsq, 5, out, s, ghz);
  run("FFMA2+LOP3", k_mix_ffma2_lop, 2, out, s, ghz);
  run("FFMA+HFMA2", k_mix_ffma_hfma2, 2, out, s, ghz);
  return 0;
}
