// Host runtime glue of the C-ABI: error channel, device query, launch counter.
#include "common.cuh"
#include "../../include/vlb200.h"

#include <atomic>
#include <stdarg.h>

namespace vl {

std::atomic<long long> g_launches{0};

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 148;
    n = prop.multiProcessorCount;
  }
  return n;
}

static std::atomic<int> g_smem_reserve{0};
int smem_reserve() { return g_smem_reserve.load(); }

}  // namespace vl

extern "C" int vl_set_smem_reserve(int32_t bytes) {
  VL_REQUIRE(bytes >= 0 && bytes <= 96 * 1024, "vl_set_smem_reserve: 0..96 KB");
  vl::g_smem_reserve.store((bytes + 1023) / 1024 * 1024);
  return 0;
}

// Pinned host staging memory.  write_combined = 1: cudaHostAllocWriteCombined - not snooped by the CPU caches, which
// leaves more of the host's memory / PCIe bandwidth to the device reads when several ranks copy at once; the host must
// only WRITE to it (reads are uncached and slow).
extern "C" void* vl_host_alloc(int64_t bytes, int32_t write_combined) {
  void* p = nullptr;
  const unsigned flags = cudaHostAllocPortable | (write_combined ? cudaHostAllocWriteCombined : 0u);
  if (bytes <= 0 || cudaHostAlloc(&p, (size_t)bytes, flags) != cudaSuccess) {
    vl::set_error("vl_host_alloc: cudaHostAlloc of %lld bytes failed", (long long)bytes);
    return nullptr;
  }
  return p;
}

extern "C" int vl_host_free(void* p) {
  if (p != nullptr) VL_CHECK_CUDA(cudaFreeHost(p));
  return 0;
}

extern "C" const char* vl_last_error(void) { return vl::last_error(); }
extern "C" int vl_version(void) { return 100; }
extern "C" int vl_device_sm_count(void) { return vl::num_sms(); }
extern "C" int64_t vl_launch_count(void) { return vl::g_launches.load(); }

extern "C" int vl_zero(void* ptr, int64_t bytes, vl_stream_t stream_) {
  VL_REQUIRE(ptr != nullptr && bytes >= 0, "vl_zero: bad arguments");
  VL_CHECK_CUDA(cudaMemsetAsync(ptr, 0, (size_t)bytes, reinterpret_cast<cudaStream_t>(stream_)));
  return 0;
}
