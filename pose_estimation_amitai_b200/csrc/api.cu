// Error plumbing and device queries of the C ABI (include/poseb200.h).
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace pb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  cudaGetLastError();  // clear the (non-sticky) error so the next launch check is not poisoned
  set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
  return PB_ERR_CUDA;
}

bool is_device_ptr(const void* p) {
  cudaPointerAttributes attr;
  cudaError_t e = cudaPointerGetAttributes(&attr, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
}

static unsigned long long g_launches = 0;
void note_launches(int n) { __atomic_fetch_add(&g_launches, (unsigned long long)n, __ATOMIC_RELAXED); }
unsigned long long launches() { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
      cudaGetLastError();
      n = 148;
    }
  }
  return n;
}

}  // namespace pb

extern "C" {

const char* pb_last_error_string(void) { return pb::g_err; }

int pb_abi_version(void) { return PB_ABI_VERSION; }

int pb_launch_count(unsigned long long* out) {
  if (out == nullptr) return PB_ERR_INVALID;
  *out = pb::launches();
  return PB_OK;
}

void pb_note_launches(int32_t n) {
  if (n > 0) pb::note_launches(n);
}

int pb_device_info(char* name, int len, int* sm, int* n_sm) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return pb::cuda_fail(e, "pb_device_info: cudaGetDevice (no GPU; there is no CPU fallback)");
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) return pb::cuda_fail(e, "pb_device_info");
  if (name && len > 0) {
    strncpy(name, prop.name, (size_t)len - 1);
    name[len - 1] = 0;
  }
  if (sm) *sm = prop.major * 10 + prop.minor;
  if (n_sm) *n_sm = prop.multiProcessorCount;
  return PB_OK;
}

}  // extern "C"
