// Host-side plumbing shared by the engines: thread-local error message, device arena, arch check, weight lookup.
#include <cstdlib>

#include "common.h"

#include <atomic>
#include <cstdarg>
#include <cstring>

namespace cc {

namespace {
thread_local char g_err[1024] = "";
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

int DevBuf::alloc(size_t n) {
  release();
  if (n == 0) return CC_OK;
  cudaError_t e = cudaMalloc(&p, n);
  if (e != cudaSuccess) {
    p = nullptr;
    set_error("cudaMalloc(%zu) failed: %s", n, cudaGetErrorString(e));
    return CC_ENOMEM;
  }
  bytes = n;
  return CC_OK;
}
void DevBuf::release() {
  if (p) cudaFree(p);
  p = nullptr;
  bytes = 0;
}

bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("CLIPCAP_B200_NO_PDL");
    return !(e != nullptr && e[0] == '1');
  }();
  return on;
}

int Arena::alloc(void** out, size_t bytes) {
  *out = nullptr;
  if (bytes == 0) bytes = 16;
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) {
    set_error("cudaMalloc(%zu) failed after %zu bytes in this handle: %s", bytes, total, cudaGetErrorString(e));
    (void)cudaGetLastError();
    return CC_ENOMEM;
  }
  ptrs.push_back(p);
  total += bytes;
  *out = p;
  return CC_OK;
}
void Arena::release() {
  for (void* p : ptrs) cudaFree(p);
  ptrs.clear();
  total = 0;
}

int check_device_sm100() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_error("no CUDA device: %s (libclipcap_b200 has no CPU fallback)", cudaGetErrorString(e));
    (void)cudaGetLastError();
    return CC_ECUDA;
  }
  int major = 0, minor = 0;
  CC_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  CC_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (major != 10) {
    set_error("device %d is sm_%d%d; libclipcap_b200 only contains sm_100a code (no fallback)", dev, major, minor);
    return CC_EARCH;
  }
  return CC_OK;
}

namespace {
std::atomic<int> g_sm_budget{[] {
  const char* e = getenv("CLIPCAP_B200_SM_BUDGET");  // development: size every launch for n SMs without partitioning
  return e != nullptr ? atoi(e) : 0;
}()};
}

void set_sm_budget(int n) { g_sm_budget.store(n > 0 ? n : 0); }
int sm_budget() { return g_sm_budget.load(); }

int device_sms() {  // per device: a process may drive several GPUs
  static std::atomic<int> cache[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  int n = cache[dev].load();
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache[dev].store(n);
  }
  return n;
}

int num_sms() {
  const int n = device_sms(), b = g_sm_budget.load();
  return (b > 0 && b < n) ? b : n;
}

int find_weight(const cc_tensor* w, int n, const std::string& name, int64_t expect_numel, Arena& arena,
                const float** out) {
  *out = nullptr;
  for (int i = 0; i < n; ++i) {
    if (w[i].name == nullptr || name != w[i].name) continue;
    CC_REQUIRE(w[i].dtype == CC_F32, CC_EINVAL, "weight '%s': only fp32 state_dict tensors are accepted", name.c_str());
    CC_REQUIRE(w[i].ndim >= 1 && w[i].ndim <= 4, CC_ESHAPE, "weight '%s': ndim %d", name.c_str(), w[i].ndim);
    int64_t numel = 1;
    for (int k = 0; k < w[i].ndim; ++k) numel *= w[i].shape[k];
    CC_REQUIRE(numel == expect_numel, CC_ESHAPE, "weight '%s': %lld elements, expected %lld", name.c_str(),
               (long long)numel, (long long)expect_numel);
    CC_REQUIRE(w[i].data != nullptr, CC_EINVAL, "weight '%s': null data", name.c_str());
    cudaPointerAttributes attr;
    cudaError_t e = cudaPointerGetAttributes(&attr, w[i].data);
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      attr.type = cudaMemoryTypeUnregistered;
    }
    if (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged) {
      *out = static_cast<const float*>(w[i].data);
      return CC_OK;
    }
    float* d = nullptr;
    CC_TRY(arena.alloc_t(&d, static_cast<size_t>(numel)));
    CC_CUDA(cudaMemcpy(d, w[i].data, static_cast<size_t>(numel) * sizeof(float), cudaMemcpyHostToDevice));
    *out = d;
    return CC_OK;
  }
  set_error("weight '%s' not found among the %d tensors passed", name.c_str(), n);
  return CC_EINVAL;
}

}  // namespace cc
