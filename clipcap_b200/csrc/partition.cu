// SM partitions (CUDA green contexts) for the two-stage serving pipeline.
//
// The hot path has two stages of opposite character: the image tower + mapper + prefill are tensor-bound (and, on B200,
// power-bound: the chip sits at its 1000 W cap with the SM clock near 1.4 GHz), the 19 decode steps are a chain of ~3200
// dependent launches bound by per-kernel latency that leaves most of the machine idle. cc_partition_create splits the
// device's SMs into a large and a small group; a caller runs the tensor-bound stages of batch i+1 on the large
// partition's stream while the decode loop of batch i runs on the small one (clipcap_b200/pipeline.py). The reference
// has no counterpart: it decodes one image at a time on one stream (clipcap/inference/demo.py:30-45).
//
// Driver entry points are resolved at run time (cudaGetDriverEntryPoint), so the library carries no link-time
// dependency on libcuda symbols newer than the runtime it was built with.
#include "common.h"

struct cc_partition {
  CUgreenCtx ctx[2] = {nullptr, nullptr};  // 0 = large, 1 = small
  CUstream stream[2] = {nullptr, nullptr};
  int sms[2] = {0, 0};
  int device = 0;
};

namespace cc {
namespace {

typedef CUresult (*DeviceGetFn)(CUdevice*, int);
typedef CUresult (*GetDevResourceFn)(CUdevice, CUdevResource*, CUdevResourceType);
typedef CUresult (*SplitByCountFn)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int,
                                   unsigned int);
typedef CUresult (*GenerateDescFn)(CUdevResourceDesc*, CUdevResource*, unsigned int);
typedef CUresult (*GreenCtxCreateFn)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int);
typedef CUresult (*GreenCtxDestroyFn)(CUgreenCtx);
typedef CUresult (*GreenCtxStreamCreateFn)(CUstream*, CUgreenCtx, unsigned int, int);
typedef CUresult (*StreamDestroyFn)(CUstream);

struct DriverFns {
  DeviceGetFn device_get = nullptr;
  GetDevResourceFn get_res = nullptr;
  SplitByCountFn split = nullptr;
  GenerateDescFn gen_desc = nullptr;
  GreenCtxCreateFn ctx_create = nullptr;
  GreenCtxDestroyFn ctx_destroy = nullptr;
  GreenCtxStreamCreateFn stream_create = nullptr;
  StreamDestroyFn stream_destroy = nullptr;
  bool ok = false;
};

template <class F>
bool resolve(const char* name, F* out) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess ||
      p == nullptr) {
    (void)cudaGetLastError();
    return false;
  }
  *out = reinterpret_cast<F>(p);
  return true;
}

const DriverFns& driver() {
  static const DriverFns fns = [] {
    DriverFns f;
    f.ok = resolve("cuDeviceGet", &f.device_get) && resolve("cuDeviceGetDevResource", &f.get_res) &&
           resolve("cuDevSmResourceSplitByCount", &f.split) && resolve("cuDevResourceGenerateDesc", &f.gen_desc) &&
           resolve("cuGreenCtxCreate", &f.ctx_create) && resolve("cuGreenCtxDestroy", &f.ctx_destroy) &&
           resolve("cuGreenCtxStreamCreate", &f.stream_create) && resolve("cuStreamDestroy", &f.stream_destroy);
    return f;
  }();
  return fns;
}

#define CC_DRV(expr)                                                                     \
  do {                                                                                   \
    CUresult r__ = (expr);                                                               \
    if (r__ != CUDA_SUCCESS) {                                                           \
      cc::set_error("%s:%d: %s -> CUresult %d", __FILE__, __LINE__, #expr, (int)r__);    \
      return CC_ECUDA;                                                                   \
    }                                                                                    \
  } while (0)

int partition_build(cc_partition* p, int device, int small_sms) {
  const DriverFns& d = driver();
  CC_REQUIRE(d.ok, CC_ECUDA, "cc_partition_create: this driver does not export the green-context entry points");
  CC_CUDA(cudaSetDevice(device));
  CC_CUDA(cudaFree(nullptr));  // the primary context exists
  CC_TRY(check_device_sm100());
  p->device = device;
  CUdevice dev;
  CC_DRV(d.device_get(&dev, device));
  CUdevResource all;
  CC_DRV(d.get_res(dev, &all, CU_DEV_RESOURCE_TYPE_SM));
  const int total = static_cast<int>(all.sm.smCount);
  CC_REQUIRE(small_sms >= 2 && small_sms <= total - 8, CC_EINVAL,
             "cc_partition_create: the small partition needs 2 .. %d SMs (got %d)", total - 8, small_sms);
  CUdevResource small_res, rest;
  unsigned int groups = 1;
  // Multiples of 8 SMs split along the GPC hierarchy (the default). Any other count asks the driver to treat the SMs
  // independently of their hierarchy (finer partitions; thread-block clusters larger than a CTA pair are then not
  // guaranteed inside a partition — this library launches none).
  const unsigned int flags = (small_sms % 8 == 0) ? 0u : static_cast<unsigned int>(CU_DEV_SM_RESOURCE_SPLIT_IGNORE_SM_COSCHEDULING);
  CC_DRV(d.split(&small_res, &groups, &all, &rest, flags, static_cast<unsigned int>(small_sms)));
  CC_REQUIRE(groups == 1 && small_res.sm.smCount > 0 && rest.sm.smCount > 0, CC_ECUDA,
             "cc_partition_create: could not split %d SMs into %d + rest", total, small_sms);
  CUdevResource res[2] = {rest, small_res};
  for (int i = 0; i < 2; ++i) {
    CUdevResourceDesc desc;
    CC_DRV(d.gen_desc(&desc, &res[i], 1));
    CC_DRV(d.ctx_create(&p->ctx[i], desc, dev, CU_GREEN_CTX_DEFAULT_STREAM));
    CC_DRV(d.stream_create(&p->stream[i], p->ctx[i], CU_STREAM_NON_BLOCKING, 0));
    p->sms[i] = static_cast<int>(res[i].sm.smCount);
  }
  return CC_OK;
}

void partition_free(cc_partition* p) {
  if (p == nullptr) return;
  const DriverFns& d = driver();
  if (d.ok) {
    for (int i = 0; i < 2; ++i) {
      if (p->stream[i]) d.stream_destroy(p->stream[i]);
      if (p->ctx[i]) d.ctx_destroy(p->ctx[i]);
    }
  }
  delete p;
}

}  // namespace
}  // namespace cc

extern "C" {

int cc_partition_create(cc_partition** out, int device, int small_sms) {
  using namespace cc;
  CC_REQUIRE(out != nullptr, CC_EINVAL, "cc_partition_create: null argument");
  *out = nullptr;
  cc_partition* p = new cc_partition();
  const int st = partition_build(p, device, small_sms);
  if (st != CC_OK) {
    partition_free(p);
    return st;
  }
  *out = p;
  return CC_OK;
}

void* cc_partition_stream(cc_partition* p, int which) {
  return (p != nullptr && (which == 0 || which == 1)) ? static_cast<void*>(p->stream[which]) : nullptr;
}

int cc_partition_sms(cc_partition* p, int which) { return (p != nullptr && (which == 0 || which == 1)) ? p->sms[which] : 0; }

void cc_partition_destroy(cc_partition* p) { cc::partition_free(p); }

}  // extern "C"
