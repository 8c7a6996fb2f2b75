// The one exchange step of the path: after every rank has encoded + mapped its own images, the prefix
// embeddings [B_local, K, d] are all-gathered so that every rank holds the whole [N * B_local, K, d] tensor. In place:
// each rank's mapper writes its block straight into slot `rank` of the gathered buffer (cc_mapper_forward takes the
// output pointer), and one ncclAllGather over NVLink / NVSwitch fills in the peers' slots. The reference has no
// multi-GPU inference path at all (single `--device`, clipcap/inference/args.py:22-27).
//
// NCCL is bound at run time (dlopen of libnccl.so.2 — the copy PyTorch already mapped into the process when there is one),
// so libclipcap_b200.so itself links against nothing but the CUDA runtime and loads on boxes without NCCL.
#include <dlfcn.h>

#include <cstring>

#include "common.h"

struct cc_comm {
  void* comm = nullptr;  // ncclComm_t
  int rank = 0, nranks = 1, device = 0;
};

namespace cc {
namespace {

struct NcclId {
  char internal[128];
};
typedef int (*GetUniqueIdFn)(NcclId*);
typedef int (*CommInitRankFn)(void**, int, NcclId, int);
typedef int (*AllGatherFn)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef int (*CommDestroyFn)(void*);
typedef const char* (*GetErrorStringFn)(int);
typedef int (*GetVersionFn)(int*);

struct Nccl {
  void* lib = nullptr;
  GetUniqueIdFn get_unique_id = nullptr;
  CommInitRankFn comm_init_rank = nullptr;
  AllGatherFn all_gather = nullptr;
  CommDestroyFn comm_destroy = nullptr;
  GetErrorStringFn error_string = nullptr;
  GetVersionFn get_version = nullptr;
  bool ok = false;
};

const Nccl& nccl() {
  static const Nccl n = [] {
    Nccl x;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      x.lib = dlopen(nm, RTLD_NOW | RTLD_NOLOAD);  // already in the process (PyTorch's bundled copy)?
      if (x.lib == nullptr) x.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (x.lib != nullptr) break;
    }
    if (x.lib == nullptr) return x;
    x.get_unique_id = reinterpret_cast<GetUniqueIdFn>(dlsym(x.lib, "ncclGetUniqueId"));
    x.comm_init_rank = reinterpret_cast<CommInitRankFn>(dlsym(x.lib, "ncclCommInitRank"));
    x.all_gather = reinterpret_cast<AllGatherFn>(dlsym(x.lib, "ncclAllGather"));
    x.comm_destroy = reinterpret_cast<CommDestroyFn>(dlsym(x.lib, "ncclCommDestroy"));
    x.error_string = reinterpret_cast<GetErrorStringFn>(dlsym(x.lib, "ncclGetErrorString"));
    x.get_version = reinterpret_cast<GetVersionFn>(dlsym(x.lib, "ncclGetVersion"));
    x.ok = x.get_unique_id && x.comm_init_rank && x.all_gather && x.comm_destroy && x.error_string;
    return x;
  }();
  return n;
}

#define CC_NCCL(expr)                                                                              \
  do {                                                                                             \
    const int r__ = (expr);                                                                        \
    if (r__ != 0) {                                                                                \
      cc::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cc::nccl().error_string(r__));   \
      return CC_ENCCL;                                                                             \
    }                                                                                              \
  } while (0)

int require_nccl() {
  CC_REQUIRE(nccl().ok, CC_ENCCL, "libnccl.so.2 could not be loaded (dlopen) or lacks the collective entry points");
  return CC_OK;
}

}  // namespace
}  // namespace cc

extern "C" {

int cc_comm_unique_id(void* id_out) {
  using namespace cc;
  CC_REQUIRE(id_out != nullptr, CC_EINVAL, "cc_comm_unique_id: null argument");
  CC_TRY(require_nccl());
  CC_NCCL(nccl().get_unique_id(static_cast<NcclId*>(id_out)));
  return CC_OK;
}

int cc_comm_create(cc_comm** out, const void* unique_id, int rank, int nranks) {
  using namespace cc;
  CC_REQUIRE(out != nullptr && unique_id != nullptr, CC_EINVAL, "cc_comm_create: null argument");
  *out = nullptr;
  CC_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, CC_EINVAL, "cc_comm_create: rank %d of %d", rank, nranks);
  CC_TRY(require_nccl());
  CC_TRY(check_device_sm100());
  cc_comm* c = new cc_comm();
  c->rank = rank;
  c->nranks = nranks;
  cudaGetDevice(&c->device);
  NcclId id;
  memcpy(&id, unique_id, sizeof(id));
  const int r = nccl().comm_init_rank(&c->comm, nranks, id, rank);
  if (r != 0) {
    set_error("ncclCommInitRank(rank %d of %d) -> %s", rank, nranks, nccl().error_string(r));
    delete c;
    return CC_ENCCL;
  }
  *out = c;
  return CC_OK;
}

int cc_allgather_prefix(cc_comm* c, void* prefix_all, size_t bytes_per_rank, void* stream) {
  using namespace cc;
  CC_REQUIRE(c != nullptr && prefix_all != nullptr, CC_EINVAL, "cc_allgather_prefix: null argument");
  CC_REQUIRE(bytes_per_rank > 0, CC_ESHAPE, "cc_allgather_prefix: empty shard");
  if (c->nranks == 1) return CC_OK;  // the local block is the whole tensor
  // in place: this rank's block already sits at slot `rank` (ncclAllGather's documented in-place form)
  const char* send = static_cast<const char*>(prefix_all) + static_cast<size_t>(c->rank) * bytes_per_rank;
  CC_NCCL(nccl().all_gather(send, prefix_all, bytes_per_rank, /*ncclInt8*/ 0, c->comm, static_cast<cudaStream_t>(stream)));
  return CC_OK;
}

int cc_comm_rank(cc_comm* c) { return c ? c->rank : -1; }
int cc_comm_nranks(cc_comm* c) { return c ? c->nranks : 0; }

int cc_nccl_version(void) {
  int v = 0;
  if (cc::nccl().ok && cc::nccl().get_version != nullptr) cc::nccl().get_version(&v);
  return v;
}

void cc_comm_destroy(cc_comm* c) {
  if (c == nullptr) return;
  if (c->comm != nullptr && cc::nccl().ok) cc::nccl().comm_destroy(c->comm);
  delete c;
}

}  // extern "C"
