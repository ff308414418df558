// Data-parallel collective inside the library: ONE all-reduce(SUM, fp32) of the flat
// gradient buffer per optimizer step over NCCL (NVLink 5 / NVSwitch), one process per GPU.  The reference has no
// multi-GPU mode; summing over ranks reproduces its gradient accumulation over mini_batch_size mini-batches
// (/root/reference/models/AcousticModel.py:386-401).
//
// NCCL is bound at run time with dlopen (the library links nothing but cudart): the copy already mapped into the
// process (torch's bundled libnccl.so.2) is preferred, then the loader's search path.  Only five entry points are
// used: ncclGetUniqueId, ncclCommInitRank, ncclAllReduce, ncclCommDestroy, ncclGetErrorString.
#include "common.cuh"
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

namespace rs {
namespace {

typedef struct { char internal[128]; } NcclUniqueId;         // ncclUniqueId: NCCL_UNIQUE_ID_BYTES = 128
typedef void* NcclComm;
typedef int (*GetUniqueIdFn)(NcclUniqueId*);
typedef int (*CommInitRankFn)(NcclComm*, int, NcclUniqueId, int);
typedef int (*AllReduceFn)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t);
typedef int (*CommDestroyFn)(NcclComm);
typedef const char* (*GetErrorStringFn)(int);
constexpr int kNcclFloat32 = 7, kNcclSum = 0;                 // ncclDataType_t / ncclRedOp_t values (nccl.h)

struct NcclApi {
  void* handle = nullptr;
  GetUniqueIdFn get_unique_id = nullptr;
  CommInitRankFn comm_init_rank = nullptr;
  AllReduceFn all_reduce = nullptr;
  CommDestroyFn comm_destroy = nullptr;
  GetErrorStringFn error_string = nullptr;
};

NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.handle ? &api : nullptr;
  tried = true;
  const char* names[] = {getenv("RS_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) {
    if (!n || !n[0]) continue;
    h = dlopen(n, RTLD_NOW | RTLD_NOLOAD);                   // already mapped (torch's bundled copy)?
    if (!h) h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) return nullptr;
  api.get_unique_id = (GetUniqueIdFn)dlsym(h, "ncclGetUniqueId");
  api.comm_init_rank = (CommInitRankFn)dlsym(h, "ncclCommInitRank");
  api.all_reduce = (AllReduceFn)dlsym(h, "ncclAllReduce");
  api.comm_destroy = (CommDestroyFn)dlsym(h, "ncclCommDestroy");
  api.error_string = (GetErrorStringFn)dlsym(h, "ncclGetErrorString");
  if (!api.get_unique_id || !api.comm_init_rank || !api.all_reduce || !api.comm_destroy) return nullptr;
  api.handle = h;
  return &api;
}

struct Comm {
  NcclComm comm;
  int rank, world;
};

int nccl_fail(NcclApi* api, const char* what, int rc) {
  set_error("%s: NCCL error %d (%s)", what, rc, api->error_string ? api->error_string(rc) : "?");
  return RS_ERR_CUDA;
}

}  // namespace
}  // namespace rs

using namespace rs;

extern "C" int rs_comm_unique_id(void* id_out, size_t id_bytes) {
  RS_REQUIRE(id_out && id_bytes >= sizeof(NcclUniqueId), RS_ERR_INVALID, "rs_comm_unique_id: need a %zu-byte buffer", sizeof(NcclUniqueId));
  NcclApi* api = nccl_api();
  RS_REQUIRE(api != nullptr, RS_ERR_UNSUPPORTED, "rs_comm_unique_id: libnccl.so.2 not found (set RS_NCCL_LIB)");
  NcclUniqueId id;
  const int rc = api->get_unique_id(&id);
  if (rc != 0) return nccl_fail(api, "ncclGetUniqueId", rc);
  memcpy(id_out, &id, sizeof(id));
  return RS_OK;
}

extern "C" int rs_comm_init(void** comm_out, const void* unique_id, size_t id_bytes, int rank, int world) {
  RS_REQUIRE(comm_out && unique_id && id_bytes >= sizeof(NcclUniqueId), RS_ERR_INVALID, "rs_comm_init: bad argument");
  RS_REQUIRE(world >= 1 && rank >= 0 && rank < world, RS_ERR_INVALID, "rs_comm_init: rank %d of %d", rank, world);
  NcclApi* api = nccl_api();
  RS_REQUIRE(api != nullptr, RS_ERR_UNSUPPORTED, "rs_comm_init: libnccl.so.2 not found (set RS_NCCL_LIB)");
  NcclUniqueId id;
  memcpy(&id, unique_id, sizeof(id));
  Comm* c = new Comm{nullptr, rank, world};
  const int rc = api->comm_init_rank(&c->comm, world, id, rank);
  if (rc != 0) { delete c; return nccl_fail(api, "ncclCommInitRank", rc); }
  *comm_out = c;
  return RS_OK;
}

extern "C" int rs_allreduce_sum(void* comm, float* buf_d, int64_t n, void* stream) {
  RS_REQUIRE(comm && buf_d && n >= 0, RS_ERR_INVALID, "rs_allreduce_sum: bad argument");
  Comm* c = (Comm*)comm;
  if (n == 0 || c->world == 1) return RS_OK;
  NcclApi* api = nccl_api();
  const int rc = api->all_reduce(buf_d, buf_d, (size_t)n, kNcclFloat32, kNcclSum, c->comm, (cudaStream_t)stream);
  if (rc != 0) return nccl_fail(api, "ncclAllReduce", rc);
  return RS_OK;
}

extern "C" void rs_comm_destroy(void* comm) {
  if (!comm) return;
  Comm* c = (Comm*)comm;
  NcclApi* api = nccl_api();
  if (api && c->comm) api->comm_destroy(c->comm);
  delete c;
}
