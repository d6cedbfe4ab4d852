// Row e / a11 of SURVEY.md section 8: the data-parallel step's ONE collective, `ncclAllReduce(SUM, fp32)` over the flat
// gradient buffer, behind the C ABI (section 8b: nabu_allreduce_grads "wraps ncclAllReduce on the same stream").
// Replaces the parameter-server gradient path of trainers/trainer.py:479-510, 556-569 (synchronous instead of async).
//
// NCCL is resolved at run time with dlopen (the process that loads this library has torch's bundled libnccl.so.2 mapped
// already; a plain C host gets the system one), so the library has no link-time dependency on it and every entry point
// that does not communicate works without NCCL installed.  One communicator per process (one process per GPU).
#include "common.cuh"
#include "nabu_b200.h"
#include <dlfcn.h>
#include <string.h>

namespace nabu {
namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;                       // ncclSuccess = 0
enum { NCCL_FLOAT32 = 7, NCCL_SUM = 0 };        // ncclDataType_t / ncclRedOp_t values of nccl.h (stable since NCCL 2.0)

struct Nccl {
  void* handle;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  const char* (*GetErrorString)(ncclResult_t);
  ncclComm_t comm;
  int rank, world;
};

Nccl& nccl() {
  static Nccl n = {};
  return n;
}

int load_nccl() {
  Nccl& n = nccl();
  if (n.handle) return 0;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    n.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (n.handle) break;
  }
  NABU_REQUIRE(n.handle != nullptr, "nabu_comm: libnccl.so.2 not found (%s)", dlerror());
  n.GetUniqueId = (ncclResult_t(*)(ncclUniqueId*))dlsym(n.handle, "ncclGetUniqueId");
  n.CommInitRank = (ncclResult_t(*)(ncclComm_t*, int, ncclUniqueId, int))dlsym(n.handle, "ncclCommInitRank");
  n.AllReduce = (ncclResult_t(*)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(n.handle, "ncclAllReduce");
  n.CommDestroy = (ncclResult_t(*)(ncclComm_t))dlsym(n.handle, "ncclCommDestroy");
  n.GetErrorString = (const char* (*)(ncclResult_t))dlsym(n.handle, "ncclGetErrorString");
  NABU_REQUIRE(n.GetUniqueId && n.CommInitRank && n.AllReduce && n.CommDestroy, "nabu_comm: NCCL symbols missing");
  return 0;
}

#define NABU_CHECK_NCCL(expr)                                                                            \
  do {                                                                                                   \
    ncclResult_t _r = (expr);                                                                            \
    if (_r != 0) {                                                                                       \
      ::nabu::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,                               \
                        nccl().GetErrorString ? nccl().GetErrorString(_r) : "nccl error");               \
      return 1;                                                                                          \
    }                                                                                                    \
  } while (0)

}  // namespace
}  // namespace nabu

using namespace nabu;

extern "C" int nabu_comm_unique_id(void* id128) {
  if (int e = load_nccl()) return e;
  ncclUniqueId id;
  NABU_CHECK_NCCL(nccl().GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return 0;
}

extern "C" int nabu_comm_init(const void* id128, int rank, int world) {
  if (int e = load_nccl()) return e;
  Nccl& n = nccl();
  NABU_REQUIRE(world >= 1 && rank >= 0 && rank < world, "nabu_comm_init: rank %d of %d", rank, world);
  if (n.comm) {
    NABU_CHECK_NCCL(n.CommDestroy(n.comm));
    n.comm = nullptr;
  }
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  NABU_CHECK_NCCL(n.CommInitRank(&n.comm, world, id, rank));
  n.rank = rank; n.world = world;
  return 0;
}

extern "C" int nabu_comm_world(void) { return nccl().comm ? nccl().world : 1; }

extern "C" int nabu_comm_destroy(void) {
  Nccl& n = nccl();
  if (n.comm) {
    NABU_CHECK_NCCL(n.CommDestroy(n.comm));
    n.comm = nullptr;
  }
  return 0;
}

extern "C" int nabu_allreduce_grads(float* grads, size_t n, void* stream) {
  Nccl& c = nccl();
  if (!c.comm) return 0;                       // a single process: the sum over ranks is the buffer itself
  NABU_CHECK_NCCL(c.AllReduce(grads, grads, n, NCCL_FLOAT32, NCCL_SUM, c.comm, (cudaStream_t)stream));
  return 0;
}
