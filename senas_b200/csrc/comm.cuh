// comm.cuh -- gradient exchange of the data-parallel search step through the C ABI (SURVEY.md section 8b item 7:
// comm_init / allreduce_bucket / comm_destroy).  One process per GPU; the communicator is NCCL's (NVLink 5 / NVSwitch),
// bound at run time with dlopen so that libsenas_b200.so has no link-time dependency on it (a single-GPU user never
// needs libnccl).  Because the communicator is owned here -- no process-group watchdog thread polling it -- its
// all-reduce can be CAPTURED into the CUDA graph of the search step (NCCL supports stream capture), which is what puts
// both gradient exchanges inside the replayed step (senas_b200/graphs.py).
#pragma once
#ifndef SENAS_EMU
#include <dlfcn.h>

#include <mutex>

namespace senas_comm {
struct UniqueId {
  char internal[128];
};
typedef void *Comm;
typedef int Result;  // ncclResult_t, 0 = success
enum { kFloat32 = 7, kSum = 0 };  // ncclFloat32, ncclSum (nccl.h)
struct Api {
  Result (*GetUniqueId)(UniqueId *) = nullptr;
  Result (*CommInitRank)(Comm *, int, UniqueId, int) = nullptr;
  Result (*AllReduce)(const void *, void *, size_t, int, int, Comm, cudaStream_t) = nullptr;
  Result (*CommDestroy)(Comm) = nullptr;
  const char *(*GetErrorString)(Result) = nullptr;
  bool ok = false;
  std::string why;
};
static Api &api() {
  static Api a;
  static std::once_flag once;
  std::call_once(once, [] {
    void *h = nullptr;
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
      h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (h) break;
    }
    if (!h) {
      a.why = "libnccl.so.2 not found (dlopen)";
      return;
    }
    a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))dlsym(h, "ncclCommInitRank");
    a.AllReduce = (decltype(a.AllReduce))dlsym(h, "ncclAllReduce");
    a.CommDestroy = (decltype(a.CommDestroy))dlsym(h, "ncclCommDestroy");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(h, "ncclGetErrorString");
    a.ok = a.GetUniqueId && a.CommInitRank && a.AllReduce && a.CommDestroy;
    if (!a.ok) a.why = "libnccl does not export the expected symbols";
  });
  return a;
}
}  // namespace senas_comm
#endif
