// NCCL binding for the one collective of the path: the all-gather of the per-rank pose blocks
// (the reference is single-process, test_kitti_pose.py:133-153 composes on one host).
//
// NCCL is bound at run time (dlopen of libnccl.so.2) instead of at link time: a host that already
// carries an NCCL (torch ships its own) must end up with ONE copy in the process, and dlopen by
// soname returns the copy that is already mapped.  Only the five entry points used are declared;
// their signatures and the two enum values are NCCL's public, ABI-stable ones (nccl.h).
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdlib>
#include <mutex>
#include <string>

namespace davo_comm {

struct UniqueId { char internal[128]; };          // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128)
typedef void* Comm;                               // ncclComm_t
constexpr int kNcclFloat = 7;                     // ncclFloat32
constexpr int kNcclSuccess = 0;

struct Api {
  void* lib = nullptr;
  int (*GetUniqueId)(UniqueId*) = nullptr;
  int (*CommInitRank)(Comm*, int, UniqueId, int) = nullptr;
  int (*CommDestroy)(Comm) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, Comm, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
  std::string why;                                // non-empty: NCCL could not be bound
};

inline const Api& api() {
  static Api a;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[3] = {std::getenv("DAVO_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (!n || !*n) continue;
      a.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (a.lib) break;
    }
    if (!a.lib) {
      const char* e = dlerror();
      a.why = std::string("libnccl.so.2 not found (") + (e ? e : "?") + "); set DAVO_B200_NCCL_LIB";
      return;
    }
    auto sym = [&](const char* s) -> void* {
      void* p = dlsym(a.lib, s);
      if (!p && a.why.empty()) a.why = std::string("NCCL symbol missing: ") + s;
      return p;
    };
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
    a.AllGather = reinterpret_cast<decltype(a.AllGather)>(sym("ncclAllGather"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
    a.GetVersion = reinterpret_cast<decltype(a.GetVersion)>(sym("ncclGetVersion"));
  });
  return a;
}

}  // namespace davo_comm
