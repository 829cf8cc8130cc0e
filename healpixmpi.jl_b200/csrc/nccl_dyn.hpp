// nccl_dyn.hpp — NCCL bound at run time (dlopen), so that the single-GPU path has no NCCL
// dependency and a host process that already loaded an NCCL (e.g. PyTorch's bundled one) shares it.
// Replaces MPI.Alltoallv! of the reference (src/sht.jl:32,133) as the transport of the leg exchange.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <string>

#include "common.hpp"

namespace hpsht {

struct NcclApi {
  void* handle = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclSend) Send = nullptr;
  decltype(&ncclRecv) Recv = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclBroadcast) Broadcast = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  decltype(&ncclGetVersion) GetVersion = nullptr;

  static NcclApi& get() {
    static NcclApi api;
    if (!api.ready) api.load();
    return api;
  }
  bool ready = false;   // every symbol resolved (a partial load leaves it false: the next get() tries again)

 private:
  void load() {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (handle) break;
    }
    if (!handle)
      throw Error(HPSHT_ERR_NCCL, std::string("cannot dlopen libnccl.so.2: ") + dlerror());
#define HP_SYM(field, name)                                                          \
  field = reinterpret_cast<decltype(field)>(dlsym(handle, name));                    \
  if (!field) throw Error(HPSHT_ERR_NCCL, std::string("libnccl lacks symbol ") + name);
    HP_SYM(GetUniqueId, "ncclGetUniqueId")
    HP_SYM(CommInitRank, "ncclCommInitRank")
    HP_SYM(CommDestroy, "ncclCommDestroy")
    HP_SYM(Send, "ncclSend")
    HP_SYM(Recv, "ncclRecv")
    HP_SYM(AllReduce, "ncclAllReduce")
    HP_SYM(Broadcast, "ncclBroadcast")
    HP_SYM(GroupStart, "ncclGroupStart")
    HP_SYM(GroupEnd, "ncclGroupEnd")
    HP_SYM(GetErrorString, "ncclGetErrorString")
    HP_SYM(GetVersion, "ncclGetVersion")
#undef HP_SYM
    ready = true;
  }
};

#define HP_NCCL(call)                                                                         \
  do {                                                                                        \
    ncclResult_t r__ = (call);                                                                \
    if (r__ != ncclSuccess)                                                                   \
      throw ::hpsht::Error(HPSHT_ERR_NCCL, std::string(#call) + ": " +                        \
                                               ::hpsht::NcclApi::get().GetErrorString(r__));  \
  } while (0)

}  // namespace hpsht
