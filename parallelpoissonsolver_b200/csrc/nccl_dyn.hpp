// nccl_dyn.hpp -- NCCL entry points resolved at run time (dlopen), only when world_size > 1.
//
// Why not link -lnccl: inside a Python process torch brings its own, newer libnccl.so.2; a DT_NEEDED on the
// system copy would make whichever of the two is loaded first win for both and break the other.  Resolution
// order: a libnccl.so.2 that is already mapped (torch's), then $PPS_NCCL_LIBRARY, then the system library.
#pragma once

#include <dlfcn.h>
#include <nccl.h>   // types only

#include <cstdlib>
#include <mutex>
#include <stdexcept>
#include <string>

namespace pps {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t*, ncclConfig_t*) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    void* handle = nullptr;
};

inline NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    static std::string error;
    std::call_once(once, []() {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_LOCAL);
        if (!h) {
            if (const char* p = std::getenv("PPS_NCCL_LIBRARY")) h = dlopen(p, RTLD_NOW | RTLD_LOCAL);
        }
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
        if (!h) {
            error = std::string("cannot load NCCL: ") + dlerror();
            return;
        }
        api.handle = h;
        auto sym = [&](const char* name) -> void* {
            void* s = dlsym(h, name);
            if (!s && error.empty()) error = std::string("NCCL symbol missing: ") + name;
            return s;
        };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.CommSplit = reinterpret_cast<decltype(api.CommSplit)>(sym("ncclCommSplit"));
        api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
        api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
        api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
        api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
        api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
        api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
        api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
    });
    if (!error.empty()) throw std::runtime_error(error);
    return api;
}

}  // namespace pps
