// rt_comm.cuh — the per-frame exchange of finished tiles INSIDE the C-ABI ("context owns all device memory
// on all GPUs"; the reference host issues one dispatch per frame, RayComputeManager.cs:84-95, and must not have to hand-roll NCCL).
//
// NCCL is bound at run time (dlopen), not at link time: a process that already carries an NCCL (torch bundles its own
// libnccl.so.2) keeps exactly that one, a plain C / C# host gets the system's libnccl.so.2, and librt_b200.so itself has no NCCL
// dependency — single-GPU hosts never load it.  Only the handful of entry points the exchange needs is declared here; the types
// below are NCCL's public ABI (ncclUniqueId = 128 opaque bytes, ncclComm_t = opaque pointer, ncclResult_t = int, ncclChar = 0).
#pragma once
#include <string>
#ifndef RT_SIMT_EMU
#include <dlfcn.h>
#include <cstdlib>
#endif

namespace rtd {

struct NcclApi
{
    typedef struct { char internal[128]; } UniqueId;
    typedef void* Comm;
    enum { Success = 0, Char = 0 };

    int (*GetVersion)(int*) = nullptr;
    int (*GetUniqueId)(UniqueId*) = nullptr;
    int (*CommInitRank)(Comm*, int, UniqueId, int) = nullptr;
    int (*CommDestroy)(Comm) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, Comm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    void* lib = nullptr;
    std::string path;

    // nullptr + `why` when no NCCL can be loaded.  RT_B200_NCCL_LIB names a specific library; otherwise the NCCL already in the
    // process (RTLD_NOLOAD) and then the loader's libnccl.so.2.
    static NcclApi* get(std::string& why)
    {
#ifdef RT_SIMT_EMU
        why = "this build has no NCCL (SIMT interpreter)"; return nullptr;
#else
        static NcclApi api; static bool tried = false; static std::string err;
        if (tried) { if (!api.lib) why = err; return api.lib ? &api : nullptr; }
        tried = true;
        const char* env = getenv("RT_B200_NCCL_LIB");
        if (env && *env) { api.lib = dlopen(env, RTLD_NOW | RTLD_LOCAL); api.path = env; }
        if (!api.lib) { api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD); api.path = "libnccl.so.2 (already loaded in this process)"; }
        if (!api.lib) { api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL); api.path = "libnccl.so.2"; }
        if (!api.lib) { const char* d = dlerror(); err = std::string("cannot load libnccl.so.2: ") + (d ? d : "?"); why = err; return nullptr; }
        struct { const char* name; void** fn; } syms[] = {
            {"ncclGetVersion", (void**)&api.GetVersion}, {"ncclGetUniqueId", (void**)&api.GetUniqueId}, {"ncclCommInitRank", (void**)&api.CommInitRank},
            {"ncclCommDestroy", (void**)&api.CommDestroy}, {"ncclAllGather", (void**)&api.AllGather}, {"ncclGroupStart", (void**)&api.GroupStart},
            {"ncclGroupEnd", (void**)&api.GroupEnd}, {"ncclGetErrorString", (void**)&api.GetErrorString}};
        for (auto& s : syms)
        {
            *s.fn = dlsym(api.lib, s.name);
            if (!*s.fn) { err = std::string("libnccl.so.2 lacks ") + s.name; why = err; api.lib = nullptr; return nullptr; }
        }
        return &api;
#endif
    }
};

} // namespace rtd
