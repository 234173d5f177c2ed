// pb200_comm.cuh - NCCL halo exchange of a row-stripped raster behind the C ABI (BASELINE configs[4]).
//
// The only non-local function of the path is the one-row stencil of np.gradient inside _compute_opera_shadow_layer
// (D:4255): a rank that owns rows [r0, r1) of an oversized raster needs DEM rows r0 - 1 and r1 from its neighbours.
// pb200_halo_exchange_dem posts both directions as ncclSend / ncclRecv pairs inside one ncclGroupStart / End on the
// caller's stream (NVLink / NVSwitch on a B200 box); pb200_comm_allreduce_u64 sums the coverage counters.
//
// NCCL is bound at run time (dlopen of libnccl.so.2, preferring the copy the host process already holds - e.g. the
// one PyTorch ships) so that the library has no link-time dependency on it and single-GPU users never load it.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>

#include <cstdlib>
#include <mutex>
#include <string>

namespace pb200 {

struct NcclUniqueId { char internal[128]; };          // ncclUniqueId (nccl.h: NCCL_UNIQUE_ID_BYTES = 128)
typedef struct ncclComm *NcclComm;
enum { NCCL_SUCCESS = 0, NCCL_UINT64 = 5, NCCL_FLOAT32 = 7, NCCL_SUM = 0 };   // ncclDataType_t / ncclRedOp_t values

struct NcclApi {
    void *handle = nullptr;
    int (*GetVersion)(int *) = nullptr;
    int (*GetUniqueId)(NcclUniqueId *) = nullptr;
    int (*CommInitRank)(NcclComm *, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    std::string error;
};

inline NcclApi *nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *env = std::getenv("PB200_NCCL_PATH");
        void *h = nullptr;
        if (env && *env) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);     // the copy the process already holds
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) {
            const char *e = dlerror();
            api.error = std::string("libnccl.so.2 not found (set PB200_NCCL_PATH): ") + (e ? e : "");
            return;
        }
        bool ok = true;
        auto sym = [&](const char *name) -> void * {
            void *p = dlsym(h, name);
            if (!p) { ok = false; api.error = std::string("NCCL symbol missing: ") + name; }
            return p;
        };
        api.GetVersion = (int (*)(int *))sym("ncclGetVersion");
        api.GetUniqueId = (int (*)(NcclUniqueId *))sym("ncclGetUniqueId");
        api.CommInitRank = (int (*)(NcclComm *, int, NcclUniqueId, int))sym("ncclCommInitRank");
        api.CommDestroy = (int (*)(NcclComm))sym("ncclCommDestroy");
        api.GroupStart = (int (*)())sym("ncclGroupStart");
        api.GroupEnd = (int (*)())sym("ncclGroupEnd");
        api.Send = (int (*)(const void *, size_t, int, int, NcclComm, cudaStream_t))sym("ncclSend");
        api.Recv = (int (*)(void *, size_t, int, int, NcclComm, cudaStream_t))sym("ncclRecv");
        api.AllReduce = (int (*)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t))sym("ncclAllReduce");
        api.GetErrorString = (const char *(*)(int))sym("ncclGetErrorString");
        if (ok) api.handle = h;
    });
    return &api;
}

struct Comm {
    NcclComm comm = nullptr;
    int rank = 0, nranks = 1;
};

}  // namespace pb200
