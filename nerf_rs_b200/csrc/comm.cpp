// comm.cpp -- data-parallel plumbing: NCCL all-reduce of the flat gradient blob.
// The reference has no multi-device code; this is the one exchange step the
// data-parallel path needs.
#include "comm.h"

#include <dlfcn.h>

#include <cstdio>
#include <cstring>

typedef struct { char internal[128]; } ncclUniqueId_t;
typedef int ncclResult_t_;
typedef void *ncclComm_t_;

struct NcclApi {
    void *handle;
    ncclResult_t_ (*GetUniqueId)(ncclUniqueId_t *);
    ncclResult_t_ (*CommInitRank)(ncclComm_t_ *, int, ncclUniqueId_t, int);
    ncclResult_t_ (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t_, cudaStream_t);
    ncclResult_t_ (*AllGather)(const void *, void *, size_t, int, ncclComm_t_, cudaStream_t);
    ncclResult_t_ (*CommDestroy)(ncclComm_t_);
    const char *(*GetErrorString)(ncclResult_t_);
};

static NcclApi *load_api(char *err, size_t errlen) {
    static NcclApi api;
    static bool loaded = false;
    if (loaded) return &api;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *n : names) {
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) {
        snprintf(err, errlen, "dlopen(libnccl.so.2) failed: %s", dlerror());
        return nullptr;
    }
    api.handle = h;
#define LOAD(field, sym)                                             \
    *(void **)(&api.field) = dlsym(h, sym);                          \
    if (!api.field) {                                                \
        snprintf(err, errlen, "dlsym(%s) failed", sym);              \
        return nullptr;                                              \
    }
    LOAD(GetUniqueId, "ncclGetUniqueId")
    LOAD(CommInitRank, "ncclCommInitRank")
    LOAD(AllReduce, "ncclAllReduce")
    LOAD(AllGather, "ncclAllGather")
    LOAD(CommDestroy, "ncclCommDestroy")
    LOAD(GetErrorString, "ncclGetErrorString")
#undef LOAD
    loaded = true;
    return &api;
}

int comm_unique_id(void *id128, char *err, size_t errlen) {
    NcclApi *api = load_api(err, errlen);
    if (!api) return -1;
    ncclUniqueId_t id;
    ncclResult_t_ r = api->GetUniqueId(&id);
    if (r != 0) {
        snprintf(err, errlen, "ncclGetUniqueId: %s", api->GetErrorString(r));
        return -1;
    }
    memcpy(id128, &id, 128);
    return 0;
}

int comm_init_rank(CommState &cs, const void *id128, int rank, int nranks, char *err, size_t errlen) {
    NcclApi *api = load_api(err, errlen);
    if (!api) return -1;
    ncclUniqueId_t id;
    memcpy(&id, id128, 128);
    ncclComm_t_ c = nullptr;
    ncclResult_t_ r = api->CommInitRank(&c, nranks, id, rank);
    if (r != 0) {
        snprintf(err, errlen, "ncclCommInitRank: %s", api->GetErrorString(r));
        return -1;
    }
    cs.api = api;
    cs.comm = c;
    cs.rank = rank;
    cs.nranks = nranks;
    return 0;
}

int comm_allreduce_sum_f32(CommState &cs, float *buf, int64_t n, cudaStream_t st, char *err, size_t errlen) {
    if (!cs.comm) return 0;
    ncclResult_t_ r = cs.api->AllReduce(buf, buf, (size_t)n, /*ncclFloat32*/ 7, /*ncclSum*/ 0, cs.comm, st);
    if (r != 0) {
        snprintf(err, errlen, "ncclAllReduce: %s", cs.api->GetErrorString(r));
        return -1;
    }
    return 0;
}

int comm_allgather_bytes(CommState &cs, const void *send, void *recv, int64_t bytes_per_rank, cudaStream_t st, char *err,
                         size_t errlen) {
    if (!cs.comm) return 0;
    ncclResult_t_ r = cs.api->AllGather(send, recv, (size_t)bytes_per_rank, /*ncclInt8*/ 0, cs.comm, st);
    if (r != 0) {
        snprintf(err, errlen, "ncclAllGather: %s", cs.api->GetErrorString(r));
        return -1;
    }
    return 0;
}

void comm_destroy(CommState &cs) {
    if (cs.comm && cs.api) cs.api->CommDestroy(cs.comm);
    cs.comm = nullptr;
    cs.nranks = 1;
    cs.rank = 0;
}
