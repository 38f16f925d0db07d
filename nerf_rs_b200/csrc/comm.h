// comm.h -- NCCL bound at run time (dlopen), so libnerf_b200.so has no link-time
// dependency on a particular libnccl: inside a torchrun process the torch-bundled
// libnccl.so.2 is already mapped and is the one resolved.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

struct NcclApi;
struct CommState {
    NcclApi *api = nullptr;
    void *comm = nullptr;
    int rank = 0, nranks = 1;
    // peer-memory path (CUDA IPC over NVLink): every rank's two gradient accumulation buffers and flag array
    bool p2p = false;
    float *peer_grads[2][8] = {};
    unsigned int *peer_flags[8] = {};
    unsigned int *my_flags = nullptr;
    unsigned int p2p_step = 0;
};

int comm_unique_id(void *id128, char *err, size_t errlen);
int comm_init_rank(CommState &cs, const void *id128, int rank, int nranks, char *err, size_t errlen);
int comm_allreduce_sum_f32(CommState &cs, float *buf, int64_t n, cudaStream_t st, char *err, size_t errlen);
int comm_allgather_bytes(CommState &cs, const void *send, void *recv, int64_t bytes_per_rank, cudaStream_t st, char *err,
                         size_t errlen);
// Exchange CUDA IPC handles of (grads0, grads1) and a freshly allocated flag array over the NCCL communicator and map every
// peer's buffers. Returns 0 and sets cs.p2p on success; on failure cs.p2p stays false (the NCCL all-reduce path is used).
int comm_p2p_setup(CommState &cs, float *grads0, float *grads1, cudaStream_t st, char *err, size_t errlen);
void comm_destroy(CommState &cs);
