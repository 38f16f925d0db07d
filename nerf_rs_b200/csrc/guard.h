// guard.h -- device allocations with guard bands: the out-of-bounds / uninitialised-read check of last resort.
// compute-sanitizer is closed on the pool this code is developed on, so the library carries its own: with NERF_B200_GUARD=1 in
// the environment every device buffer the context (and the MLP state) allocates gets a 4 KB band of a known pattern on both
// sides and is itself filled with 0xFF bytes (NaNs: anything read before it is written poisons the outputs the parity tests
// check for finiteness). nerf_debug_check_guards verifies the bands after the kernels have run. Without the variable these are
// plain cudaMalloc / cudaFree.
#pragma once
#include <cuda_runtime.h>

cudaError_t guard_malloc(void **p, size_t bytes);
template <typename T>
inline cudaError_t guard_malloc(T **p, size_t bytes) { return guard_malloc(reinterpret_cast<void **>(p), bytes); }
cudaError_t guard_free(void *p);
// number of allocations whose guard bands were overwritten (0 = clean), -1 when the guard mode is off; synchronises the device
int guard_check(int *n_allocations);
