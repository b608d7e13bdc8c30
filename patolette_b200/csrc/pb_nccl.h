// pb_nccl.h - the library's run-time-resolved NCCL communicator (pb_nccl.cpp).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

// 128-byte ncclUniqueId of a new communicator (call on one rank, hand the bytes to the others); 0 on success
int pb_nccl_unique_id(char out[128]);
// collective over all ranks: (re)creates the communicator on the current CUDA device; world == 1 just drops it
int pb_nccl_init(int rank, int world, const char id_bytes[128]);
void pb_nccl_destroy();
bool pb_nccl_active();
int pb_nccl_rank();
int pb_nccl_world();
int pb_nccl_version(); // e.g. 22809, 0 when NCCL cannot be loaded
// in-place capable all-gather of `bytes` bytes per rank on stream st; throws pb_cuda_error on failure
void pb_nccl_allgather(const void *send, void *recv, size_t bytes, cudaStream_t st);
