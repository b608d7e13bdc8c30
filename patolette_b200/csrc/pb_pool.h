// pb_pool.h - caching device allocator.
//
// One quantisation allocates ~100 bytes per pixel of working buffers.  cudaMalloc / cudaFree
// cost ~100-200 ms per call at 4096^2 (cudaFree also synchronises the device), which is as much
// as the whole pipeline takes, so freed blocks are kept per device and handed out again to the
// next call asking for (almost) the same size.  patolette_b200_release_cache() returns them.
#pragma once
#include <stddef.h>

void *pb_pool_alloc(size_t bytes); // throws pb_cuda_error on failure (after flushing the cache once)
void pb_pool_free(void *p);
void pb_pool_release_all();
size_t pb_pool_cached_bytes();
