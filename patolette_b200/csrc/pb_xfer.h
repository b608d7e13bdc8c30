// pb_xfer.h - host <-> device copies for caller-owned (usually pageable) buffers.
//
// patolette()'s contract is plain host pointers.  cudaMemcpy from pageable memory is staged by the
// driver through one thread (~10 GB/s: 74 ms of copies around 77 ms of compute at 4096^2).  These
// helpers stripe the buffer over a few host threads, each with its own pinned bounce buffers and
// stream, which reaches the PCIe rate.  Pinned / registered buffers go straight to cudaMemcpyAsync.
// Both calls return when the copy is complete.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

void pb_copy_h2d(void *dst_dev, const void *src_host, size_t bytes, cudaStream_t st);
void pb_copy_d2h(void *dst_host, const void *src_dev, size_t bytes, cudaStream_t st);
// true for pinned / registered / managed host memory (DMA straight from the caller's buffer)
bool pb_host_is_pinned(const void *p);
