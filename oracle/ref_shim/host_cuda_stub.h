/* TEST INFRASTRUCTURE — host-only variant of the reference build (libptref_host.so).
 * There is no GPU in the build container, so the reference's managed-memory allocations
 * (e.g. Core/triangle_mesh.cpp:575-579, scene_parser.cpp:484-504) are redirected to the host
 * heap. This lets the reference's OWN scene pipeline (JSON -> OBJ -> transforms -> material
 * binding) run here to produce golden fixtures. Kernels are never launched in this variant.
 * Pre-included AFTER the thrust headers so thrust itself still sees the real runtime API. */
#pragma once
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>
static inline cudaError_t ptb_stub_malloc_managed(void** p, size_t n) { *p = calloc(1, n ? n : 1); return cudaSuccess; }
static inline cudaError_t ptb_stub_memcpy(void* d, const void* s, size_t n) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t ptb_stub_free(void* p) { free(p); return cudaSuccess; }
#define cudaMallocManaged(p, n) ptb_stub_malloc_managed((void**)(p), (n))
#define cudaMemcpy(d, s, n, k) ptb_stub_memcpy((d), (s), (n))
#define cudaFree(p) ptb_stub_free((void*)(p))
#define cudaDeviceSynchronize() cudaSuccess
#define cudaMalloc(p, n) ptb_stub_malloc_managed((void**)(p), (n))
#define PTB_REF_HOST_ONLY 1
