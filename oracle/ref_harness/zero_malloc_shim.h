/* Force-included (-include) in front of the UNMODIFIED reference sources.
 * The reference never initialises its cudaMalloc'ed buffers, yet reads bin N/2 of several of
 * them (conv.cu:162-179,233).  To make the compiled reference a
 * deterministic oracle every cudaMalloc is followed by a cudaMemset(0).  It also supplies the
 * headers GCC 13 no longer pulls in transitively (log.h:17, settings.h:17, midi.cu:45). */
#pragma once
#include <cstdint>
#include <cmath>
#include <unistd.h>
#include <cuda_runtime.h>
template <class T>
static inline cudaError_t oracle_zero_malloc(T **p, size_t n)
{
    cudaError_t rc = cudaMalloc((void **)p, n);
    if (rc == cudaSuccess) rc = cudaMemset(*p, 0, n);
    return rc;
}
#define cudaMalloc(p, n) oracle_zero_malloc(p, n)
