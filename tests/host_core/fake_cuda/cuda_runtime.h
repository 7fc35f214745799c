/* tests/host_core/fake_cuda/cuda_runtime.h -- stands in for <cuda_runtime.h> when a .cu file of the library is compiled for the host by
 * the CPU test suite (tests/host_core/context_emu.cpp): the emulation of cuda_emu.h plus inert versions of the stream / event / attribute
 * calls the host orchestration makes.  "Device" memory is host memory; everything is synchronous.  Never shipped. */
#ifndef CS_TEST_FAKE_CUDA_RUNTIME_H
#define CS_TEST_FAKE_CUDA_RUNTIME_H
#include "../cuda_emu.h"

typedef void *cudaEvent_t;
struct int2 {
    int x, y;
};
struct int4 {
    int x, y, z, w;
};
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
enum cudaDeviceAttr { cudaDevAttrMaxSharedMemoryPerBlockOptin = 97 };
struct cudaFuncAttributes {
    size_t sharedSizeBytes = 0;
};
static inline cudaError_t cudaGetDeviceCount(int *n)
{
    *n = 1;
    return 0;
}
static inline cudaError_t cudaGetDevice(int *d)
{
    *d = 0;
    return 0;
}
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned)
{
    *s = (cudaStream_t)1;
    return 0;
}
static inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned, int)
{
    *s = (cudaStream_t)1;
    return 0;
}
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
static inline cudaError_t cudaDeviceGetStreamPriorityRange(int *lo, int *hi)
{
    *lo = 0;
    *hi = -1;
    return 0;
}
static inline cudaError_t cudaEventCreate(cudaEvent_t *e)
{
    *e = (cudaEvent_t)1;
    return 0;
}
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned)
{
    *e = (cudaEvent_t)1;
    return 0;
}
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return 0; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t)
{
    *ms = 0;
    return 0;
}
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return 0; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind)
{
    std::memcpy(d, s, n);
    return 0;
}
static inline cudaError_t cudaMemcpy2D(void *d, size_t dpitch, const void *s, size_t spitch, size_t width, size_t height, cudaMemcpyKind)
{
    for (size_t y = 0; y < height; y++) std::memcpy((char *)d + y * dpitch, (const char *)s + y * spitch, width);
    return 0;
}
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t)
{
    std::memset(d, v, n);
    return 0;
}
static inline cudaError_t cudaFreeHost(void *p)
{
    std::free(p);
    return 0;
}
template <typename K> static inline cudaError_t cudaFuncGetAttributes(cudaFuncAttributes *, K) { return 1; }
static inline cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr, int)
{
    *v = 0;
    return 0;
}
template <typename K> static inline cudaError_t cudaFuncSetAttribute(K, cudaFuncAttribute, int) { return 0; }
#endif
