/* cs_debug_atan2: the device side of cs_pmath.h, exposed so that a test can check it against the host evaluation bit for bit. */
#include <cuda_runtime.h>
#include <stdint.h>

#include "cs_internal.h"
#include "cs_pmath.h"

namespace {
__global__ void k_debug_atan2(const double *__restrict__ y, const double *__restrict__ x, double *__restrict__ out, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = cs_pm_atan2(y[i], x[i]);
}
}  // namespace

extern "C" int cs_debug_atan2(cs_ctx *c, const double *y, const double *x, double *out, int n)
{
    if (!c || !y || !x || !out || n < 0) return CS_ERR_INVALID_ARG;
    if (n == 0) return CS_OK;
    cudaSetDevice(cs_ctx_device(c));
    double *d = nullptr;
    if (cudaMalloc(&d, (size_t)n * 24) != cudaSuccess) return cs_ctx_fail(c, CS_ERR_CUDA, "cudaMalloc failed");
    cudaMemcpy(d, y, (size_t)n * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(d + n, x, (size_t)n * 8, cudaMemcpyHostToDevice);
    k_debug_atan2<<<(n + 255) / 256, 256>>>(d, d + n, d + 2 * (size_t)n, n);
    cudaMemcpy(out, d + 2 * (size_t)n, (size_t)n * 8, cudaMemcpyDeviceToHost);
    const cudaError_t e = cudaGetLastError();
    cudaFree(d);
    return e == cudaSuccess ? CS_OK : cs_ctx_fail(c, CS_ERR_CUDA, "cs_debug_atan2 failed: %s", cudaGetErrorString(e));
}

/* the host evaluation of the same header (what a C++ caller of the library would get) */
extern "C" double cs_atan2_host(double y, double x) { return cs_pm_atan2(y, x); }
