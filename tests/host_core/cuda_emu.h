/* tests/host_core/cuda_emu.h -- a minimal emulation of the CUDA execution model and of the few runtime calls cs_lbd.cu makes, for the CPU
 * test suite (never shipped):
 *   one std::thread per CUDA thread of a block, blocks one after the other;
 *   __syncthreads()      -> a std::barrier over the block's threads;
 *   __shared__           -> a function-local static (one block runs at a time, so its threads share it exactly as a CTA shares smem);
 *   __shfl_xor_sync      -> exchange through a block-wide array between two barriers (valid where every thread of the block executes the
 *                           shuffle together, as in k_lbd_match);
 *   threadIdx / blockIdx -> thread_local structs set by the launcher;
 *   cudaMalloc / cudaMemcpyAsync / ... -> the host heap and memcpy ("device" pointers are host pointers). */
#ifndef CS_TEST_CUDA_EMU_H
#define CS_TEST_CUDA_EMU_H
#include <barrier>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

struct EmuDim3 {
    unsigned x = 0, y = 0, z = 0;
};
struct uint4 {
    unsigned x, y, z, w;
};
static thread_local EmuDim3 threadIdx, blockIdx;
static std::barrier<> *g_block_barrier = nullptr;
static unsigned long long g_shfl[1024];

#ifdef CS_EMU_DROP_BARRIER
/* negative control for the racecheck build: every thread skips its CS_EMU_DROP_BARRIER-th __syncthreads (1-based) -- ThreadSanitizer must
 * then report the accesses that barrier was ordering */
static thread_local int g_sync_count = 0;
static inline void __syncthreads()
{
    if (++g_sync_count == CS_EMU_DROP_BARRIER) return;
    g_block_barrier->arrive_and_wait();
}
#else
static inline void __syncthreads() { g_block_barrier->arrive_and_wait(); }
#endif
static inline unsigned long long __shfl_xor_sync(unsigned, unsigned long long v, int lane_mask)
{
    g_shfl[threadIdx.x] = v;
    g_block_barrier->arrive_and_wait();
    const unsigned long long r = g_shfl[(threadIdx.x & ~31u) | ((threadIdx.x ^ (unsigned)lane_mask) & 31u)];
    g_block_barrier->arrive_and_wait();
    return r;
}

#define __global__
#define __shared__ static
#define __restrict__
#define __launch_bounds__(n)

template <typename F> static void emu_launch(unsigned grid, unsigned block, F &&kernel)
{
    for (unsigned b = 0; b < grid; b++) {
        std::barrier<> bar((std::ptrdiff_t)block);
        g_block_barrier = &bar;
        std::vector<std::thread> th;
        for (unsigned t = 0; t < block; t++)
            th.emplace_back([&, t, b] {
                threadIdx.x = t;
                blockIdx.x = b;
#ifdef CS_EMU_DROP_BARRIER
                g_sync_count = 0;
#endif
                kernel();
                bar.arrive_and_drop(); /* a thread that is done must not hold up the barriers the others still reach */
            });
        for (auto &x : th) x.join();
    }
}
#define CS_LBD_EMU_LAUNCH(grid, block, fn) emu_launch((grid), (block), (fn))

/* ---- the runtime calls cs_lbd.cu makes */
typedef int cudaError_t;
typedef void *cudaStream_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
static inline cudaError_t cudaMalloc(void **p, size_t n)
{
    *p = std::aligned_alloc(256, (n + 255) / 256 * 256);
    if (*p) std::memset(*p, 0xA5, n); /* device memory is not zeroed: poison it */
    return *p ? 0 : 2;
}
static inline cudaError_t cudaFree(void *p)
{
    std::free(p);
    return 0;
}
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t)
{
    std::memcpy(d, s, n);
    return 0;
}
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline const char *cudaGetErrorString(cudaError_t) { return "emulated"; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
#endif
