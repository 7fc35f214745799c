/* tests/host_core/cuda_emu_full.h -- a fuller emulation of the CUDA execution model for the CPU test suite (never shipped): 1-D thread
 * blocks in up to 3-D grids, warp-level primitives, atomics, dynamic shared memory, and the runtime calls a .cu file's host code makes.
 *   one std::thread per CUDA thread of a block, blocks one after the other;
 *   __syncthreads()                 a std::barrier over the block's threads;
 *   __shfl*_sync / __ballot_sync /  exchange through per-thread slots between two arrivals at a per-WARP std::barrier (warps of a block run
 *   __syncwarp                      independently of each other, as on the GPU); lanes that have returned no longer take part;
 *   __shared__                      a function-local static (one block runs at a time); `extern __shared__ T name[]` is rewritten by the
 *                                   test's preprocessing step into a pointer to one block-wide byte array;
 *   threadIdx / blockIdx            thread_local; blockDim / gridDim globals set by the launcher;
 *   cudaMalloc / cudaMemcpy...      the host heap ("device" memory is host memory; everything is synchronous).
 * Transcendentals inside kernels come from the host's libm here (the GPU's differ by an ulp or two): a rehearsal of logic and indexing,
 * not of the last bit of a threshold test -- the committed goldens and the GPU tests are for that. */
#ifndef CS_TEST_CUDA_EMU_FULL_H
#define CS_TEST_CUDA_EMU_FULL_H
#include <algorithm>
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint2 {
    unsigned x, y;
};
struct uint4 {
    unsigned x, y, z, w;
};
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static thread_local dim3 threadIdx(0, 0, 0), blockIdx(0, 0, 0);
static dim3 blockDim, gridDim;

struct EmuBlockState {
    std::barrier<> *bar = nullptr;
    std::vector<std::unique_ptr<std::barrier<>>> warp;
    unsigned long long slot[1024];
    std::atomic<int> active[1024];
    alignas(128) unsigned char dyn[232448];
};
static EmuBlockState g_blk;

static inline std::barrier<> &emu_warp_bar() { return *g_blk.warp[threadIdx.x >> 5]; }
static inline void __syncthreads() { g_blk.bar->arrive_and_wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu_warp_bar().arrive_and_wait(); }
template <typename T> static inline T emu_exchange(T v, int src_lane)
{
    static_assert(sizeof(T) <= 8, "shuffle payload");
    unsigned long long raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    g_blk.slot[threadIdx.x] = raw;
    emu_warp_bar().arrive_and_wait();
    const unsigned src = (threadIdx.x & ~31u) | ((unsigned)src_lane & 31u);
    const unsigned long long got = g_blk.slot[src < blockDim.x ? src : threadIdx.x];
    emu_warp_bar().arrive_and_wait();
    T r;
    std::memcpy(&r, &got, sizeof(T));
    return r;
}
template <typename T> static inline T __shfl_sync(unsigned, T v, int src_lane) { return emu_exchange(v, src_lane); }
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int lane_mask) { return emu_exchange(v, (int)((threadIdx.x & 31u) ^ (unsigned)lane_mask)); }
template <typename T> static inline T __shfl_up_sync(unsigned, T v, unsigned delta)
{
    const int lane = (int)(threadIdx.x & 31u);
    return emu_exchange(v, lane >= (int)delta ? lane - (int)delta : lane);
}
static inline unsigned __ballot_sync(unsigned, int pred)
{
    g_blk.slot[threadIdx.x] = pred ? 1ull : 0ull;
    emu_warp_bar().arrive_and_wait();
    unsigned m = 0;
    const unsigned base = threadIdx.x & ~31u;
    for (unsigned l = 0; l < 32 && base + l < blockDim.x; l++)
        if (g_blk.active[base + l].load() && g_blk.slot[base + l]) m |= 1u << l;
    emu_warp_bar().arrive_and_wait();
    return m;
}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
template <typename T> static inline T __ldg(const T *p) { return *p; }
static inline float __int_as_float(int v)
{
    float f;
    std::memcpy(&f, &v, 4);
    return f;
}
static inline int __float_as_int(float f)
{
    int v;
    std::memcpy(&v, &f, 4);
    return v;
}
template <typename T> static inline T atomicOr(T *p, T v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
template <typename T> static inline T atomicAdd(T *p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline size_t __cvta_generic_to_shared(const void *p) { return (size_t)p; }
using std::max;
using std::min;

#define __global__
#define __device__
#define __host__
#define __shared__ static
#define __restrict__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __grid_constant__
#define __align__(n) __attribute__((aligned(n)))

template <typename K> struct EmuLaunch {
    K kernel;
    dim3 grid;
    unsigned block;
    template <typename... A> void operator()(A... args)
    {
        blockDim = dim3(block);
        gridDim = grid;
        for (unsigned bz = 0; bz < grid.z; bz++)
            for (unsigned by = 0; by < grid.y; by++)
                for (unsigned bx = 0; bx < grid.x; bx++) {
                    std::barrier<> bar((std::ptrdiff_t)block);
                    g_blk.bar = &bar;
                    g_blk.warp.clear();
                    for (unsigned w0 = 0; w0 < block; w0 += 32) g_blk.warp.emplace_back(new std::barrier<>((std::ptrdiff_t)std::min(32u, block - w0)));
                    for (unsigned t = 0; t < block; t++) g_blk.active[t].store(1);
                    std::vector<std::thread> th;
                    for (unsigned t = 0; t < block; t++)
                        th.emplace_back([&, t, bx, by, bz] {
                            threadIdx = dim3(t, 0, 0);
                            blockIdx = dim3(bx, by, bz);
                            kernel(args...);
                            g_blk.active[t].store(0); /* a lane that has returned takes no part in later ballots or barriers */
                            g_blk.warp[t >> 5]->arrive_and_drop();
                            bar.arrive_and_drop();
                        });
                    for (auto &x : th) x.join();
                }
    }
};
#define EMU_LAUNCH(kernel, grid, block) EmuLaunch<decltype(&kernel)>{&kernel, dim3(grid), (unsigned)(block)}

/* ---- runtime calls */
typedef int cudaError_t;
typedef void *cudaStream_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
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
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind)
{
    std::memcpy(d, s, n);
    return 0;
}
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t)
{
    std::memset(d, v, n);
    return 0;
}
static inline cudaError_t cudaMemset(void *d, int v, size_t n)
{
    std::memset(d, v, n);
    return 0;
}
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline const char *cudaGetErrorString(cudaError_t) { return "emulated"; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
template <typename K> static inline cudaError_t cudaFuncSetAttribute(K, cudaFuncAttribute, int) { return 0; }

/* cs_tma.cuh's interface, inert: no tile goes through a copy engine here */
struct CUtensorMap {
    unsigned char opaque[128];
};
static inline bool cs_make_tmap_bytes(CUtensorMap *, const void *, int64_t, int64_t, int64_t, int, int) { return false; }
static inline void cs_mbar_init(unsigned long long *) {}
static inline void cs_tma_load_2d(const CUtensorMap *, void *, unsigned long long *, int, int, uint32_t) {}
static inline bool cs_mbar_wait(unsigned long long *, uint32_t) { return true; }
#endif
