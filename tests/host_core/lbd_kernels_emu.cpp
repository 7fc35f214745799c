/* tests/host_core/lbd_kernels_emu.cpp -- the SOURCE of the two descriptor / matcher kernels (cube_slam_b200/csrc/cs_lbd_kernels.cuh)
 * compiled for the host against a minimal emulation of the CUDA execution model, and run launch by launch:
 *   one std::thread per CUDA thread of a block, blocks one after the other;
 *   __syncthreads()      -> a std::barrier over the block's threads;
 *   __shared__           -> a function-local static (one block runs at a time, so its threads share it exactly as a CTA shares smem);
 *   __shfl_xor_sync      -> exchange through a block-wide array between two barriers (every thread of the block executes the shuffles of
 *                           k_lbd_match together, so a block barrier is a valid stand-in for warp lock step);
 *   threadIdx / blockIdx -> thread_local structs set by the launcher.
 * tests/test_lbd_host_core.py compares what the launches write with the oracle.  This checks what the arithmetic-only harness
 * (lbd_core_host.cpp) cannot: the kernels' own index arithmetic, phase split and barrier placement.  Test infrastructure, never shipped.
 * g++ -std=c++20 -O2 -ffp-contract=off -pthread. */
#include <barrier>
#include <cstdint>
#include <cstring>
#include <memory>
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

static inline void __syncthreads() { g_block_barrier->arrive_and_wait(); }
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

#include "../../cube_slam_b200/csrc/cs_lbd_core.h"
namespace {
#include "../../cube_slam_b200/csrc/cs_lbd_kernels.cuh"
}

template <typename F> static void launch(unsigned grid, unsigned block, F &&kernel)
{
    for (unsigned b = 0; b < grid; b++) {
        std::barrier<> bar((std::ptrdiff_t)block);
        g_block_barrier = &bar;
        std::vector<std::thread> th;
        for (unsigned t = 0; t < block; t++)
            th.emplace_back([&, t, b] {
                threadIdx.x = t;
                blockIdx.x = b;
                kernel();
                bar.arrive_and_drop(); /* a thread that is done must not hold up the barriers the others still reach */
            });
        for (auto &x : th) x.join();
    }
}

/* k_lbd_describe<<<n_lines, 64>>> */
extern "C" void emu_lbd_describe(const void *lines, int n_lines, const int16_t *dx_all, const int16_t *dy_all, int w, int h, const float *coef, uint8_t *desc,
                                 float *fdesc)
{
    launch((unsigned)n_lines, 64, [&] { k_lbd_describe((const CsLbdLine *)lines, n_lines, dx_all, dy_all, w, h, coef, desc, fdesc); });
}

/* k_lbd_match<<<n_queries, 128>>>; q_all / t_all must be 16-byte aligned like device memory */
extern "C" void emu_lbd_match(const void *q_all, const void *t_all, const int32_t *pair_of_query, const int32_t *t_off, int n_queries, unsigned long long *keys)
{
    launch((unsigned)n_queries, 128, [&] { k_lbd_match((const uint4 *)q_all, (const uint4 *)t_all, pair_of_query, t_off, n_queries, keys); });
}
