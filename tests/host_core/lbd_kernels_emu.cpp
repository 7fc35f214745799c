/* tests/host_core/lbd_kernels_emu.cpp -- the SOURCE of the two descriptor / matcher kernels (cube_slam_b200/csrc/cs_lbd_kernels.cuh)
 * compiled for the host against a minimal emulation of the CUDA execution model (cuda_emu.h: a std::thread per CUDA thread, a std::barrier
 * for __syncthreads, function-local statics for __shared__), and run launch by launch through the same launch wrappers the library uses.
 * tests/test_lbd_host_core.py compares what the launches write with the oracle.  This checks what the arithmetic-only harness
 * (lbd_core_host.cpp) cannot: the kernels' own index arithmetic, phase split and barrier placement.  Test infrastructure, never shipped.
 * g++ -std=c++20 -O2 -ffp-contract=off -pthread. */
#include "cuda_emu.h"

#include "../../cube_slam_b200/csrc/cs_lbd_core.h"
namespace {
#include "../../cube_slam_b200/csrc/cs_lbd_kernels.cuh"
}

/* k_lbd_describe<<<n_lines, 64>>> */
extern "C" void emu_lbd_describe(const void *lines, int n_lines, const int16_t *dx_all, const int16_t *dy_all, int w, int h, const float *coef, uint8_t *desc,
                                 float *fdesc)
{
    launch_lbd_describe((unsigned)n_lines, nullptr, (const CsLbdLine *)lines, n_lines, dx_all, dy_all, w, h, coef, desc, fdesc);
}

/* k_lbd_match<<<n_queries, 128>>>; q_all / t_all must be 16-byte aligned like device memory */
extern "C" void emu_lbd_match(const void *q_all, const void *t_all, const int32_t *pair_of_query, const int32_t *t_off, int n_queries, unsigned long long *keys)
{
    launch_lbd_match((unsigned)n_queries, nullptr, (const uint4 *)q_all, (const uint4 *)t_all, pair_of_query, t_off, n_queries, keys);
}
