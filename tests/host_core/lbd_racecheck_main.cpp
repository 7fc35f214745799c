/* tests/host_core/lbd_racecheck_main.cpp -- a "racecheck" for k_lbd_describe / k_lbd_match without a GPU: the kernels' source under the
 * emulation of cuda_emu.h (a std::thread per CUDA thread, std::barrier for __syncthreads), built with -fsanitize=thread and run on random
 * data.  ThreadSanitizer then reports any pair of accesses to the emulated shared / global memory that the kernels' barriers do not
 * order -- a missing __syncthreads shows up as a data race.  tests/test_lbd_host_core.py builds and runs it.  Never shipped. */
#include "cuda_emu.h"

#include <cstdio>
#include <random>

#include "../../cube_slam_b200/csrc/cs_lbd_core.h"
namespace {
#include "../../cube_slam_b200/csrc/cs_lbd_kernels.cuh"
}

int main()
{
    std::mt19937 rng(7);
    const int w = 96, h = 64, n_lines = 12;
    std::vector<int16_t> dx((size_t)w * h), dy((size_t)w * h);
    for (auto &v : dx) v = (int16_t)((int)(rng() % 800) - 400);
    for (auto &v : dy) v = (int16_t)((int)(rng() % 800) - 400);
    std::vector<CsLbdLine> lines(n_lines);
    for (int i = 0; i < n_lines; i++) {
        const float a = (float)(rng() % 6283) / 1000.f - 3.14f;
        lines[i] = CsLbdLine{(float)(rng() % w), (float)(rng() % h), cosf(a), sinf(a), (int32_t)(1 + rng() % 60), 0};
    }
    std::vector<float> coef(84, 0.5f), fdesc((size_t)n_lines * 72);
    std::vector<uint8_t> desc((size_t)n_lines * 32);
    launch_lbd_describe(n_lines, nullptr, lines.data(), n_lines, dx.data(), dy.data(), w, h, coef.data(), desc.data(), fdesc.data());
    const int nq = 6, nt = 300;
    std::vector<uint4> q((size_t)nq * 2), t((size_t)nt * 2);
    for (auto &v : q) v = uint4{(unsigned)rng(), (unsigned)rng(), (unsigned)rng(), (unsigned)rng()};
    for (auto &v : t) v = uint4{(unsigned)rng(), (unsigned)rng(), (unsigned)rng(), (unsigned)rng()};
    for (int i = 0; i < 2 * 4; i++) t[i] = q[i % (2 * nq)]; /* some codes close to a query */
    std::vector<int32_t> pq(nq, 0), toff = {0, nt};
    std::vector<unsigned long long> keys(nq);
    launch_lbd_match(nq, nullptr, q.data(), t.data(), pq.data(), toff.data(), nq, keys.data());
    unsigned long long sum = 0;
    for (auto k : keys) sum ^= k;
    for (auto d : desc) sum += d;
    printf("racecheck done %llx\n", sum);
    return 0;
}
