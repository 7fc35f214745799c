/* tests/host_core/lbd_core_host.cpp -- the product's descriptor / matcher arithmetic (cube_slam_b200/csrc/cs_lbd_core.h, the functions the
 * CUDA kernels k_lbd_describe / k_lbd_match are made of) compiled for the HOST, driven the way the kernels drive it: one "thread" index at
 * a time, phase after phase, shared-memory arrays as locals.  tests/test_lbd_host_core.py compares the result with the oracle bit for bit,
 * so the kernels' arithmetic, indexing and phase split are checked on machines without a GPU.  Test infrastructure, never shipped.
 * g++ -O2 -ffp-contract=off (IEEE + - * / sqrt, like nvcc -fmad=false -prec-div=true -prec-sqrt=true on the device). */
#include <cstdint>
#include <cstring>

#include "../../cube_slam_b200/csrc/cs_lbd_core.h"

/* k_lbd_describe, CTA by CTA: lines = n x CsLbdLine (24 bytes, from cs_lbd_debug_prepare), coef = F_g 63 then F_l 21 */
extern "C" void host_lbd_describe(const void *lines_, int n_lines, const int16_t *dx_all, const int16_t *dy_all, int w, int h, const float *coef, uint8_t *desc,
                                  float *fdesc)
{
    const CsLbdLine *lines = (const CsLbdLine *)lines_;
    for (int li = 0; li < n_lines; li++) {
        float s_rows[CS_LBD_ROWS * 4], s_sums[CS_LBD_DESC], s_des[CS_LBD_DESC], s_coefL[3 * CS_LBD_BAND_WIDTH];
        const CsLbdLine L = lines[li];
        for (int tid = 0; tid < 64; tid++) {
            if (tid < 3 * CS_LBD_BAND_WIDTH) s_coefL[tid] = coef[CS_LBD_ROWS + tid];
            if (tid < CS_LBD_ROWS) {
                const size_t off = (size_t)L.frame * w * h;
                float r[4];
                cs_lbd_row(L, tid, dx_all + off, dy_all + off, w, h, coef[tid], r);
                for (int k = 0; k < 4; k++) s_rows[tid * 4 + k] = r[k];
            }
        }
        for (int tid = 0; tid < 64; tid++)
            for (int t = tid; t < CS_LBD_DESC; t += 64) s_sums[t] = cs_lbd_band_sum(t, s_rows, s_coefL);
        for (int tid = 0; tid < 64; tid++)
            if (tid < CS_LBD_BANDS) cs_lbd_band_stats(tid, s_sums, s_des);
        cs_lbd_finish(s_des);
        for (int tid = 0; tid < 64; tid++) {
            if (tid < CS_LBD_BYTES) desc[(size_t)li * CS_LBD_BYTES + tid] = cs_lbd_byte(tid, s_des);
            if (fdesc)
                for (int t = tid; t < CS_LBD_DESC; t += 64) fdesc[(size_t)li * CS_LBD_DESC + t] = s_des[t];
        }
    }
}

/* k_lbd_match, CTA by CTA: keys[qi] = the smallest key over the query's train set */
extern "C" void host_lbd_match(const uint8_t *q_all, const uint8_t *t_all, const int32_t *pair_of_query, const int32_t *t_off, int n_queries, uint64_t *keys)
{
    for (int qi = 0; qi < n_queries; qi++) {
        uint32_t q[8];
        std::memcpy(q, q_all + (size_t)qi * 32, 32);
        const int p = pair_of_query[qi], t0 = t_off[p], t1 = t_off[p + 1];
        uint64_t s_best[128];
        for (int tid = 0; tid < 128; tid++) {
            uint64_t best = ~0ull;
            for (int j = t0 + tid; j < t1; j += 128) {
                uint32_t t[8];
                std::memcpy(t, t_all + (size_t)j * 32, 32);
                const uint64_t key = cs_lbd_match_key(q, t, (uint32_t)(j - t0));
                best = key < best ? key : best;
            }
            s_best[tid] = best;
        }
        uint64_t b = ~0ull;
        for (int tid = 0; tid < 128; tid++) b = s_best[tid] < b ? s_best[tid] : b;
        keys[qi] = b;
    }
}

extern "C" int host_lbd_round(float x) { return cs_lbd_round(x); }
extern "C" int host_lbd_key_dist(uint64_t key) { return CS_LBD_KEY_DIST(key); }
extern "C" uint32_t host_lbd_key_train(uint64_t key) { return CS_LBD_KEY_TRAIN(key); }
