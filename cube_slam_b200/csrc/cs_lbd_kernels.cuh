/* cs_lbd_kernels.cuh -- the two kernels of cs_lbd.cu (included there, inside its unnamed namespace): k_lbd_describe and k_lbd_match.
 * They live in a file of their own so that the CPU test suite can compile this very source against a small emulation of the CUDA
 * execution model (tests/host_core/lbd_kernels_emu.cpp: one std::thread per CUDA thread, a std::barrier for __syncthreads, function-local
 * statics for __shared__) and run whole launches against the oracle -- index arithmetic, phase split and barrier placement included -- on
 * machines without a GPU.  The arithmetic itself is cs_lbd_core.h; see cs_lbd.cu's header for what each kernel does. */
__global__ void __launch_bounds__(64) k_lbd_describe(const CsLbdLine *__restrict__ lines, int n_lines, const int16_t *__restrict__ dx_all,
                                                     const int16_t *__restrict__ dy_all, int w, int h, const float *__restrict__ coef /* F_g 63, F_l 21 */,
                                                     uint8_t *__restrict__ desc, float *__restrict__ fdesc)
{
    __shared__ float s_rows[CS_LBD_ROWS * 4];
    __shared__ float s_sums[CS_LBD_DESC];
    __shared__ float s_des[CS_LBD_DESC];
    __shared__ float s_coefL[3 * CS_LBD_BAND_WIDTH];
    const int li = blockIdx.x, tid = threadIdx.x;
    if (li >= n_lines) return; /* the whole CTA leaves together */
    const CsLbdLine L = lines[li];
    if (tid < 3 * CS_LBD_BAND_WIDTH) s_coefL[tid] = coef[CS_LBD_ROWS + tid];
    if (tid < CS_LBD_ROWS) {
        const size_t off = (size_t)L.frame * w * h;
        float r[4];
        cs_lbd_row(L, tid, dx_all + off, dy_all + off, w, h, coef[tid], r);
        s_rows[tid * 4 + 0] = r[0];
        s_rows[tid * 4 + 1] = r[1];
        s_rows[tid * 4 + 2] = r[2];
        s_rows[tid * 4 + 3] = r[3];
    }
    __syncthreads();
    for (int t = tid; t < CS_LBD_DESC; t += 64) s_sums[t] = cs_lbd_band_sum(t, s_rows, s_coefL);
    __syncthreads();
    if (tid < CS_LBD_BANDS) cs_lbd_band_stats(tid, s_sums, s_des);
    __syncthreads();
    if (tid == 0) cs_lbd_finish(s_des);
    __syncthreads();
    if (tid < CS_LBD_BYTES) desc[(size_t)li * CS_LBD_BYTES + tid] = cs_lbd_byte(tid, s_des);
    if (fdesc)
        for (int t = tid; t < CS_LBD_DESC; t += 64) fdesc[(size_t)li * CS_LBD_DESC + t] = s_des[t];
}

__global__ void __launch_bounds__(128) k_lbd_match(const uint4 *__restrict__ q_all, const uint4 *__restrict__ t_all, const int32_t *__restrict__ pair_of_query,
                                                   const int32_t *__restrict__ t_off, int n_queries, unsigned long long *__restrict__ keys)
{
    __shared__ unsigned long long s_best[4];
    const int qi = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (qi >= n_queries) return;
    const uint4 qa = q_all[2 * (size_t)qi], qb = q_all[2 * (size_t)qi + 1];
    const uint32_t q[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
    const int p = pair_of_query[qi], t0 = t_off[p], t1 = t_off[p + 1];
    unsigned long long best = ~0ull;
    for (int j = t0 + tid; j < t1; j += 128) {
        const uint4 ta = t_all[2 * (size_t)j], tb = t_all[2 * (size_t)j + 1];
        const uint32_t t[8] = {ta.x, ta.y, ta.z, ta.w, tb.x, tb.y, tb.z, tb.w};
        const unsigned long long key = cs_lbd_match_key(q, t, (uint32_t)(j - t0));
        best = key < best ? key : best;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other < best ? other : best;
    }
    if (lane == 0) s_best[wid] = best;
    __syncthreads();
    if (tid == 0) {
        unsigned long long b = s_best[0];
        for (int k = 1; k < 4; k++) b = s_best[k] < b ? s_best[k] : b;
        keys[qi] = b;
    }
}

/* the two launches: a CTA of 64 threads per key line, a CTA of 128 threads per query.  Under the CPU emulation of the test suite
 * (CS_LBD_EMU_LAUNCH defined by tests/host_core/*.cpp) the same grids run as threads of the host. */
#if defined(__CUDACC__)
inline void launch_lbd_describe(unsigned grid, cudaStream_t st, const CsLbdLine *lines, int n_lines, const int16_t *dx_all, const int16_t *dy_all, int w, int h,
                                const float *coef, uint8_t *desc, float *fdesc)
{
    k_lbd_describe<<<grid, 64, 0, st>>>(lines, n_lines, dx_all, dy_all, w, h, coef, desc, fdesc);
}
inline void launch_lbd_match(unsigned grid, cudaStream_t st, const uint4 *q_all, const uint4 *t_all, const int32_t *pair_of_query, const int32_t *t_off, int n_queries,
                             unsigned long long *keys)
{
    k_lbd_match<<<grid, 128, 0, st>>>(q_all, t_all, pair_of_query, t_off, n_queries, keys);
}
#elif defined(CS_LBD_EMU_LAUNCH)
inline void launch_lbd_describe(unsigned grid, cudaStream_t, const CsLbdLine *lines, int n_lines, const int16_t *dx_all, const int16_t *dy_all, int w, int h,
                                const float *coef, uint8_t *desc, float *fdesc)
{
    CS_LBD_EMU_LAUNCH(grid, 64, [&] { k_lbd_describe(lines, n_lines, dx_all, dy_all, w, h, coef, desc, fdesc); });
}
inline void launch_lbd_match(unsigned grid, cudaStream_t, const uint4 *q_all, const uint4 *t_all, const int32_t *pair_of_query, const int32_t *t_off, int n_queries,
                             unsigned long long *keys)
{
    CS_LBD_EMU_LAUNCH(grid, 128, [&] { k_lbd_match(q_all, t_all, pair_of_query, t_off, n_queries, keys); });
}
#endif
