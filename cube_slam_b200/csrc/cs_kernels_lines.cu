/*
 * cs_kernels_lines.cu -- per-ROI line selection and merging.
 *
 *   k_roi_lines: align_left_right_edges (object_3d_util.cpp:147-158) -> both-endpoints-inside-ROI filter
 *                (box_proposal_detail.cpp:166-174) -> merge_break_lines (object_3d_util.cpp:300-376)
 *                -> per-line angle + midpoint (box_proposal_detail.cpp:185-191).
 *
 * One CTA per ROI job; the ROI's line set lives in shared memory.  merge_break_lines is a
 * sequential-semantics algorithm ("merge the lexicographically first mergeable pair, restart").
 * The kernel reproduces exactly that pair sequence, but instead of rescanning every pair after
 * each merge it keeps two markers:
 *     d : the row whose content changed in the last merge  (pairs (a,d), a<d, and row d must be re-tested)
 *     F : rows in (d, F) are known to contain no mergeable pair; rows >= F are unknown
 * so the number of pair tests is O(m^2 + merges*m) instead of O(merges*m^2), and every round of
 * tests is spread across the CTA with the first hit selected by an atomicMin on (row,col).
 */
#include "cs_pmath.h"
#include <cuda_runtime.h>
#include <stdint.h>

#include "cs_internal.h"
#include "cs_kernels.h"

#define CS_PI 3.14159265358979323846
#define LN_THREADS 128

struct LineSet {
    double x1[CS_LINE_CAP], y1[CS_LINE_CAP], x2[CS_LINE_CAP], y2[CS_LINE_CAP], ang[CS_LINE_CAP];
};

__device__ __forceinline__ double dmin(double a, double b) { return (b < a) ? b : a; } /* std::min */

/* the body of the seg1/seg2 test of merge_break_lines (object_3d_util.cpp:319-342) */
__device__ __forceinline__ bool merge_test(const LineSet &L, int s1, int s2, double angle_thre, double dist_thre, double &msx, double &msy,
                                           double &mex, double &mey, double &merged_angle)
{
    const double diff = fabs(L.ang[s1] - L.ang[s2]);
    const double angle_diff = dmin(diff, CS_PI - diff);
    if (!(angle_diff < angle_thre)) return false;
    const double ax = L.x2[s1] - L.x1[s2], ay = L.y2[s1] - L.y1[s2];
    const double bx = L.x2[s2] - L.x1[s1], by = L.y2[s2] - L.y1[s1];
    const double d12 = sqrt(ax * ax + ay * ay);
    const double d21 = sqrt(bx * bx + by * by);
    if (!((d12 < dist_thre) || (d21 < dist_thre))) return false;
    if (L.x1[s1] < L.x1[s2]) {
        msx = L.x1[s1];
        msy = L.y1[s1];
    } else {
        msx = L.x1[s2];
        msy = L.y1[s2];
    }
    if (L.x2[s1] > L.x2[s2]) {
        mex = L.x2[s1];
        mey = L.y2[s1];
    } else {
        mex = L.x2[s2];
        mey = L.y2[s2];
    }
    merged_angle = cs_pm_atan2(mey - msy, mex - msx);
    const double temp = fabs(L.ang[s1] - merged_angle);
    const double merge_angle_diff = dmin(temp, CS_PI - temp);
    return merge_angle_diff < angle_thre;
}

extern __shared__ unsigned char ln_smem_raw[];

__global__ void __launch_bounds__(LN_THREADS) k_roi_lines(const CsJob *__restrict__ jobs, const CsFrame *__restrict__ frames,
                                                          const double *__restrict__ lines /* batch, M x 4 */, const float *__restrict__ lines_f32 /* online mode: detector output */,
                                                          const int32_t *__restrict__ n_lines_dev, int f32_pitch, double *__restrict__ out_lines,
                                                          int32_t *__restrict__ out_counts /* n_jobs x 2: inside, merged */,
                                                          int32_t *__restrict__ err_flag, double dist_thre, double angle_thre_deg,
                                                          double len_thre)
{
    LineSet &L = *reinterpret_cast<LineSet *>(ln_smem_raw);
    __shared__ int s_warp_cnt[LN_THREADS / 32];
    __shared__ int s_total;
    __shared__ int s_best;

    const int job = blockIdx.x;
    const CsJob jb = jobs[job];
    CsFrame fr = frames[jb.frame];
    if (n_lines_dev) { /* online mode: the line detector left n x 4 float32 per frame at a fixed pitch */
        fr.n_lines = n_lines_dev[jb.frame];
        if (fr.n_lines > f32_pitch) { /* the detector found more segments than the context's line capacity */
            if (threadIdx.x == 0) atomicOr(err_flag, 4);
            fr.n_lines = f32_pitch;
        }
        fr.line_off = jb.frame * f32_pitch;
    }
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const double bl = jb.roi_l, bt = jb.roi_t, br = jb.roi_r, bb = jb.roi_b;

    /* ---- ordered selection of the frame's lines that lie inside the ROI ------------------- */
    if (tid == 0) s_total = 0;
    __syncthreads();
    for (int base = 0; base < fr.n_lines; base += LN_THREADS) {
        const int i = base + tid;
        bool in = false;
        double x1 = 0, y1 = 0, x2 = 0, y2 = 0;
        if (i < fr.n_lines) {
            if (lines_f32) { /* float -> double exactly as main_obj.cpp:429-432 */
                const float *p = lines_f32 + (size_t)(fr.line_off + i) * 4;
                x1 = p[0];
                y1 = p[1];
                x2 = p[2];
                y2 = p[3];
            } else {
                const double *p = lines + (size_t)(fr.line_off + i) * 4;
                x1 = p[0];
                y1 = p[1];
                x2 = p[2];
                y2 = p[3];
            }
            if (x2 < x1) { /* align_left_right_edges */
                double t = x1;
                x1 = x2;
                x2 = t;
                t = y1;
                y1 = y2;
                y2 = t;
            }
            in = (bl <= x1 && x1 <= br && bt <= y1 && y1 <= bb) && (bl <= x2 && x2 <= br && bt <= y2 && y2 <= bb);
        }
        const unsigned m = __ballot_sync(0xffffffffu, in);
        if (lane == 0) s_warp_cnt[wid] = __popc(m);
        __syncthreads();
        int off = s_total;
        for (int k = 0; k < wid; k++) off += s_warp_cnt[k];
        if (in) {
            const int slot = off + __popc(m & ((1u << lane) - 1u));
            if (slot < CS_LINE_CAP) {
                L.x1[slot] = x1;
                L.y1[slot] = y1;
                L.x2[slot] = x2;
                L.y2[slot] = y2;
                L.ang[slot] = cs_pm_atan2(y2 - y1, x2 - x1);
            }
        }
        __syncthreads();
        if (tid == 0) {
            int t = s_total;
            for (int k = 0; k < LN_THREADS / 32; k++) t += s_warp_cnt[k];
            s_total = t;
        }
        __syncthreads();
    }
    int total = s_total;
    const int n_inside = total;
    if (total > CS_LINE_CAP) {
        if (tid == 0) atomicOr(err_flag, 1);
        total = CS_LINE_CAP;
    }

    /* ---- merge_break_lines ------------------------------------------------------------------ */
    const double angle_thre = angle_thre_deg / 180.0 * CS_PI;
    int d = -1, F = 0, counter = 0;
    const int KEY = CS_LINE_CAP;
    bool more = true;
    while (more && counter < 500) {
        counter++;
        int hit = INT_MAX;
        /* round 0: pairs (a, d) for a < d, then row d */
        if (d >= 0) {
            if (tid == 0) s_best = INT_MAX;
            __syncthreads();
            double t0, t1, t2, t3, t4;
            for (int a = tid; a < d; a += LN_THREADS)
                if (merge_test(L, a, d, angle_thre, dist_thre, t0, t1, t2, t3, t4)) atomicMin(&s_best, a * KEY + d);
            for (int b = d + 1 + tid; b < total; b += LN_THREADS)
                if (merge_test(L, d, b, angle_thre, dist_thre, t0, t1, t2, t3, t4)) atomicMin(&s_best, d * KEY + b);
            __syncthreads();
            hit = s_best;
            __syncthreads();
        }
        /* rows >= F, eight rows per round */
        int row0 = F;
        while (hit == INT_MAX && row0 < total - 1) {
            if (tid == 0) s_best = INT_MAX;
            __syncthreads();
            const int RB = 8;
            double t0, t1, t2, t3, t4;
            for (int idx = tid; idx < RB * total; idx += LN_THREADS) {
                const int a = row0 + idx / total, b = idx % total;
                if (a < total - 1 && b > a && merge_test(L, a, b, angle_thre, dist_thre, t0, t1, t2, t3, t4)) atomicMin(&s_best, a * KEY + b);
            }
            __syncthreads();
            hit = s_best;
            __syncthreads();
            if (hit == INT_MAX) row0 += RB;
        }
        if (hit == INT_MAX) {
            more = false;
        } else {
            const int a = hit / KEY, b = hit % KEY;
            /* clean frontier before the hit */
            int Feff = F;
            if (a >= F) Feff = a + 1;
            if (tid == 0) {
                double msx, msy, mex, mey, mang;
                merge_test(L, a, b, angle_thre, dist_thre, msx, msy, mex, mey, mang);
                L.x1[a] = msx;
                L.y1[a] = msy;
                L.x2[a] = mex;
                L.y2[a] = mey;
                L.ang[a] = mang; /* == atan2 of the stored endpoints, what the next restart recomputes */
                const int last = total - 1; /* fast_RemoveRow (matrix_utils.cpp:172-176) */
                L.x1[b] = L.x1[last];
                L.y1[b] = L.y1[last];
                L.x2[b] = L.x2[last];
                L.y2[b] = L.y2[last];
                L.ang[b] = L.ang[last];
            }
            total--;
            d = a;
            F = min(Feff, b);
            if (F < d + 1) F = d + 1;
            __syncthreads();
        }
    }

    /* ---- drop short lines, emit angle + midpoint ------------------------------------------- */
    if (tid == 0) s_total = 0;
    __syncthreads();
    double *o = out_lines + (size_t)job * CS_MAXL_OUT * 7;
    for (int base = 0; base < total; base += LN_THREADS) {
        const int i = base + tid;
        bool keep = false;
        if (i < total) {
            const double dx = L.x2[i] - L.x1[i], dy = L.y2[i] - L.y1[i];
            keep = (len_thre > 0) ? (sqrt(dx * dx + dy * dy) > len_thre) : true;
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_warp_cnt[wid] = __popc(m);
        __syncthreads();
        int off = s_total;
        for (int k = 0; k < wid; k++) off += s_warp_cnt[k];
        if (keep) {
            const int slot = off + __popc(m & ((1u << lane) - 1u));
            if (slot < CS_MAXL_OUT) {
                o[0 * CS_MAXL_OUT + slot] = L.x1[i];
                o[1 * CS_MAXL_OUT + slot] = L.y1[i];
                o[2 * CS_MAXL_OUT + slot] = L.x2[i];
                o[3 * CS_MAXL_OUT + slot] = L.y2[i];
                o[4 * CS_MAXL_OUT + slot] = L.ang[i];
                o[5 * CS_MAXL_OUT + slot] = (L.x1[i] + L.x2[i]) / 2;
                o[6 * CS_MAXL_OUT + slot] = (L.y1[i] + L.y2[i]) / 2;
            }
        }
        __syncthreads();
        if (tid == 0) {
            int t = s_total;
            for (int k = 0; k < LN_THREADS / 32; k++) t += s_warp_cnt[k];
            s_total = t;
        }
        __syncthreads();
    }
    if (tid == 0) {
        int n = s_total;
        if (n > CS_MAXL_OUT) {
            atomicOr(err_flag, 2);
            n = CS_MAXL_OUT;
        }
        out_counts[job * 2 + 0] = n_inside;
        out_counts[job * 2 + 1] = n;
    }
}

void cs_launch_roi_lines(const CsJob *d_jobs, int n_jobs, const CsFrame *d_frames, const double *d_lines, const float *d_lines_f32,
                         const int32_t *d_n_lines_dev, int f32_pitch, double *d_out_lines,
                         int32_t *d_out_counts, int32_t *d_err, double dist_thre, double angle_thre_deg, double len_thre, cudaStream_t st,
                         int64_t *launches)
{
    if (n_jobs <= 0) return;
    CS_ONCE_PER_DEVICE(cudaFuncSetAttribute(k_roi_lines, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LineSet)));
    CS_APPLY_CARVEOUT(k_roi_lines);
    k_roi_lines<<<n_jobs, LN_THREADS, sizeof(LineSet), st>>>(d_jobs, d_frames, d_lines, d_lines_f32, d_n_lines_dev, f32_pitch, d_out_lines, d_out_counts, d_err, dist_thre,
                                                             angle_thre_deg, len_thre);
    (*launches)++;
}
