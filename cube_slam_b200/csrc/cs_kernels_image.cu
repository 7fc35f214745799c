/*
 * cs_kernels_image.cu -- per-frame / per-ROI image stages of detect_cuboid for sm_100a.
 *
 *   k_bgr2gray      cv::cvtColor(BGR2GRAY)                      box_proposal_detail.cpp:62-66
 *   k_canny_nms     cv::Canny: Sobel 3x3 + L1 magnitude + NMS    box_proposal_detail.cpp:197
 *   k_canny_hyst    cv::Canny: hysteresis (order independent)    box_proposal_detail.cpp:197
 *   k_chamfer_dt    cv::distanceTransform(DIST_L2, 3)            box_proposal_detail.cpp:199
 *
 * All stages are integer / fixed-point and therefore bit-exact against the oracle restatements
 * (oracle/cuboid_oracle.cpp: orc_bgr2gray, orc_canny, orc_chamfer_dt).
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "cs_internal.h"
#include "cs_kernels.h"

/* ------------------------------------------------------------------------------------------
 * BGR -> gray.  Pure stream: 3 B read + 1 B written per pixel, the one HBM-bound kernel of the
 * path.  Contiguous frames are treated as one flat pixel array so every thread moves 48 B in
 * (three 16-byte loads) and 16 B out, all naturally aligned.
 * ------------------------------------------------------------------------------------------ */
__device__ __forceinline__ uint32_t luma15(uint32_t b, uint32_t g, uint32_t r)
{
    return (b * 3735u + g * 19235u + r * 9798u + (1u << 14)) >> 15;
}

__global__ void __launch_bounds__(256) k_bgr2gray_flat(const uint4 *__restrict__ bgr, uint4 *__restrict__ gray,
                                                        int64_t n_groups /* of 16 pixels */)
{
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n_groups; g += (int64_t)gridDim.x * blockDim.x) {
        const uint4 a = __ldcs(bgr + 3 * g), b = __ldcs(bgr + 3 * g + 1), c = __ldcs(bgr + 3 * g + 2);
        const uint32_t w[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
        uint32_t out[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            /* 4 pixels = 12 bytes = words 3q .. 3q+2 */
            const uint32_t w0 = w[3 * q], w1 = w[3 * q + 1], w2 = w[3 * q + 2];
            const uint32_t p0 = luma15(w0 & 255u, (w0 >> 8) & 255u, (w0 >> 16) & 255u);
            const uint32_t p1 = luma15(w0 >> 24, w1 & 255u, (w1 >> 8) & 255u);
            const uint32_t p2 = luma15((w1 >> 16) & 255u, w1 >> 24, w2 & 255u);
            const uint32_t p3 = luma15((w2 >> 8) & 255u, (w2 >> 16) & 255u, w2 >> 24);
            out[q] = p0 | (p1 << 8) | (p2 << 16) | (p3 << 24);
        }
        gray[g] = make_uint4(out[0], out[1], out[2], out[3]);
    }
}

/* generic (strided / tail) variant: one pixel per thread */
__global__ void __launch_bounds__(256) k_bgr2gray_generic(const uint8_t *__restrict__ img, uint8_t *__restrict__ gray, int n_frames,
                                                           int w, int h, int stride, int channels, int64_t first_pixel)
{
    const int64_t total = (int64_t)n_frames * w * h;
    for (int64_t p = first_pixel + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
        const int64_t f = p / ((int64_t)w * h);
        const int64_t r = p - f * (int64_t)w * h;
        const int y = (int)(r / w), x = (int)(r - (int64_t)y * w);
        const uint8_t *s = img + ((size_t)f * h + y) * stride + (size_t)x * channels;
        gray[p] = (channels == 3) ? (uint8_t)luma15(s[0], s[1], s[2]) : s[0];
    }
}

/* ------------------------------------------------------------------------------------------
 * Canny part 1: Sobel + |dx|+|dy| + non-maximum suppression over 32x32 tiles of each ROI.
 * The gray tile (+2 halo, replicated at the ROI border: OpenCV filters an ROI in isolation) is staged
 * in shared memory; magnitudes of the +1 halo are held there too (zero outside the ROI).
 * Map codes: 0 = not an edge, 1 = weak candidate, 2 = strong.  Strong pixels are appended to the
 * job's hysteresis queue with one atomic per warp.
 * ------------------------------------------------------------------------------------------ */
#define CT 32 /* tile edge */

__global__ void __launch_bounds__(256) k_canny_nms(const uint8_t *__restrict__ gray, int img_w, int img_h, const CsJob *__restrict__ jobs,
                                                   const CsTile *__restrict__ tiles, uint8_t *__restrict__ map_arena,
                                                   int32_t *__restrict__ queue_arena, int32_t *__restrict__ q_tail, int low, int high)
{
    __shared__ uint8_t s_g[CT + 4][CT + 4];
    __shared__ int16_t s_dx[CT + 2][CT + 2];
    __shared__ int16_t s_dy[CT + 2][CT + 2];
    __shared__ uint16_t s_m[CT + 2][CT + 2 + 2];

    const CsTile t = tiles[blockIdx.x];
    const CsJob jb = jobs[t.job];
    const int w = jb.roi_w, h = jb.roi_h;
    const int x0 = t.tx * CT, y0 = t.ty * CT;
    const uint8_t *src = gray + ((size_t)jb.frame * img_h + jb.roi_t) * img_w + jb.roi_l;
    const int tid = threadIdx.x;

    for (int i = tid; i < (CT + 4) * (CT + 4); i += 256) {
        const int ly = i / (CT + 4), lx = i - ly * (CT + 4);
        int gy = y0 + ly - 2, gx = x0 + lx - 2;
        gy = min(max(gy, 0), h - 1);
        gx = min(max(gx, 0), w - 1);
        s_g[ly][lx] = src[(size_t)gy * img_w + gx];
    }
    __syncthreads();
    for (int i = tid; i < (CT + 2) * (CT + 2); i += 256) {
        const int ly = i / (CT + 2), lx = i - ly * (CT + 2);
        const int gy = y0 + ly - 1, gx = x0 + lx - 1; /* ROI coordinates of this magnitude */
        int dx = 0, dy = 0, m = 0;
        if (gy >= 0 && gy < h && gx >= 0 && gx < w) {
            const int cy = ly + 1, cx = lx + 1; /* centre in s_g */
            const int a = s_g[cy - 1][cx - 1], b = s_g[cy - 1][cx], c = s_g[cy - 1][cx + 1];
            const int d = s_g[cy][cx - 1], f = s_g[cy][cx + 1];
            const int g = s_g[cy + 1][cx - 1], hh = s_g[cy + 1][cx], k = s_g[cy + 1][cx + 1];
            dx = (c + 2 * f + k) - (a + 2 * d + g);
            dy = (g + 2 * hh + k) - (a + 2 * b + c);
            m = abs(dx) + abs(dy);
        }
        s_dx[ly][lx] = (int16_t)dx;
        s_dy[ly][lx] = (int16_t)dy;
        s_m[ly][lx] = (uint16_t)m;
    }
    __syncthreads();

    const int TG22 = 13573; /* (int)(0.4142135623730950488016887242097 * (1 << 15) + 0.5) */
    uint8_t *map = map_arena + jb.px_off;
    int32_t *queue = queue_arena + jb.px_off;
    const int lane = tid & 31;
    for (int i = tid; i < CT * CT; i += 256) {
        const int ly = i >> 5, lx = i & 31;
        const int gy = y0 + ly, gx = x0 + lx;
        const bool in = (gy < h && gx < w);
        int code = 0;
        if (in) {
            const int my = ly + 1, mx = lx + 1;
            const int m = s_m[my][mx];
            if (m > low) {
                const int xs = s_dx[my][mx], ys = s_dy[my][mx];
                const int ax = abs(xs), ay = abs(ys) << 15;
                const int tg22x = ax * TG22;
                bool is_max;
                if (ay < tg22x) {
                    is_max = (m > s_m[my][mx - 1]) && (m >= s_m[my][mx + 1]);
                } else {
                    const int tg67x = tg22x + (ax << 16);
                    if (ay > tg67x)
                        is_max = (m > s_m[my - 1][mx]) && (m >= s_m[my + 1][mx]);
                    else {
                        const int s = ((xs ^ ys) < 0) ? -1 : 1;
                        is_max = (m > s_m[my - 1][mx - s]) && (m > s_m[my + 1][mx + s]);
                    }
                }
                if (is_max) code = (m > high) ? 2 : 1;
            }
            map[(size_t)gy * w + gx] = (uint8_t)code;
        }
        /* warp-aggregated append of strong pixels */
        const unsigned strong = __ballot_sync(0xffffffffu, code == 2);
        if (strong) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&q_tail[t.job], __popc(strong));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (code == 2) queue[base + __popc(strong & ((1u << lane) - 1u))] = gy * w + gx;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Canny part 2: hysteresis.  One CTA per ROI runs a level-synchronous flood fill from the strong
 * pixels; a weak pixel (1) is claimed exactly once by an atomicOr on its byte's word (1 -> 3).
 * The result (all weak pixels 8-connected to a strong one) does not depend on visiting order, so
 * it equals OpenCV's stack-based flood fill bit for bit.
 * ------------------------------------------------------------------------------------------ */
__global__ void __launch_bounds__(256) k_canny_hyst(const CsJob *__restrict__ jobs, uint8_t *__restrict__ map_arena,
                                                    int32_t *__restrict__ queue_arena, int32_t *__restrict__ q_tail)
{
    __shared__ int s_new;
    const CsJob jb = jobs[blockIdx.x];
    const int w = jb.roi_w, h = jb.roi_h;
    uint8_t *map = map_arena + jb.px_off;
    int32_t *queue = queue_arena + jb.px_off;
    int head = 0, tail = q_tail[blockIdx.x];
    const int tid = threadIdx.x;
    while (tail > head) {
        if (tid == 0) s_new = 0;
        __syncthreads();
        for (int i = head + tid; i < tail; i += 256) {
            const int p = queue[i];
            const int y = p / w, x = p - y * w;
#pragma unroll
            for (int dy = -1; dy <= 1; dy++) {
                const int ny = y + dy;
                if (ny < 0 || ny >= h) continue;
#pragma unroll
                for (int dx = -1; dx <= 1; dx++) {
                    if (dx == 0 && dy == 0) continue;
                    const int nx = x + dx;
                    if (nx < 0 || nx >= w) continue;
                    const size_t q = (size_t)ny * w + nx;
                    if (map[q] == 1) {
                        const uintptr_t addr = (uintptr_t)(map + q);
                        unsigned *word = (unsigned *)(addr & ~(uintptr_t)3);
                        const unsigned sh = (unsigned)(addr & 3) * 8u;
                        const unsigned old = atomicOr(word, 2u << sh);
                        if (((old >> sh) & 255u) == 1u) queue[tail + atomicAdd(&s_new, 1)] = ny * w + nx;
                    }
                }
            }
        }
        __syncthreads();
        head = tail;
        tail += s_new;
        __syncthreads();
    }
}

/* ------------------------------------------------------------------------------------------
 * 3x3 chamfer distance transform, two raster passes in 16.16 fixed point (a = 62587, b = 89738).
 * One warp per ROI; a lane owns PPL consecutive columns of the current row in registers.  The
 * in-row recurrence tmp[j] = min(u[j], tmp[j-1] + a) is a min-plus prefix scan:
 *      tmp[j] = a*j + min_{k<=j} (u[k] - a*k)
 * evaluated as a local scan + a 5-step warp scan, all in exact integer arithmetic, so the result
 * equals the sequential OpenCV loop bit for bit.  "Infinity" is a sentinel BIG (> any reachable
 * distance) that is re-clamped every row; it is emitted as OpenCV's saturated DIST_MAX.
 * ------------------------------------------------------------------------------------------ */
#define DT_HV 62587
#define DT_DG 89738
#define DT_BIG (1 << 30)

template <int PPL>
__device__ __forceinline__ void dt_warp(const uint8_t *__restrict__ map, uint32_t *__restrict__ tmp, int w, int h)
{
    const int lane = threadIdx.x & 31;
    const int c0 = lane * PPL;
    const unsigned FULL = 0xffffffffu;
    int up[PPL];
#pragma unroll
    for (int k = 0; k < PPL; k++) up[k] = DT_BIG;

    /* forward pass */
    for (int i = 0; i < h; i++) {
        const uint8_t *mrow = map + (size_t)i * w;
        int upL = __shfl_up_sync(FULL, up[PPL - 1], 1);
        int upR = __shfl_down_sync(FULL, up[0], 1);
        if (lane == 0) upL = DT_BIG;
        if (lane == 31) upR = DT_BIG;
        int v[PPL];
#pragma unroll
        for (int k = 0; k < PPL; k++) {
            const int col = c0 + k;
            const int l = (k == 0) ? upL : up[k - 1];
            const int r = (k == PPL - 1) ? upR : up[k + 1];
            int u = min(min(l + DT_DG, up[k] + DT_HV), r + DT_DG);
            if (col < w) {
                if (mrow[col] & 2) u = 0;
            } else
                u = DT_BIG;
            v[k] = u - DT_HV * col;
        }
#pragma unroll
        for (int k = 1; k < PPL; k++) v[k] = min(v[k], v[k - 1]);
        int incl = v[PPL - 1];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl = min(incl, o);
        }
        int excl = __shfl_up_sync(FULL, incl, 1);
        if (lane == 0) excl = INT_MAX;
        uint32_t *trow = tmp + (size_t)i * w;
#pragma unroll
        for (int k = 0; k < PPL; k++) {
            const int col = c0 + k;
            int t = min(v[k], excl) + DT_HV * col;
            if (t >= DT_BIG || col >= w) t = DT_BIG;
            up[k] = t;
            if (col < w) trow[col] = (uint32_t)t;
        }
    }

    /* backward pass: `up` now plays the role of the row below */
#pragma unroll
    for (int k = 0; k < PPL; k++) up[k] = DT_BIG;
    const float scale = 1.f / 65536.f;
    const float dist_max = (float)(0xffffffffu - (uint32_t)DT_DG) * scale;
    for (int i = h - 1; i >= 0; i--) {
        uint32_t *trow = tmp + (size_t)i * w;
        int dnL = __shfl_up_sync(FULL, up[PPL - 1], 1);
        int dnR = __shfl_down_sync(FULL, up[0], 1);
        if (lane == 0) dnL = DT_BIG;
        if (lane == 31) dnR = DT_BIG;
        int v[PPL];
#pragma unroll
        for (int k = 0; k < PPL; k++) {
            const int col = c0 + k;
            const int l = (k == 0) ? dnL : up[k - 1];
            const int r = (k == PPL - 1) ? dnR : up[k + 1];
            int cur = (col < w) ? (int)trow[col] : DT_BIG;
            cur = min(min(cur, r + DT_DG), min(up[k] + DT_HV, l + DT_DG));
            if (col >= w) cur = DT_BIG;
            v[k] = cur + DT_HV * col;
        }
#pragma unroll
        for (int k = PPL - 2; k >= 0; k--) v[k] = min(v[k], v[k + 1]);
        int incl = v[0];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_down_sync(FULL, incl, d);
            if (lane + d < 32) incl = min(incl, o);
        }
        int excl = __shfl_down_sync(FULL, incl, 1);
        if (lane == 31) excl = INT_MAX;
        float *drow = reinterpret_cast<float *>(trow);
#pragma unroll
        for (int k = 0; k < PPL; k++) {
            const int col = c0 + k;
            int t = min(v[k], excl) - DT_HV * col;
            if (t >= DT_BIG || col >= w) t = DT_BIG;
            up[k] = t;
            if (col < w) drow[col] = (t >= DT_BIG) ? dist_max : (float)(uint32_t)t * scale;
        }
    }
}

template <int PPL>
__global__ void __launch_bounds__(128) k_chamfer_dt(const CsJob *__restrict__ jobs, const int32_t *__restrict__ job_ids, int n_ids,
                                                    const uint8_t *__restrict__ map_arena, float *__restrict__ dist_arena)
{
    const int slot = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (slot >= n_ids) return;
    const CsJob jb = jobs[job_ids[slot]];
    const uint8_t *map = map_arena + jb.px_off;
    uint32_t *tmp = reinterpret_cast<uint32_t *>(dist_arena + jb.px_off);
    if (jb.roi_w <= 0 || jb.roi_h <= 0) return;
    dt_warp<PPL>(map, tmp, jb.roi_w, jb.roi_h);
}

/* ------------------------------------------------------------------------------------------ launchers */
void cs_launch_gray(const uint8_t *d_img, uint8_t *d_gray, int n_frames, int w, int h, int stride, int channels, cudaStream_t st,
                    int64_t *launches)
{
    const int64_t total = (int64_t)n_frames * w * h;
    int64_t done = 0;
    if (channels == 3 && stride == 3 * w && (((uintptr_t)d_img) & 15) == 0 && (((uintptr_t)d_gray) & 15) == 0) {
        const int64_t groups = total / 16;
        if (groups > 0) {
            const int blocks = (int)((groups + 255) / 256 < 148 * 16 ? (groups + 255) / 256 : 148 * 16);
            k_bgr2gray_flat<<<blocks, 256, 0, st>>>((const uint4 *)d_img, (uint4 *)d_gray, groups);
            (*launches)++;
        }
        done = groups * 16;
    }
    if (done < total) {
        const int64_t rest = total - done;
        const int blocks = (int)((rest + 255) / 256 < 148 * 16 ? (rest + 255) / 256 : 148 * 16);
        k_bgr2gray_generic<<<blocks, 256, 0, st>>>(d_img, d_gray, n_frames, w, h, stride, channels, done);
        (*launches)++;
    }
}

void cs_launch_canny(const uint8_t *d_gray, int img_w, int img_h, const CsJob *d_jobs, int n_jobs, const CsTile *d_tiles, int n_tiles,
                     uint8_t *d_map, int32_t *d_queue, int32_t *d_qtail, int low, int high, cudaStream_t st, int64_t *launches)
{
    if (n_jobs <= 0) return;
    cudaMemsetAsync(d_qtail, 0, sizeof(int32_t) * n_jobs, st);
    if (n_tiles > 0) {
        k_canny_nms<<<n_tiles, 256, 0, st>>>(d_gray, img_w, img_h, d_jobs, d_tiles, d_map, d_queue, d_qtail, low, high);
        (*launches)++;
    }
}

void cs_launch_hyst(const CsJob *d_jobs, int n_jobs, uint8_t *d_map, int32_t *d_queue, int32_t *d_qtail, cudaStream_t st, int64_t *launches)
{
    if (n_jobs <= 0) return;
    k_canny_hyst<<<n_jobs, 256, 0, st>>>(d_jobs, d_map, d_queue, d_qtail);
    (*launches)++;
}

/* width classes of the distance transform: a lane owns PPL columns */
const int cs_dt_class_ppl[CS_DT_CLASSES] = {4, 8, 12, 16, 24, 40, 64};

int cs_dt_class_of(int roi_w)
{
    for (int c = 0; c < CS_DT_CLASSES; c++)
        if (roi_w <= 32 * cs_dt_class_ppl[c]) return c;
    return -1;
}

template <int PPL>
static void launch_dt_class(const CsJob *d_jobs, const int32_t *d_ids, int n, const uint8_t *d_map, float *d_dist, cudaStream_t st,
                            int64_t *launches)
{
    if (n <= 0) return;
    const int wpb = (PPL >= 24) ? 2 : 4;
    k_chamfer_dt<PPL><<<(n + wpb - 1) / wpb, wpb * 32, 0, st>>>(d_jobs, d_ids, n, d_map, d_dist);
    (*launches)++;
}

/* d_ids: job ids grouped by class; class c occupies [class_off[c], class_off[c+1]) */
void cs_launch_dt(const CsJob *d_jobs, const int32_t *d_ids, const int *class_off, const uint8_t *d_map, float *d_dist, cudaStream_t st,
                  int64_t *launches)
{
    launch_dt_class<4>(d_jobs, d_ids + class_off[0], class_off[1] - class_off[0], d_map, d_dist, st, launches);
    launch_dt_class<8>(d_jobs, d_ids + class_off[1], class_off[2] - class_off[1], d_map, d_dist, st, launches);
    launch_dt_class<12>(d_jobs, d_ids + class_off[2], class_off[3] - class_off[2], d_map, d_dist, st, launches);
    launch_dt_class<16>(d_jobs, d_ids + class_off[3], class_off[4] - class_off[3], d_map, d_dist, st, launches);
    launch_dt_class<24>(d_jobs, d_ids + class_off[4], class_off[5] - class_off[4], d_map, d_dist, st, launches);
    launch_dt_class<40>(d_jobs, d_ids + class_off[5], class_off[6] - class_off[5], d_map, d_dist, st, launches);
    launch_dt_class<64>(d_jobs, d_ids + class_off[6], class_off[7] - class_off[6], d_map, d_dist, st, launches);
}
