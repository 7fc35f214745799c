/*
 * cs_lsd.cu -- line_lbd_detect::detect_filter_lines, LSD flavour, for sm_100a (kernel group (i) of the north star).
 *
 * Replaces, for one octave (the only one filter_lines keeps, line_lbd/class/line_lbd_allclass.cpp:200-207):
 *   LSDDetector::detectImpl                line_lbd/libs/LSDDetector.cpp:153-256
 *   LineSegmentDetectorImpl::flsd & co.    line_lbd/libs/lsd.cpp:440-1154   (LSD_REFINE_ADV, default parameters)
 *   filter_lines + keylines_to_mat         line_lbd/class/line_lbd_allclass.cpp:26-36,200-221
 *
 * Streaming stages (FP64, evaluation order of OpenCV's C paths, -fmad=false):
 *   k_lsd_blur    cvtColor + the 7 x 7 Gaussian (sigma 0.6/0.8), both passes, on 64 x 16 tiles; interior tiles' BGR bytes come in by TMA
 *                 (k_lsd_hblur / k_lsd_vblur: the two-pass pair of round 1, A/B path, cs_set_profiling bit 7)     lsd.cpp:452-457
 *   k_lsd_resize  cv::resize(x0.8, INTER_LINEAR) on doubles               lsd.cpp:459
 *   k_lsd_grad    2x2 gradient, modulus, fastAtan2 angle                  lsd.cpp:562-586
 *   (no ordering pass: ll_angle's 1024-bin pseudo-ordering, lsd.cpp:588-634, links the pixels by gradient bin, but flsd walks the node
 *   vector by index, lsd.cpp:478-480, i.e. in the raster order the nodes were allocated in -- established in round 2 by compiling the
 *   reference's own lsd.cpp, oracle/ref/; the counting-sort kernels of round 1 are gone)
 *
 * Seed loop (lsd.cpp:476-535).  The reference visits the ordered pixel list one seed at a time; a seed grows a region over the pixels no
 * earlier seed used, so the result is defined by the order -- but only through the `used` map.  It is cut in two:
 *   k_lsd_grow_seq   everything that reads or writes `used`: the raster scan for seeds, region_grow, region2rect, the density test, refine /
 *                    reduce_region_radius (lsd.cpp:478-519), one warp per frame (one 32-thread CTA, 8 KB of shared memory for the region
 *                    list, `used` as a bit per pixel in HBM read through L2).  Emits the candidate rectangles in seed order.
 *   k_lsd_val_count / k_lsd_val_nfa (six rounds)   rect_improve and the NFA test (lsd.cpp:520-534, 873-1136) read the angle map only: one
 *                    warp per undecided candidate rectangle, all frames at once; a round's up-to-five rectangles counted in one scan of the
 *                    angle map, the binomial tails in a kernel of their own (k_lsd_validate: the same as ONE kernel, kept as the A/B path --
 *                    its 6.6 k instructions against a 32 KB instruction cache cost 40 % of its issue stalls).
 *   k_lsd_emit       accepted candidates in seed order + the key-line filter.
 * Why this shape (profiles/r2_lsd_seed_ncu.md): the one-warp loop is bound by instruction issue (an instruction every 6 cycles, ~12 M per
 * frame in round 1), not by memory latency; what stops more frames from sharing an SM is the instruction cache, so the sequential kernel is
 * kept small (grow -> rectangle -> density is ONE two-pass loop, 7.4 k instructions instead of 39 k) and everything order-free runs where
 * thousands of warps execute the same code.  An ordered-speculation kernel (many warps per frame claiming pixels by rank) was built and
 * measured in this round and removed: 74 % of the candidates it started were refused and redone, it needed ~25 barrier-separated rounds
 * per frame, was not faster than one warp per frame, and emitted duplicate segments on dense frames (git history: k_lsd_grow_par).
 *
 * Inside one candidate the warp parallelises what is order-free (the 3x3 neighbour tests of three region points per step from ONE 16-byte
 * record per pixel, the addends of the ordered sums, rectangle pixel counts over rows, min/max extents, the binomial tail's break tests) and
 * keeps every floating-point accumulation in the reference's order; neighbours whose angle difference is clear of the tolerance by more
 * than the region angle can drift within a round are accepted / rejected without re-deriving the angle per pixel (lsd_region_grow).
 */
#include <cuda_runtime.h>
#include <float.h>
#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "cs_internal.h"
#include "cs_nfa.cuh"
#include "cs_tma.cuh"

#define LSD_PI 3.1415926535897932384626433832795
#define LSD_NOTDEF (-1024.0)
#define LSD_DEG2RAD (LSD_PI / 180)
#define LSD_3_2_PI ((3 * LSD_PI) / 2)
#define LSD_2PI (2 * LSD_PI)
#define LSD_LN10 2.30258509299404568402
#define LSD_NBINS 1024
#define LSD_CHUNK_ROWS 8

#define LSD_CAND_CAP 2048    /* candidate rectangles per frame handed from the seed loop to k_lsd_validate */
#define LSD_SEQ_SCAP 2048    /* region entries k_lsd_grow_seq keeps in shared memory (the rest spill to HBM; small, so that many frames share an SM) */
#define LSD_HDR 8            /* ints of a candidate record header in the arena: n1, n2, has_line, x1 y1 x2 y2 (float bits), pad */

namespace {

/* cv2 4.x getGaussianKernel(7, 0.6 / 0.8, CV_64F): lsd.cpp:453 divides, sigma = 0.7499999999999999, not 0.75 */
__constant__ double c_gauss7[7] = {0x1.763496d347532p-13, 0x1.f1e23259cfdc1p-7, 0x1.bfd7fac1bd5a8p-3, 0x1.10562a79786afp-1,
                                   0x1.bfd7fac1bd5a8p-3, 0x1.f1e23259cfdc1p-7, 0x1.763496d347532p-13};

__device__ __forceinline__ int reflect101(int p, int n)
{
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = (p < 0) ? -p : 2 * n - 2 - p;
    return p;
}

/* cv::fastAtan2 (degrees) */
__device__ __forceinline__ float fast_atan2(float y, float x)
{
    const float p1 = 0.9997878412794807f * (float)(180 / LSD_PI);
    const float p3 = -0.3258083974640975f * (float)(180 / LSD_PI);
    const float p5 = 0.1555786518463281f * (float)(180 / LSD_PI);
    const float p7 = -0.04432655554792128f * (float)(180 / LSD_PI);
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

/* ---------------------------------------------------------------------------------------- streaming stages */
__global__ void __launch_bounds__(256) k_lsd_hblur(const uint8_t *__restrict__ img, int n_frames, int w, int h, int stride, int channels,
                                                   double *__restrict__ tmp)
{
    /* grid: x over the pixels of one frame (32-bit arithmetic), y = frame */
    const int f = blockIdx.y;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= w * h) return;
    const size_t p = (size_t)f * w * h + r;
    {
        const int y = r / w, x = r - y * w;
        const uint8_t *row = img + ((size_t)f * h + y) * stride;
        double s = 0;
#pragma unroll
        for (int k = 0; k < 7; k++) {
            const int xx = reflect101(x + k - 3, w);
            int g;
            if (channels == 3) {
                const uint8_t *q = row + 3 * xx;
                g = (int)((q[0] * 3735u + q[1] * 19235u + q[2] * 9798u + (1u << 14)) >> 15);
            } else
                g = row[xx];
            const double t = c_gauss7[k] * (double)g;
            s = (k == 0) ? t : s + t;
        }
        tmp[p] = s;
    }
}

__global__ void __launch_bounds__(256) k_lsd_vblur(const double *__restrict__ tmp, int n_frames, int w, int h, double *__restrict__ blur)
{
    /* grid: x over the pixels of one frame (32-bit arithmetic), y = frame */
    const int f = blockIdx.y;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= w * h) return;
    const size_t p = (size_t)f * w * h + r;
    {
        const int y = r / w, x = r - y * w;
        const double *base = tmp + (size_t)f * w * h;
        double s = c_gauss7[3] * base[(size_t)y * w + x];
#pragma unroll
        for (int k = 1; k <= 3; k++) s += c_gauss7[3 + k] * (base[(size_t)reflect101(y + k, h) * w + x] + base[(size_t)reflect101(y - k, h) * w + x]);
        blur[p] = s;
    }
}

/* cvtColor + the 7 x 7 Gaussian (both passes) on 64 x 16 tiles: the gray value of a pixel is computed once (k_lsd_hblur recomputes it for
 * each of its seven taps) and the horizontal sums never leave shared memory.  Same operations in the same order as k_lsd_hblur followed
 * by k_lsd_vblur (which stay as the A/B path, cs_set_profiling bit 7). */
#define LSB_TW 64
#define LSB_TH 16
/* kTma: the BGR bytes of an interior tile (no reflection needed) are fetched by the copy engine -- one cp.async.bulk.tensor.2d of
 * 224 x 22 bytes, completion on an mbarrier -- instead of three byte loads per pixel; border tiles keep the reflected loads. */
#define LSB_BOXW 240 /* 3 * (LSB_TW + 6) = 210 bytes of BGR + up to 15 of alignment slack (a TMA box starts at a multiple of 16 bytes), a multiple of 16 */
template <bool kTma>
__global__ void __launch_bounds__(256) k_lsd_blur(const __grid_constant__ CUtensorMap tmap, const uint8_t *__restrict__ img, int w, int h, int stride,
                                                  int channels, double *__restrict__ blur, int32_t *__restrict__ err_flag)
{
    __shared__ uint8_t s_g[LSB_TH + 6][LSB_TW + 8];
    __shared__ double s_h[LSB_TH + 6][LSB_TW];
    __shared__ __align__(128) uint8_t s_rgb[kTma ? (LSB_TH + 6) * LSB_BOXW : 16];
    __shared__ __align__(8) unsigned long long s_bar;
    const int f = blockIdx.z, x0 = blockIdx.x * LSB_TW, y0 = blockIdx.y * LSB_TH, tid = threadIdx.x;
    const uint8_t *frame = img + (size_t)f * h * stride;
    const bool interior = kTma && x0 - 3 >= 0 && x0 + LSB_TW + 3 <= w - 1 && y0 - 3 >= 0 && y0 + LSB_TH + 3 <= h - 1;
    if (interior) {
        if (tid == 0) cs_mbar_init(&s_bar);
        __syncthreads();
        const int bx = 3 * (x0 - 3), boff = bx & 15; /* the box starts at a multiple of 16 bytes */
        if (tid == 0) cs_tma_load_2d(&tmap, s_rgb, &s_bar, bx - boff, f * h + y0 - 3, (LSB_TH + 6) * LSB_BOXW);
        if (!cs_mbar_wait(&s_bar, 0) && tid == 0) atomicOr(err_flag, 8);
        for (int i = tid; i < (LSB_TH + 6) * (LSB_TW + 6); i += 256) {
            const int r = i / (LSB_TW + 6), c = i - r * (LSB_TW + 6);
            const uint8_t *q = &s_rgb[r * LSB_BOXW + boff + 3 * c];
            s_g[r][c] = (uint8_t)((q[0] * 3735u + q[1] * 19235u + q[2] * 9798u + (1u << 14)) >> 15);
        }
    } else
        for (int i = tid; i < (LSB_TH + 6) * (LSB_TW + 6); i += 256) {
            const int r = i / (LSB_TW + 6), c = i - r * (LSB_TW + 6);
            const int yy = reflect101(y0 - 3 + r, h), xx = reflect101(x0 - 3 + c, w);
            const uint8_t *q = frame + (size_t)yy * stride;
            uint32_t g;
            if (channels == 3) {
                q += 3 * xx;
                g = (q[0] * 3735u + q[1] * 19235u + q[2] * 9798u + (1u << 14)) >> 15;
            } else
                g = q[xx];
            s_g[r][c] = (uint8_t)g;
        }
    __syncthreads();
    for (int i = tid; i < (LSB_TH + 6) * LSB_TW; i += 256) {
        const int r = i / LSB_TW, c = i - r * LSB_TW;
        double s = 0;
#pragma unroll
        for (int k = 0; k < 7; k++) {
            const double t = c_gauss7[k] * (double)(int)s_g[r][c + k];
            s = (k == 0) ? t : s + t;
        }
        s_h[r][c] = s;
    }
    __syncthreads();
    for (int i = tid; i < LSB_TH * LSB_TW; i += 256) {
        const int r = i / LSB_TW, c = i - r * LSB_TW;
        const int x = x0 + c, y = y0 + r;
        if (x >= w || y >= h) continue;
        double s = c_gauss7[3] * s_h[r + 3][c];
#pragma unroll
        for (int k = 1; k <= 3; k++) s += c_gauss7[3 + k] * (s_h[r + 3 + k][c] + s_h[r + 3 - k][c]);
        blur[((size_t)f * h + y) * w + x] = s;
    }
}

__global__ void __launch_bounds__(256) k_lsd_resize(const double *__restrict__ blur, int n_frames, int sw, int sh, int dw, int dh, double inv_scale,
                                                    double *__restrict__ scaled)
{
    /* grid: x over the pixels of one scaled frame, y = frame */
    const int f = blockIdx.y;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= dw * dh) return;
    const size_t p = (size_t)f * dw * dh + r;
    {
        const int dy = r / dw, dx = r - dy * dw;
        float fx = (float)((dx + 0.5) * inv_scale - 0.5);
        int sx = (int)floorf(fx);
        fx -= sx;
        bool single = false;
        if (sx < 0) {
            fx = 0;
            sx = 0;
        }
        if (sx + 1 >= sw) {
            single = true; /* dx >= xmax: one tap (OpenCV's HResize tail) */
            if (sx >= sw - 1) {
                fx = 0;
                sx = sw - 1;
            }
        }
        float fy = (float)((dy + 0.5) * inv_scale - 0.5);
        int sy = (int)floorf(fy);
        fy -= sy;
        const int y0 = (sy >= 0) ? (sy < sh ? sy : sh - 1) : 0;
        const int y1 = (sy + 1 >= 0) ? (sy + 1 < sh ? sy + 1 : sh - 1) : 0;
        const double *S0 = blur + ((size_t)f * sh + y0) * sw, *S1 = blur + ((size_t)f * sh + y1) * sw;
        const float a0 = 1.f - fx, a1 = fx, b0 = 1.f - fy, b1 = fy;
        double r0, r1;
        if (!single) {
            r0 = S0[sx] * a0 + S0[sx + 1] * a1;
            r1 = S1[sx] * a0 + S1[sx + 1] * a1;
        } else {
            r0 = S0[sx] * 1.0;
            r1 = S1[sx] * 1.0;
        }
        scaled[p] = r0 * b0 + r1 * b1;
    }
}

/* Per pixel: modulus (f64), level-line angle as the float fastAtan2 returns it (degrees; -1 = NOTDEF; the reference's double angle is
 * (double)deg * DEG2RAD, recomputed where needed) and the 16-byte growth record {deg, cos, sin, claim}: (cos, sin) of float(angle),
 * each the correctly rounded float -- what region_grow accumulates (lsd.cpp:680-681). */
__global__ void __launch_bounds__(256) k_lsd_grad(const double *__restrict__ scaled, int n_frames, int W, int H, double threshold,
                                                  double *__restrict__ modgrad, float *__restrict__ angf, uint4 *__restrict__ pix)
{
    /* grid: x over the pixels of one scaled frame, y = frame */
    const int f = blockIdx.y;
    const int addr = blockIdx.x * blockDim.x + threadIdx.x;
    const bool inside = addr < W * H;
    const size_t p = (size_t)f * W * H + addr;
    {
        const int y = addr / W, x = addr - y * W;
        const double *im = scaled + (size_t)f * W * H;
        double norm = 0;
        float deg = -1.f;
        if (inside && x < W - 1 && y < H - 1) {
            const double DA = im[addr + W + 1] - im[addr];
            const double BC = im[addr + 1] - im[addr + W];
            const double gx = DA + BC, gy = DA - BC;
            norm = sqrt((gx * gx + gy * gy) / 4);
            if (!(norm <= threshold)) deg = fast_atan2((float)gx, (float)(-gy));
        }
        if (!inside) return;
        modgrad[p] = norm;
        angf[p] = deg;
        uint4 r = make_uint4(__float_as_uint(deg), 0u, 0u, 0u);
        if (deg >= 0.f) {
            const double ang = (double)deg * LSD_DEG2RAD;
            const double af = (double)(float)ang;
            r.y = __float_as_uint((float)cos(af));
            r.z = __float_as_uint((float)sin(af));
        }
        pix[p] = r;
    }
}

/* ---------------------------------------------------------------------------------------- the seed loop */
/* cycle counters of the seed loop's phases (diagnostics, cs_debug_lsd_prof): grow, region2rect, refine, rectangle counts, binomial tails,
 * candidates, list scan */
__device__ unsigned long long g_lsd_prof[16];
#define LSD_PROF_T0() const long long prof_t0__ = clock64()
#define LSD_PROF_ADD(slot)                                                                         \
    do {                                                                                           \
        if ((threadIdx.x & 31) == 0) atomicAdd(&g_lsd_prof[slot], (unsigned long long)(clock64() - prof_t0__)); \
    } while (0)

struct LsdFrame {
    int W, H;
    uint4 *pix;            /* {deg, cos, sin, 0}: one 16-byte record per neighbour test */
    const float *angf;     /* deg plane for the rectangle scans */
    const double *modgrad;
    double LOG_NT;
    /* the `used` map as a bit per pixel (HBM, read and written through L2 only, so that it costs no shared memory: the seed loop issues
     * one instruction every ~6 cycles, and the more frames share an SM the better) */
    uint32_t *ubits;
    struct LsdSpan *span; /* k_lsd_validate: room for five row-span records per warp (shared memory) */
    double *stage;        /* k_lsd_grow_seq: 96 doubles of shared memory per warp (ordered sums) */
    __device__ __forceinline__ bool used_bit(int addr) const { return (__ldcg(ubits + (addr >> 5)) >> (addr & 31)) & 1u; }
    const double *lgam; /* log_gamma of small integers (cs_nfa.cuh) */
    unsigned long long wmagic; /* ceil(2^40 / W): row of a pixel address without an integer division (exact for addresses < 2^20 .. 2^30 / W) */
    __device__ __forceinline__ int row_of(int addr) const { return (int)(((unsigned long long)addr * wmagic) >> 40); }
};

/* the region list of the candidate a warp works on: the first `scap` entries in shared memory, the rest in HBM */
struct LsdReg {
    int *s;
    int *g;
    int scap;
    int cap;
    __device__ __forceinline__ int get(int i) const { return i < scap ? s[i] : __ldcg(g + i); }
    __device__ __forceinline__ void put(int i, int v) const
    {
        if (i < scap)
            s[i] = v;
        else
            g[i] = v;
    }
};

__device__ __forceinline__ void lsd_prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

struct LsdRect {
    double x1, y1, x2, y2, width, x, y, theta, dx, dy, prec, p;
};

__device__ __forceinline__ double lsd_dist_sq(double x1, double y1, double x2, double y2) { return (x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1); }
__device__ __forceinline__ double lsd_dist(double x1, double y1, double x2, double y2) { return sqrt(lsd_dist_sq(x1, y1, x2, y2)); }
__device__ __forceinline__ double lsd_angle_diff_signed(double a, double b)
{
    double diff = a - b;
    while (diff <= -LSD_PI) diff += LSD_2PI;
    while (diff > LSD_PI) diff -= LSD_2PI;
    return diff;
}
__device__ __forceinline__ bool lsd_double_equal(double a, double b)
{
    if (a == b) return true;
    const double abs_diff = fabs(a - b);
    const double aa = fabs(a), bb = fabs(b);
    double abs_max = (aa > bb) ? aa : bb;
    if (abs_max < DBL_MIN) abs_max = DBL_MIN;
    return (abs_diff / abs_max) <= (100.0 * DBL_EPSILON);
}
/* lsd.cpp:1138-1154 on the float level-line angle in degrees (negative = NOTDEF).  The decision is the reference's double comparison; a
 * float evaluation of the same difference (error < 1e-5 rad) settles every case that is not within 2e-4 rad of the tolerance, so the
 * double arithmetic (half-rate pipe, and this test runs for every pixel of every rectangle scan and every neighbour of every growth step)
 * is only executed on the rare borderline pixel. */
__device__ __forceinline__ bool lsd_aligned_deg(float deg, double theta, double prec)
{
    if (deg < 0.f) return false;
    {
        float nf = fabsf((float)theta - deg * 0.017453292f);
        if (nf > 4.712389f) nf = fabsf(nf - 6.2831855f);
        const float pf = (float)prec;
        if (nf < pf - 2e-4f) return true;
        if (nf > pf + 2e-4f) return false;
    }
    const double a = (double)deg * LSD_DEG2RAD;
    double n_theta = theta - a;
    if (n_theta < 0) n_theta = -n_theta;
    if (n_theta > LSD_3_2_PI) {
        n_theta -= LSD_2PI;
        if (n_theta < 0) n_theta = -n_theta;
    }
    return n_theta <= prec;
}

__device__ __noinline__ double lsd_log_gamma(double x)
{
    if (x > 15.0) return 0.918938533204673 + (x - 0.5) * log(x) - x + 0.5 * x * log(x * sinh(1 / x) + 1 / (810.0 * pow(x, 6.0)));
    const double q[7] = {75122.6331530, 80916.6278952, 36308.2951477, 8687.24529705, 1168.92649479, 83.8676043424, 2.50662827511};
    double a = (x + 0.5) * log(x + 5.5) - (x + 5.5);
    double b = 0;
    for (int n = 0; n < 7; ++n) {
        a -= log(x + (double)n);
        b += q[n] * pow(x, (double)n);
    }
    return a + log(b);
}

/* lsd.cpp:1100-1136; called with different (n, k, p) on different lanes */
__device__ __noinline__ double lsd_nfa(int n, int k, double p, double LOG_NT)
{
    if (n == 0 || k == 0) return -LOG_NT;
    if (n == k) return -LOG_NT - (double)n * log10(p);
    const double p_term = p / (1 - p);
    const double log1term = ((double)n + 1) - lsd_log_gamma((double)k + 1) - lsd_log_gamma((double)(n - k) + 1) + (double)k * log(p) + (double)(n - k) * log(1.0 - p);
    double term = exp(log1term);
    if (lsd_double_equal(term, 0)) {
        if (k > n * p) return -log1term / LSD_LN10 - LOG_NT;
        return -LOG_NT;
    }
    double bin_tail = term;
    const double tolerance = 0.1;
    for (int i = k + 1; i <= n; ++i) {
        const double bin_term = (double)(n - i + 1) / (double)i;
        const double mult_term = bin_term * p_term;
        term *= mult_term;
        bin_tail += term;
        if (bin_term < 1) {
            const double err = term * ((1 - pow(mult_term, (double)(n - i + 1))) / (1 - mult_term) - 1);
            if (err < tolerance * fabs(-log10(bin_tail) - LOG_NT) * bin_tail) break;
        }
    }
    return -log10(bin_tail) - LOG_NT;
}

/* lsd.cpp:637-688, one warp.  Three region points per round, their 3 x 3 neighbourhoods on lanes 0..26 in the reference's (point, yy, xx)
 * order, so "the first lane" is "the next pixel the reference would test".
 *
 * The reference updates the region angle after EVERY added pixel and tests the next neighbour against the new angle: a chain of
 * fastAtan2 -> compare -> add per pixel.  Most of those tests cannot come out differently, though.  Let S be the running sum of unit
 * vectors, L = |S|.  A pixel that passes the test lies within A = prec + D of the (computed) region direction, so adding it turns S by at
 * most sin(A + e) / L <= (A + e) / L and does not shorten it (e <= 1e-3 rad: error of the fastAtan2 polynomial, measured 1.7e-4).  With
 * at most m additions in this round, every angle the round will test against is within
 *     D = m (prec + 0.2 + e) / L + 2 e + float slack
 * of the angle at the start of the round (D <= 0.2 required).  A neighbour whose difference from the round-start angle is below prec - D
 * is accepted whenever its turn comes, one above prec + D is rejected whenever: neither needs the angle at its turn.  Only the pixels in
 * between are tested the reference's way, against fastAtan2 of the sums as they stand at their turn (the sums are always added in the
 * reference's order, in float, so they are the reference's bits).  prec + D stays below pi / 2, where the reference's wrapped difference
 * (lsd.cpp:1138-1154) equals the circular distance or exceeds pi / 2, so the argument holds across the 0 / 2 pi seam. */
__device__ void lsd_region_grow(const LsdFrame &F, const LsdReg &R, int base, int s_addr, int &reg_size, double &reg_angle, double prec)
{
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    float reg_deg = F.angf[s_addr];
    reg_angle = (double)reg_deg * LSD_DEG2RAD;
    float sumdx = (float)cos(reg_angle);
    float sumdy = (float)sin(reg_angle);
    __syncwarp(); /* every lane is done reading the previous region's list before its slots are written again */
    if (lane == 0) {
        atomicOr(F.ubits + (s_addr >> 5), 1u << (s_addr & 31));
        R.put(base, s_addr);
    }
    reg_size = 1;
    __syncwarp();
    const int grp = lane / 9, kk = lane - grp * 9;
    const int ky = kk / 3 - 1, kx = kk - (kk / 3) * 3 - 1; /* (yy, xx) in the reference's loop order */
    const float precf = (float)prec;
    const unsigned lt_mask = (1u << lane) - 1u;
    for (int i = 0; i < reg_size;) {
        const int navail = min(3, reg_size - i);
        bool cand = false;
        int c_addr = -1;
        float deg = -1.f, csx = 0.f, csy = 0.f;
        if (grp < navail) {
            const int pa = R.get(base + i + grp);
            const int py = F.row_of(pa), px = pa - py * F.W;
            const int yy = py + ky, xx = px + kx;
            if (yy >= 0 && yy < F.H && xx >= 0 && xx < F.W) {
                c_addr = yy * F.W + xx;
                const uint4 r = __ldg(F.pix + c_addr); /* {deg, cos, sin, -}: read-only */
                const uint32_t uw = __ldcg(F.ubits + (c_addr >> 5)); /* issued together with the record: one round trip per round, not two */
                deg = __uint_as_float(r.x);
                if (deg >= 0.f && !((uw >> (c_addr & 31)) & 1u)) {
                    cand = true;
                    csx = __uint_as_float(r.y);
                    csy = __uint_as_float(r.z);
                }
            }
        }
        unsigned maybe = __ballot_sync(FULL, cand);
        if (maybe) {
            /* the float difference from the round-start angle (error < 1e-5 rad) and the drift bound D */
            float nf = fabsf(reg_deg * 0.017453292f - deg * 0.017453292f);
            if (nf > 4.712389f) nf = fabsf(nf - 6.2831855f);
            const float inv_l = rsqrtf(sumdx * sumdx + sumdy * sumdy) * 1.001f;
            const float per_add = (precf + 0.201f) * inv_l;
            unsigned sure = 0u;
            float D = (float)__popc(maybe) * per_add + 0.0022f;
            if (D <= 0.2f && precf + D < 1.5f) {
                const unsigned m1 = __ballot_sync(FULL, cand && nf < precf + D); /* everything else is rejected whatever the angle */
                D = (float)__popc(m1) * per_add + 0.0022f;                          /* at most popc(m1) additions this round */
                sure = __ballot_sync(FULL, cand && nf < precf - D);
                maybe = __ballot_sync(FULL, cand && nf < precf + D);
            }
            unsigned accepted = 0u;
            bool stale = false; /* sums changed since reg_deg was computed */
            while (maybe) {
                const int fl = __ffs(maybe) - 1;
                maybe &= ~(1u << fl);
                if (!((sure >> fl) & 1u)) {
                    /* the reference's own test, at this pixel's turn */
                    if (stale) {
                        reg_deg = fast_atan2(sumdy, sumdx);
                        stale = false;
                    }
                    const bool ok = lsd_aligned_deg(deg, (double)reg_deg * LSD_DEG2RAD, prec);
                    if (!((__ballot_sync(FULL, ok) >> fl) & 1u)) continue;
                }
                const int addrf = __shfl_sync(FULL, c_addr, fl);
                /* cos(float(angle)), sin(float(angle)): precomputed per pixel by k_lsd_grad (pinned to the correctly rounded float) */
                sumdx += __shfl_sync(FULL, csx, fl);
                sumdy += __shfl_sync(FULL, csy, fl);
                stale = true;
                accepted |= 1u << fl;
                maybe &= ~__ballot_sync(FULL, c_addr == addrf); /* the same pixel seen from a later point of this round */
            }
            if (accepted) {
                if ((accepted >> lane) & 1u) {
                    R.put(base + reg_size + __popc(accepted & lt_mask), c_addr);
                    atomicOr(F.ubits + (c_addr >> 5), 1u << (c_addr & 31));
                    /* this pixel is a region point now: its 3 x 3 neighbourhood will be gathered a few rounds from here.  Ask L2 for the three
                     * 48-byte runs of records (a DRAM miss is ~3x an L2 hit, and the gather is on the warp's critical path). */
                    const uint4 *q = F.pix + c_addr - 1;
                    if (c_addr >= F.W + 1) lsd_prefetch_l2(q - F.W);
                    lsd_prefetch_l2(q);
                    if (c_addr + F.W + 1 < F.W * F.H) lsd_prefetch_l2(q + F.W);
                }
                reg_size += __popc(accepted);
                if (stale) reg_deg = fast_atan2(sumdy, sumdx);
                __syncwarp();
            }
        }
        i += navail;
    }
    reg_angle = (double)reg_deg * LSD_DEG2RAD;
}

/* Sums over the region in the reference's order (the order decides the last bits of a double sum): 32 region points at a time, every lane
 * computes the addend(s) of ITS point -- the products round the same wherever they are computed -- and stages them in shared memory
 * (F.stage, 96 doubles per warp); then the additions run as one chain over the staged values, read as broadcasts. */
/* lsd.cpp:690-784 */
__device__ void lsd_region2rect(const LsdFrame &F, const LsdReg &R, int base, int reg_size, double reg_angle, double prec, double p, LsdRect &rec)
{
    const int lane = threadIdx.x & 31;
    double *st = F.stage;
    double x = 0, y = 0, sum = 0;
    for (int i0 = 0; i0 < reg_size; i0 += 32) {
        const int n = min(32, reg_size - i0);
        __syncwarp();
        if (lane < n) {
            const int addr = R.get(base + i0 + lane);
            const double weight = F.modgrad[addr];
            const int ry = F.row_of(addr), rx = addr - ry * F.W;
            st[lane] = (double)rx * weight;
            st[32 + lane] = (double)ry * weight;
            st[64 + lane] = weight;
        }
        __syncwarp();
        for (int j = 0; j < n; j++) {
            x += st[j];
            y += st[32 + j];
            sum += st[64 + j];
        }
    }
    x /= sum;
    y /= sum;
    /* get_theta */
    double Ixx = 0.0, Iyy = 0.0, Ixy = 0.0;
    for (int i0 = 0; i0 < reg_size; i0 += 32) {
        const int n = min(32, reg_size - i0);
        __syncwarp();
        if (lane < n) {
            const int addr = R.get(base + i0 + lane);
            const double weight = F.modgrad[addr];
            const int ry = F.row_of(addr), rx = addr - ry * F.W;
            const double ddx = (double)rx - x, ddy = (double)ry - y;
            st[lane] = ddy * ddy * weight;
            st[32 + lane] = ddx * ddx * weight;
            st[64 + lane] = ddx * ddy * weight;
        }
        __syncwarp();
        for (int j = 0; j < n; j++) {
            Ixx += st[j];
            Iyy += st[32 + j];
            Ixy -= st[64 + j];
        }
    }
    const double lambda = 0.5 * (Ixx + Iyy - sqrt((Ixx - Iyy) * (Ixx - Iyy) + 4.0 * Ixy * Ixy));
    double theta = (fabs(Ixx) > fabs(Iyy)) ? (double)fast_atan2((float)(lambda - Ixx), (float)Ixy) : (double)fast_atan2((float)Ixy, (float)(lambda - Iyy));
    theta *= LSD_DEG2RAD;
    if (fabs(lsd_angle_diff_signed(theta, reg_angle)) > prec) theta += LSD_PI;
    const double dx = cos(theta), dy = sin(theta);
    /* extents: min / max are order-free, so lanes split the region */
    double l_min = 0, l_max = 0, w_min = 0, w_max = 0;
    for (int i = lane; i < reg_size; i += 32) {
        const int addr = R.get(base + i);
        const int ry = F.row_of(addr), rx = addr - ry * F.W;
        const double regdx = (double)rx - x, regdy = (double)ry - y;
        const double l = regdx * dx + regdy * dy;
        const double w = -regdx * dy + regdy * dx;
        l_max = fmax(l_max, l);
        l_min = fmin(l_min, l);
        w_max = fmax(w_max, w);
        w_min = fmin(w_min, w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        l_max = fmax(l_max, __shfl_xor_sync(0xffffffffu, l_max, o));
        l_min = fmin(l_min, __shfl_xor_sync(0xffffffffu, l_min, o));
        w_max = fmax(w_max, __shfl_xor_sync(0xffffffffu, w_max, o));
        w_min = fmin(w_min, __shfl_xor_sync(0xffffffffu, w_min, o));
    }
    rec.x1 = x + l_min * dx;
    rec.y1 = y + l_min * dy;
    rec.x2 = x + l_max * dx;
    rec.y2 = y + l_max * dy;
    rec.width = w_max - w_min;
    rec.x = x;
    rec.y = y;
    rec.theta = theta;
    rec.dx = dx;
    rec.dy = dy;
    rec.prec = prec;
    rec.p = p;
    if (rec.width < 1.0) rec.width = 1.0;
}

__device__ __noinline__ void lsd_region2rect_cold(const LsdFrame &F, const LsdReg &R, int base, int reg_size, double reg_angle, double prec, double p, LsdRect &rec);

/* lsd.cpp:834-871 (lane 0 replays the reference's in-place compaction; it fixes the order later sums run in).
 * Returns 0 ok, 3 region rejected. */
__device__ __noinline__ int lsd_reduce_region_radius(const LsdFrame &F, const LsdReg &R, int base, int &reg_size, double reg_angle, double prec, double p,
                                        LsdRect &rec, double density, double density_th)
{
    const int lane = threadIdx.x & 31;
    const int a0 = R.get(base);
    const double xc = (double)(a0 - F.row_of(a0) * F.W), yc = (double)F.row_of(a0);
    const double radSq1 = lsd_dist_sq(xc, yc, rec.x1, rec.y1), radSq2 = lsd_dist_sq(xc, yc, rec.x2, rec.y2);
    double radSq = radSq1 > radSq2 ? radSq1 : radSq2;
    while (density < density_th) {
        radSq *= 0.75 * 0.75;
        int rs = reg_size;
        if (lane == 0) {
            for (int i = 0; i < rs; ++i) {
                const int addr = R.get(base + i);
                if (lsd_dist_sq(xc, yc, (double)(addr - F.row_of(addr) * F.W), (double)F.row_of(addr)) > radSq) {
                    atomicAnd(F.ubits + (addr >> 5), ~(1u << (addr & 31)));
                    const int last = R.get(base + rs - 1);
                    R.put(base + rs - 1, addr);
                    R.put(base + i, last);
                    --rs;
                    --i;
                }
            }
        }
        reg_size = __shfl_sync(0xffffffffu, rs, 0);
        __syncwarp();
        if (reg_size < 2) return 3;
        lsd_region2rect_cold(F, R, base, reg_size, reg_angle, prec, p, rec);
        density = (double)reg_size / (lsd_dist(rec.x1, rec.y1, rec.x2, rec.y2) * rec.width);
    }
    return 0;
}

/* lsd.cpp:977-1098 with the vendored slips kept: the (total, aligned) pixel counts of a rectangle.  The edge stepping of the
 * reference adds integer-valued steps (its slopes are int / int divisions) to integer starts, once per row INSIDE the image, so the
 * span of a row has a closed form (LsdSpan) and rows are scanned by separate lanes. */
struct LsdSpan {
    int mx, ly, ry, y0, y1; /* start column, the rows where the left / right edge changes slope, first and last row inside the image */
    int fl, sl, fr, sr;     /* the four integer steps */
};

__device__ void lsd_span_setup(const LsdFrame &F, const LsdRect &rec, LsdSpan &P)
{
    const double half_width = rec.width / 2.0;
    const double dyhw = rec.dy * half_width, dxhw = rec.dx * half_width;
    int ox[4], oy[4];
    ox[0] = (int)(rec.x1 - dyhw);
    oy[0] = (int)(rec.y1 + dxhw);
    ox[1] = (int)(rec.x2 - dyhw);
    oy[1] = (int)(rec.y2 + dxhw);
    ox[2] = (int)(rec.x2 + dyhw);
    oy[2] = (int)(rec.y2 - dxhw);
    ox[3] = (int)(rec.x1 + dyhw);
    oy[3] = (int)(rec.y1 - dxhw);
    /* sort by (x, y) */
#pragma unroll
    for (int i = 1; i < 4; i++) {
        const int vx = ox[i], vy = oy[i];
        int j = i - 1;
        while (j >= 0 && ((vx == ox[j]) ? (vy < oy[j]) : (vx < ox[j]))) {
            ox[j + 1] = ox[j];
            oy[j + 1] = oy[j];
            j--;
        }
        ox[j + 1] = vx;
        oy[j + 1] = vy;
    }
    int imin = 0, imax = 0;
    for (int i = 1; i < 4; ++i) {
        if (oy[imin] > oy[i]) imin = i;
        if (oy[imax] < oy[i]) imax = i;
    }
    unsigned taken = 1u << imin;
    int ileft = -1;
    for (int i = 0; i < 4; ++i)
        if (!((taken >> i) & 1u)) {
            if (ileft < 0)
                ileft = i;
            else if (ox[ileft] > ox[i])
                ileft = i;
        }
    taken |= 1u << ileft;
    int iright = -1;
    for (int i = 0; i < 4; ++i)
        if (!((taken >> i) & 1u)) {
            if (iright < 0)
                iright = i;
            else if (ox[iright] < ox[i])
                iright = i;
        }
    taken |= 1u << iright;
    int itail = -1;
    for (int i = 0; i < 4; ++i)
        if (!((taken >> i) & 1u)) {
            if (itail < 0)
                itail = i;
            else if (ox[itail] > ox[i])
                itail = i;
        }
    const int mx = ox[imin], my = oy[imin], lx = ox[ileft], ly = oy[ileft], rx = ox[iright], ry = oy[iright], tx = ox[itail];
    P.mx = mx;
    P.ly = ly;
    P.ry = ry;
    P.fl = (my != ly) ? (mx - lx) / (my - ly) : 0;
    P.sl = (ly != tx) ? (lx - tx) / (ly - tx) : 0;
    P.fr = (my != ry) ? (mx - rx) / (my - ry) : 0;
    P.sr = (ry != tx) ? (rx - tx) / (ry - tx) : 0;
    /* rows outside the image skip the edge stepping too (as the reference): only rows y0..y1 count */
    P.y0 = max(my, 0);
    P.y1 = min(oy[imax], F.H - 1);
}

/* the columns [lo, hi] of row y (y0 <= y <= y1); empty when hi < lo */
__device__ __forceinline__ void lsd_span_row(const LsdSpan &P, int y, int W, int &lo, int &hi)
{
    /* steps added before row y: one per earlier inside row y' in [y0, y); row y' adds the second slope iff y' >= ly (ry) */
    const long long n_l2 = (long long)max(0, y - max(P.ly, P.y0)), n_l1 = (long long)(y - P.y0) - n_l2;
    const long long n_r2 = (long long)max(0, y - max(P.ry, P.y0)), n_r1 = (long long)(y - P.y0) - n_r2;
    const long long left_x = (long long)P.mx + n_l1 * (long long)P.fl + n_l2 * (long long)P.sl;
    const long long right_x = (long long)P.mx + n_r1 * (long long)P.fr + n_r2 * (long long)P.sr;
    lo = (int)(left_x > 0 ? left_x : 0);
    hi = (int)(right_x < (long long)(W - 1) ? right_x : (long long)(W - 1));
}

__device__ void lsd_rect_count(const LsdFrame &F, const LsdRect &rec, int &total_pts, int &alg_pts)
{
    const int lane = threadIdx.x & 31;
    LsdSpan P;
    lsd_span_setup(F, rec, P);
    int tot = 0, alg = 0;
    /* lanes over rows AND over the pixels of a row: G lanes share a row, G the largest power of two with rows * G <= 32 (a nearly horizontal
     * segment has a handful of long rows, a nearly vertical one many short rows) */
    const int n_rows = P.y1 - P.y0 + 1;
    int G = 1;
    while (G < 32 && n_rows * G * 2 <= 32) G <<= 1;
    const int rows_per_step = 32 / G, sub = lane & (G - 1), rsel = lane / G;
    for (int yb = P.y0; yb <= P.y1; yb += rows_per_step) {
        const int y = yb + rsel;
        if (y > P.y1) continue;
        int lo, hi;
        lsd_span_row(P, y, F.W, lo, hi);
        if (hi >= lo) {
            if (sub == 0) tot += hi - lo + 1;
            const float *row = F.angf + (size_t)y * F.W;
            for (int x = lo + sub; x <= hi; x += G) alg += lsd_aligned_deg(row[x], rec.theta, rec.prec) ? 1 : 0;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        tot += __shfl_xor_sync(0xffffffffu, tot, o);
        alg += __shfl_xor_sync(0xffffffffu, alg, o);
    }
    total_pts = tot;
    alg_pts = alg;
}

/* The (total, aligned) counts of the up to five rectangles of one rect_improve phase in ONE scan of the angle map.  Within a phase the
 * rectangles share the direction theta and differ either in the tolerance only (lsd.cpp:889-903, 957-972: p halved, same geometry) or in
 * the geometry only (lsd.cpp:905-955: width reduced / one side moved, same tolerance): the level-line angle of a pixel is loaded and
 * compared once, the per-rectangle part is a threshold or a column-range test.  The spans are built by lanes 0..n-1 in parallel and
 * shared through F.span (shared memory, 5 LsdSpan per warp). */
__device__ void lsd_rect_count_multi(const LsdFrame &F, const LsdRect *r, int n, bool same_geom, int *tot, int *alg)
{
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    LsdSpan *sp = F.span;
    const int n_geom = same_geom ? 1 : n;
    __syncwarp();
    if (lane < n_geom) lsd_span_setup(F, r[lane], sp[lane]);
    __syncwarp();
    int y_min = sp[0].y0, y_max = sp[0].y1;
    for (int t = 1; t < n_geom; t++) {
        y_min = min(y_min, sp[t].y0);
        y_max = max(y_max, sp[t].y1);
    }
    const double theta = r[0].theta;
    const float thetaf = (float)theta;
    float pf[5];
#pragma unroll
    for (int t = 0; t < 5; t++) pf[t] = (float)r[t < n ? t : 0].prec;
    int c_tot[5] = {0, 0, 0, 0, 0}, c_alg[5] = {0, 0, 0, 0, 0};
    const int n_rows = y_max - y_min + 1;
    int G = 1;
    while (G < 32 && n_rows * G * 2 <= 32) G <<= 1;
    const int rows_per_step = 32 / G, sub = lane & (G - 1), rsel = lane / G;
    for (int yb = y_min; yb <= y_max; yb += rows_per_step) {
        const int y = yb + rsel;
        if (y > y_max) continue;
        int lo[5], hi[5];
        int x_lo = 0x7fffffff, x_hi = -1;
#pragma unroll
        for (int t = 0; t < 5; t++) {
            lo[t] = 1;
            hi[t] = 0;
            if (t < n_geom && y >= sp[t].y0 && y <= sp[t].y1) {
                lsd_span_row(sp[t], y, F.W, lo[t], hi[t]);
                if (hi[t] >= lo[t]) {
                    x_lo = min(x_lo, lo[t]);
                    x_hi = max(x_hi, hi[t]);
                    if (sub == 0) c_tot[t] += hi[t] - lo[t] + 1;
                }
            }
        }
        if (x_hi < 0) continue; /* no rectangle has pixels in this row */
        const float *row = F.angf + (size_t)y * F.W;
        for (int x = x_lo + sub; x <= x_hi; x += G) {
            const float deg = row[x];
            if (deg < 0.f) continue;
            float nf = fabsf(thetaf - deg * 0.017453292f);
            if (nf > 4.712389f) nf = fabsf(nf - 6.2831855f);
            if (same_geom) {
#pragma unroll
                for (int t = 0; t < 5; t++)
                    if (t < n) {
                        bool a = nf < pf[t] - 2e-4f;
                        if (!a && !(nf > pf[t] + 2e-4f)) a = lsd_aligned_deg(deg, theta, r[t].prec); /* borderline: the reference's doubles */
                        c_alg[t] += a ? 1 : 0;
                    }
            } else {
                bool a = nf < pf[0] - 2e-4f;
                if (!a && !(nf > pf[0] + 2e-4f)) a = lsd_aligned_deg(deg, theta, r[0].prec);
                if (a) {
#pragma unroll
                    for (int t = 0; t < 5; t++) c_alg[t] += (x >= lo[t] && x <= hi[t]) ? 1 : 0;
                }
            }
        }
    }
    if (same_geom) {
#pragma unroll
        for (int t = 1; t < 5; t++) c_tot[t] = c_tot[0];
    }
#pragma unroll
    for (int t = 0; t < 5; t++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            c_tot[t] += __shfl_xor_sync(FULL, c_tot[t], o);
            c_alg[t] += __shfl_xor_sync(FULL, c_alg[t], o);
        }
        if (t < n) {
            tot[t] = c_tot[t];
            alg[t] = c_alg[t];
        }
    }
}

/* the NFA of up to five rectangles of one rect_improve phase: the pixel counts in one scan (rows across lanes), then the binomial tails,
 * each by the whole warp (cs_nfa.cuh) */
__device__ __noinline__ void lsd_rect_nfa5(const LsdFrame &F, const LsdRect *r, int n, bool same_geom, double *v)
{
    int tot[5], alg[5];
    {
        LSD_PROF_T0();
        if (n == 1)
            lsd_rect_count(F, r[0], tot[0], alg[0]);
        else
            lsd_rect_count_multi(F, r, n, same_geom, tot, alg);
        LSD_PROF_ADD(3);
    }
    {
        LSD_PROF_T0();
        for (int t = 0; t < n; t++) v[t] = cs_nfa_warp(F.lgam, tot[t], alg[t], r[t].p, F.LOG_NT, true); /* the whole warp on one binomial tail */
        LSD_PROF_ADD(4);
    }
}

/* lsd.cpp:873-975: the rectangles of a phase do not depend on the NFA values of that phase, so they are evaluated together */
__device__ double lsd_rect_improve(const LsdFrame &F, LsdRect &rec)
{
    const double delta = 0.5, delta_2 = delta / 2.0, LOG_EPS = 0;
    LsdRect cand[5];
    double v[5];
    lsd_rect_nfa5(F, &rec, 1, true, v);
    double log_nfa = v[0];
    if (log_nfa > LOG_EPS) return log_nfa;
    LsdRect r = rec;
    for (int n = 0; n < 5; ++n) {
        r.p /= 2;
        r.prec = r.p * LSD_PI;
        cand[n] = r;
    }
    lsd_rect_nfa5(F, cand, 5, true, v);
    for (int n = 0; n < 5; ++n)
        if (v[n] > log_nfa) {
            log_nfa = v[n];
            rec = cand[n];
        }
    if (log_nfa > LOG_EPS) return log_nfa;
    for (int phase = 0; phase < 4; phase++) {
        r = rec;
        int m = 0;
        for (int n = 0; n < 5; ++n)
            if ((r.width - delta) >= 0.5) {
                if (phase == 0)
                    r.width -= delta;
                else if (phase == 1) {
                    r.x1 += -r.dy * delta_2;
                    r.y1 += r.dx * delta_2;
                    r.x2 += -r.dy * delta_2;
                    r.y2 += r.dx * delta_2;
                    r.width -= delta;
                } else if (phase == 2) {
                    r.x1 -= -r.dy * delta_2;
                    r.y1 -= r.dx * delta_2;
                    r.x2 -= -r.dy * delta_2;
                    r.y2 -= r.dx * delta_2;
                    r.width -= delta;
                } else {
                    r.p /= 2;
                    r.prec = r.p * LSD_PI;
                }
                cand[m++] = r;
            }
        if (m) {
            lsd_rect_nfa5(F, cand, m, phase == 3, v);
            for (int n = 0; n < m; ++n)
                if (v[n] > log_nfa) {
                    rec = cand[n];
                    log_nfa = v[n];
                }
        }
        if (phase < 3 && log_nfa > LOG_EPS) return log_nfa;
    }
    return log_nfa;
}

/* The first half of one seed (lsd.cpp:478-519): grow, rectangle, density refinement.  has_rect = 1 when a rectangle comes out that
 * rect_improve / the NFA test still have to judge -- which they can do later, in any order and in parallel: they read the level-line
 * angles only and never touch the `used` map. */
__device__ __noinline__ void lsd_region2rect_cold(const LsdFrame &F, const LsdReg &R, int base, int reg_size, double reg_angle, double prec, double p, LsdRect &rec)
{
    lsd_region2rect(F, R, base, reg_size, reg_angle, prec, p, rec);
}

__device__ void lsd_grow_candidate(const LsdFrame &F, const LsdReg &R, int s_addr, int min_reg_size, double prec, double p, int &n_all,
                                  int &has_rect, LsdRect &rec)
{
    /* region_grow -> region2rect -> density test, at most twice: the second round is refine()'s re-grow with the tolerance estimated from
     * the first region (lsd.cpp:786-832).  One loop, so that the two big inlined bodies exist once in the kernel (instruction cache). */
    const double DENSITY_TH = 0.7;
    const int lane = threadIdx.x & 31;
    int base = 0, reg_size = 0, seed = s_addr;
    double reg_angle = 0, tau = prec;
    has_rect = 0;
    n_all = 0;
    for (int pass = 0; pass < 2; pass++) {
        {
            LSD_PROF_T0();
            lsd_region_grow(F, R, base, seed, reg_size, reg_angle, tau);
            LSD_PROF_ADD(pass == 0 ? 0 : 2);
        }
        if (pass == 0 && lane == 0) {
            atomicAdd(&g_lsd_prof[5], 1ull);
            atomicAdd(&g_lsd_prof[7], (unsigned long long)reg_size);
        }
        n_all = base + reg_size;
        if (reg_size < (pass == 0 ? min_reg_size : 2)) return;
        {
            LSD_PROF_T0();
            lsd_region2rect(F, R, base, reg_size, reg_angle, prec, p, rec);
            LSD_PROF_ADD(1);
        }
        const double density = (double)reg_size / (lsd_dist(rec.x1, rec.y1, rec.x2, rec.y2) * rec.width);
        if (density >= DENSITY_TH) {
            has_rect = 1;
            return;
        }
        if (pass == 1) { /* still too sparse after the re-grow: shrink it around the seed (lsd.cpp:834-871) */
            if (lsd_reduce_region_radius(F, R, base, reg_size, reg_angle, prec, p, rec, density, DENSITY_TH) == 0) has_rect = 1;
            return;
        }
        /* refine(): tolerance from the angle spread near the seed, give the region back, grow again from the same seed */
        LSD_PROF_T0();
        const int a0 = R.get(base);
        const double xc = (double)(a0 - F.row_of(a0) * F.W), yc = (double)F.row_of(a0);
        const double ang_c = (double)F.angf[a0] * LSD_DEG2RAD;
        double sum = 0, s_sum = 0;
        int n = 0;
        for (int i0 = 0; i0 < reg_size; i0 += 32) {
            const int m = min(32, reg_size - i0);
            bool near = false;
            __syncwarp();
            if (lane < m) {
                const int addr = R.get(base + i0 + lane);
                const int ry = F.row_of(addr), rx = addr - ry * F.W;
                if (lsd_dist(xc, yc, (double)rx, (double)ry) < rec.width) {
                    near = true;
                    const double ang_d = lsd_angle_diff_signed((double)F.angf[addr] * LSD_DEG2RAD, ang_c);
                    F.stage[lane] = ang_d;
                    F.stage[32 + lane] = ang_d * ang_d;
                }
            }
            __syncwarp();
            unsigned todo = __ballot_sync(0xffffffffu, near);
            n += __popc(todo);
            while (todo) {
                const int j = __ffs(todo) - 1;
                todo &= todo - 1;
                sum += F.stage[j];
                s_sum += F.stage[32 + j];
            }
        }
        for (int i = lane; i < reg_size; i += 32) {
            const int addr = R.get(base + i);
            atomicAnd(F.ubits + (addr >> 5), ~(1u << (addr & 31)));
        }
        __syncwarp();
        const double mean_angle = sum / (double)n;
        tau = 2.0 * sqrt((s_sum - 2.0 * mean_angle * sum) / (double)n + mean_angle * mean_angle);
        seed = a0;
        base += reg_size;
        LSD_PROF_ADD(2);
    }
}

/* The second half (lsd.cpp:520-534): rect_improve, NFA test, back to image coordinates.  Returns whether the rectangle is a line. */
__device__ bool lsd_validate_rect(const LsdFrame &F, LsdRect &rec, double scale, float *line)
{
    const double LOG_EPS = 0;
    const double log_nfa = lsd_rect_improve(F, rec);
    if (log_nfa <= LOG_EPS) return false;
    rec.x1 += 0.5;
    rec.y1 += 0.5;
    rec.x2 += 0.5;
    rec.y2 += 0.5;
    rec.x1 /= scale;
    rec.y1 /= scale;
    rec.x2 /= scale;
    rec.y2 /= scale;
    line[0] = (float)rec.x1;
    line[1] = (float)rec.y1;
    line[2] = (float)rec.x2;
    line[3] = (float)rec.y2;
    return true;
}

/* checkLineExtremes + 10-px border rejection + length filter (LSDDetector.cpp:75-101,226-238; filter_lines): true = keep */
__device__ __forceinline__ bool lsd_keyline_filter(const float *raw, int img_w, int img_h, float line_length_thres, float *o)
{
    const float pre_boundary_thre = 10;
    float e[4] = {raw[0], raw[1], raw[2], raw[3]};
    if (e[0] < 0) e[0] = 0;
    if (e[0] >= img_w) e[0] = (float)img_w - 1.0f;
    if (e[2] < 0) e[2] = 0;
    if (e[2] >= img_w) e[2] = (float)img_w - 1.0f;
    if (e[1] < 0) e[1] = 0;
    if (e[1] >= img_h) e[1] = (float)img_h - 1.0f;
    if (e[3] < 0) e[3] = 0;
    if (e[3] >= img_h) e[3] = (float)img_h - 1.0f;
    const float sx = e[0], sy = e[1], ex = e[2], ey = e[3];
    if (((sx < pre_boundary_thre) && (ex < pre_boundary_thre)) || ((sx > img_w - pre_boundary_thre) && (ex > img_w - pre_boundary_thre)) ||
        ((sy < pre_boundary_thre) && (ey < pre_boundary_thre)) || ((sy > img_h - pre_boundary_thre) && (ey > img_h - pre_boundary_thre)))
        return false;
    const double ddx = (double)(e[0] - e[2]), ddy = (double)(e[1] - e[3]);
    const float line_length = (float)sqrt(ddx * ddx + ddy * ddy);
    if (!(line_length > line_length_thres)) return false;
    o[0] = sx;
    o[1] = sy;
    o[2] = ex;
    o[3] = ey;
    return true;
}

struct LsdGrowArgs {
    int W, H, img_w, img_h;
    uint4 *pix;
    const float *angf;
    const double *modgrad;
    const int32_t *list;
    const int32_t *list_len;
    uint32_t *st;
    int32_t *arena;
    int arena_cap;       /* ints per frame */
    int32_t *spill;      /* LSD_SPILL ints per warp (the first LSD_SCAP unused) */
    double LOG_NT;
    int min_reg_size;
    double prec, p, scale;
    float line_length_thres;
    float *raw;
    int32_t *n_raw;
    float *out;
    int32_t *n_out;
    int cap;
    LsdRect *cand;       /* cand_cap rectangles per frame, seed order: k_lsd_grow_seq -> k_lsd_validate */
    int32_t *n_cand;
    int cand_cap;
    int32_t *cand_line;  /* per candidate: {is a line, 4 floats} */
    struct LsdCandState *cand_state; /* per candidate: the state of the phase-split validation */
    int32_t *err;        /* bit 2: more candidates in a frame than cand_cap */
    const double *lgam;  /* log_gamma table (cs_nfa.cuh) */
    uint32_t *ubits;     /* (W * H + 31) / 32 words per frame: the used map of k_lsd_grow_seq */
    int32_t *redo;       /* per frame: 1 = the sequential kernel must redo this frame */
    int32_t *stats;      /* per frame: rounds, candidates processed, refused, invalidated (diagnostics) */
};

/* The order-dependent half of the seed loop, one warp per frame (see the file header). */
template <int kMinCtas>
__global__ void __launch_bounds__(32, kMinCtas) k_lsd_grow_seq(LsdGrowArgs A, int scap)
{
    const int f = blockIdx.x, lane = threadIdx.x;
    LSD_PROF_T0();
    const size_t npx = (size_t)A.W * A.H;
    LsdFrame F;
    F.W = A.W;
    F.H = A.H;
    F.pix = A.pix + f * npx;
    F.angf = A.angf + f * npx;
    F.modgrad = A.modgrad + f * npx;
    F.LOG_NT = A.LOG_NT;
    extern __shared__ uint32_t s_seq[];
    const int n_words = (int)((npx + 31) >> 5);
    F.ubits = A.ubits + (size_t)f * n_words;
    F.span = nullptr;
    F.stage = reinterpret_cast<double *>(s_seq + scap);
    F.lgam = A.lgam;
    F.wmagic = ((1ull << 40) + (unsigned long long)A.W - 1) / (unsigned long long)A.W;
    for (int i = lane; i < n_words; i += 32) F.ubits[i] = 0u;
    __syncwarp();
    LsdReg R;
    R.s = (int *)s_seq; /* the first LSD_SEQ_SCAP region entries in shared memory, the rest in the (otherwise unused) record arena */
    R.g = A.arena + (size_t)f * A.arena_cap;
    R.scap = scap;
    R.cap = A.arena_cap;
    const int n_px = (int)npx;
    int n_cand = 0;
    /* Seeds in RASTER order: flsd walks its coorlist vector by index (lsd.cpp:478-480), and ll_angle fills that vector in scan order; the
     * gradient-bin links it also builds (lsd.cpp:588-634) are never followed.  128 pixels per step, four independent loads per lane in
     * flight: pixels without a defined angle (which includes the last row and column) or already used are skipped by ballot.  The used
     * words are read when no region is in progress: a bit that is set then belongs to a finished region and stays set, and a pixel that
     * gets used after the fetch is caught by the re-check at its turn. */
    for (int i0 = 0; i0 < n_px; i0 += 128) {
        float a4[4];
        uint32_t u4[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int adx = i0 + 32 * j + lane;
            a4[j] = adx < n_px ? F.angf[adx] : -1.f;
            u4[j] = i0 + 32 * j < n_px ? __ldcg(F.ubits + ((i0 + 32 * j) >> 5)) : 0xffffffffu;
        }
#pragma unroll 1 /* the body holds the whole seed pipeline: once in the kernel (instruction cache) */
        for (int j = 0; j < 4; j++) {
        const int adx = i0 + 32 * j + lane;
        const float a_j = j == 0 ? a4[0] : j == 1 ? a4[1] : j == 2 ? a4[2] : a4[3];
        const uint32_t u_j = j == 0 ? u4[0] : j == 1 ? u4[1] : j == 2 ? u4[2] : u4[3];
        const bool seed = a_j >= 0.f && !((u_j >> lane) & 1u);
        unsigned todo = __ballot_sync(0xffffffffu, seed);
        while (todo) {
            const int sl = __ffs(todo) - 1;
            todo &= todo - 1;
            const int s_addr = __shfl_sync(0xffffffffu, adx, sl);
            if (F.used_bit(s_addr)) continue; /* used by a region grown since the ballot */
            int n_all = 0, has_rect = 0;
            LsdRect rec;
            lsd_grow_candidate(F, R, s_addr, A.min_reg_size, A.prec, A.p, n_all, has_rect, rec); /* arena_cap >= 2 W H: both passes fit */
            if (!has_rect) continue;
            /* the rectangle goes to k_lsd_validate (rect_improve + NFA never touch the used map): candidates in seed order */
            if (lane == 0 && n_cand < A.cand_cap) A.cand[(size_t)f * A.cand_cap + n_cand] = rec;
            n_cand++;
        }
        }
    }
    if (lane == 0) {
        A.n_cand[f] = n_cand;
        if (n_cand > A.cand_cap) atomicOr(A.err, 4);
    }
    LSD_PROF_ADD(6);
}

/* rect_improve + NFA of every candidate rectangle (lsd.cpp:520-534), one warp each: thousands of warps running the same code.  (A CTA of
 * five warps per candidate, one warp per rectangle of a rect_improve phase, was measured slower: 3.5 vs 2.45 ms per 256 frames.) */
template <int kMinCtas>
__global__ void __launch_bounds__(128, kMinCtas) k_lsd_validate(LsdGrowArgs A)
{
    const int f = blockIdx.y, lane = threadIdx.x & 31;
    const size_t npx = (size_t)A.W * A.H;
    LsdFrame F;
    F.W = A.W;
    F.H = A.H;
    F.pix = A.pix + f * npx;
    F.angf = A.angf + f * npx;
    F.modgrad = A.modgrad + f * npx;
    F.LOG_NT = A.LOG_NT;
    F.ubits = nullptr;
    __shared__ LsdSpan s_span[4][5];
    F.span = s_span[threadIdx.x >> 5];
    F.stage = nullptr;
    F.lgam = A.lgam;
    F.wmagic = 0;
    const int n = min(A.n_cand[f], A.cand_cap);
    const int wpb = blockDim.x >> 5;
    for (int c = blockIdx.x * wpb + (threadIdx.x >> 5); c < n; c += gridDim.x * wpb) {
        LsdRect rec = A.cand[(size_t)f * A.cand_cap + c];
        float line[4] = {0.f, 0.f, 0.f, 0.f};
        const bool ok = lsd_validate_rect(F, rec, A.scale, line);
        if (lane == 0) {
            int32_t *o = A.cand_line + ((size_t)f * A.cand_cap + c) * 5;
            o[0] = ok ? 1 : 0;
            for (int k = 0; k < 4; k++) o[1 + k] = __float_as_int(line[k]);
        }
    }
}

/* ---- the same, split by phase into small kernels ---------------------------------------------------------------------------------
 * k_lsd_validate is ~6.6 k instructions (the scans, the double-precision log / exp / pow of the binomial tail, the improvement logic), the
 * L1.5 instruction cache holds 2 k, and its warps sit in different places of that code: ncu shows 40 % of its issue stalls as
 * `no_instruction`.  rect_improve is six rounds of {count the pixels of up to five rectangles, turn the counts into NFA values, keep the
 * best} (lsd.cpp:873-975); run as twelve launches -- k_lsd_val_count(round), k_lsd_val_nfa(round) -- every warp of a launch is in the
 * same few hundred instructions.  State between launches: the candidate's current rectangle (A.cand, updated in place), its best
 * log_nfa and the counts of the round (A.cand_state). */
struct LsdCandState {
    double log_nfa;
    int32_t status; /* 0 = still being improved, 1 = decided (A.cand_line holds the verdict) */
    int32_t n;      /* rectangles counted in this round */
    int32_t tot[5], alg[5];
};

/* the rectangles round `round` evaluates, from the current one (lsd.cpp:873-975; round 0 = the rectangle itself) */
__device__ int lsd_improve_variants(const LsdRect &rec, int round, LsdRect *cand)
{
    const double delta = 0.5, delta_2 = delta / 2.0;
    LsdRect r = rec;
    if (round == 0) {
        cand[0] = r;
        return 1;
    }
    if (round == 1) {
        for (int n = 0; n < 5; ++n) {
            r.p /= 2;
            r.prec = r.p * LSD_PI;
            cand[n] = r;
        }
        return 5;
    }
    const int phase = round - 2;
    int m = 0;
    for (int n = 0; n < 5; ++n)
        if ((r.width - delta) >= 0.5) {
            if (phase == 0)
                r.width -= delta;
            else if (phase == 1) {
                r.x1 += -r.dy * delta_2;
                r.y1 += r.dx * delta_2;
                r.x2 += -r.dy * delta_2;
                r.y2 += r.dx * delta_2;
                r.width -= delta;
            } else if (phase == 2) {
                r.x1 -= -r.dy * delta_2;
                r.y1 -= r.dx * delta_2;
                r.x2 -= -r.dy * delta_2;
                r.y2 -= r.dx * delta_2;
                r.width -= delta;
            } else {
                r.p /= 2;
                r.prec = r.p * LSD_PI;
            }
            cand[m++] = r;
        }
    return m;
}

__global__ void __launch_bounds__(128, 5) k_lsd_val_count(LsdGrowArgs A, int round)
{
    const int f = blockIdx.y, lane = threadIdx.x & 31;
    const size_t npx = (size_t)A.W * A.H;
    LsdFrame F;
    F.W = A.W;
    F.H = A.H;
    F.pix = nullptr;
    F.angf = A.angf + f * npx;
    F.modgrad = nullptr;
    F.LOG_NT = A.LOG_NT;
    F.ubits = nullptr;
    __shared__ LsdSpan s_span[4][5];
    F.span = s_span[threadIdx.x >> 5];
    F.stage = nullptr;
    F.lgam = A.lgam;
    F.wmagic = 0;
    const int n = min(A.n_cand[f], A.cand_cap);
    const int wpb = blockDim.x >> 5;
    for (int c = blockIdx.x * wpb + (threadIdx.x >> 5); c < n; c += gridDim.x * wpb) {
        LsdCandState *S = A.cand_state + (size_t)f * A.cand_cap + c;
        if (round > 0 && S->status != 0) continue;
        const LsdRect rec = A.cand[(size_t)f * A.cand_cap + c];
        LsdRect cv[5];
        const int m = lsd_improve_variants(rec, round, cv);
        int tot[5] = {0, 0, 0, 0, 0}, alg[5] = {0, 0, 0, 0, 0};
        if (m == 1)
            lsd_rect_count(F, cv[0], tot[0], alg[0]);
        else if (m > 1)
            lsd_rect_count_multi(F, cv, m, round == 1 || round == 5, tot, alg);
        if (lane == 0) {
            S->n = m;
            for (int t = 0; t < 5; t++) {
                S->tot[t] = tot[t];
                S->alg[t] = alg[t];
            }
        }
    }
}

__global__ void __launch_bounds__(128, 5) k_lsd_val_nfa(LsdGrowArgs A, int round)
{
    const int f = blockIdx.y, lane = threadIdx.x & 31;
    const int n = min(A.n_cand[f], A.cand_cap);
    const int wpb = blockDim.x >> 5;
    for (int c = blockIdx.x * wpb + (threadIdx.x >> 5); c < n; c += gridDim.x * wpb) {
        LsdCandState *S = A.cand_state + (size_t)f * A.cand_cap + c;
        if (round > 0 && S->status != 0) continue;
        LsdRect rec = A.cand[(size_t)f * A.cand_cap + c];
        LsdRect cv[5];
        const int m = lsd_improve_variants(rec, round, cv);
        double log_nfa = round == 0 ? 0.0 : S->log_nfa;
        bool changed = false;
        for (int t = 0; t < m; t++) {
            const double v = cs_nfa_warp(A.lgam, S->tot[t], S->alg[t], cv[t].p, A.LOG_NT, true); /* the whole warp on one binomial tail */
            if (round == 0 || v > log_nfa) {
                log_nfa = v;
                rec = cv[t];
                changed = round > 0;
            }
        }
        const bool done = round == 5 || log_nfa > 0.0; /* lsd.cpp:887,903,921,938,955: leave as soon as the rectangle is meaningful */
        __syncwarp();
        if (lane == 0) {
            if (done) {
                int32_t *o = A.cand_line + ((size_t)f * A.cand_cap + c) * 5;
                const bool ok = log_nfa > 0.0;
                o[0] = ok ? 1 : 0;
                /* lsd.cpp:524-534: the half-pixel offset, back to the scale of the input image */
                o[1] = __float_as_int((float)((rec.x1 + 0.5) / A.scale));
                o[2] = __float_as_int((float)((rec.y1 + 0.5) / A.scale));
                o[3] = __float_as_int((float)((rec.x2 + 0.5) / A.scale));
                o[4] = __float_as_int((float)((rec.y2 + 0.5) / A.scale));
                S->status = 1;
            } else {
                if (changed) A.cand[(size_t)f * A.cand_cap + c] = rec;
                S->log_nfa = log_nfa;
                S->status = 0;
            }
        }
    }
}

/* kernels one LSD run launches: blur, resize, gradient, seed loop, validation (12 phase kernels, or the single kernel of the A/B variants), emit */
static int cs_lsd_launches_per_run()
{
    static const int v_val = getenv("CS_LSD_VAL_VARIANT") ? atoi(getenv("CS_LSD_VAL_VARIANT")) : 0;
    return 3 + 1 + (v_val == 0 ? 12 : 1) + 1;
}

/* the accepted candidates of a frame in seed order: raw segments, and those that pass the key-line filter */
__global__ void __launch_bounds__(256) k_lsd_emit(LsdGrowArgs A)
{
    __shared__ int s_warp_cnt[8], s_base_raw, s_base_out;
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const unsigned FULL = 0xffffffffu;
    const int n = min(A.n_cand[f], A.cand_cap);
    float *raw = A.raw + (size_t)f * A.cap * 4;
    float *out = A.out + (size_t)f * A.cap * 4;
    if (tid == 0) {
        s_base_raw = 0;
        s_base_out = 0;
    }
    __syncthreads();
    for (int b = 0; b < n; b += 256) {
        const int i = b + tid;
        bool has = false, kept = false;
        float ln[4], fo[4];
        if (i < n) {
            const int32_t *o = A.cand_line + ((size_t)f * A.cand_cap + i) * 5;
            if (o[0]) {
                has = true;
                for (int k = 0; k < 4; k++) ln[k] = __int_as_float(o[1 + k]);
                kept = lsd_keyline_filter(ln, A.img_w, A.img_h, A.line_length_thres, fo);
            }
        }
        const unsigned mh = __ballot_sync(FULL, has), mk = __ballot_sync(FULL, kept);
        if (lane == 0) s_warp_cnt[wid] = __popc(mh) | (__popc(mk) << 16);
        __syncthreads();
        int pre_r = s_base_raw, pre_o = s_base_out;
        for (int k = 0; k < wid; k++) {
            pre_r += s_warp_cnt[k] & 0xffff;
            pre_o += s_warp_cnt[k] >> 16;
        }
        if (has) {
            const int slot = pre_r + __popc(mh & ((1u << lane) - 1u));
            if (slot < A.cap)
                for (int k = 0; k < 4; k++) raw[4 * slot + k] = ln[k];
        }
        if (kept) {
            const int slot = pre_o + __popc(mk & ((1u << lane) - 1u));
            if (slot < A.cap)
                for (int k = 0; k < 4; k++) out[4 * slot + k] = fo[k];
        }
        __syncthreads();
        if (tid == 0) {
            int tr = 0, to = 0;
            for (int k = 0; k < 8; k++) {
                tr += s_warp_cnt[k] & 0xffff;
                to += s_warp_cnt[k] >> 16;
            }
            s_base_raw += tr;
            s_base_out += to;
        }
        __syncthreads();
    }
    if (tid == 0) {
        /* more candidate rectangles than the hand-off buffer holds: report it the way a segment overflow is reported (count > capacity) */
        const bool overflow = A.n_cand[f] > A.cand_cap;
        A.n_raw[f] = overflow ? A.cap + 1 : s_base_raw;
        A.n_out[f] = overflow ? A.cap + 1 : s_base_out;
    }
}

/* ---------------------------------------------------------------------------------------- host side */
struct Buf {
    void *p = nullptr;
    size_t cap = 0;
};

struct LsdState {
    Buf img, tmp, blur, scaled, modgrad, angf, pix, arena, raw, nraw, out, nout, redo, stats, lgam, ubits, cand, ncand, candline, candstate, err;
    bool lgam_filled = false;
    int last_frames = 0, last_W = 0, last_H = 0, cap = 0;
};

int ensure(cs_ctx *c, Buf &b, size_t bytes)
{
    if (bytes <= b.cap) return CS_OK;
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
    const size_t want = bytes + bytes / 16 + 256;
    if (cudaMalloc(&b.p, want) != cudaSuccess) return cs_ctx_fail(c, CS_ERR_CUDA, "cudaMalloc(%zu) failed in the line detector", want);
    b.cap = want;
    return CS_OK;
}

inline int grid_for(int64_t n) { return (int)std::min<int64_t>((n + 255) / 256, 148 * 32); }

/* raise a kernel's dynamic shared-memory limit when a launch needs more than it was last given (per device) */
#define CS_LSD_SET_SMEM(kernel, bytes)                                                                  \
    do {                                                                                                \
        static size_t set_[64] = {};                                                                    \
        const int dev_ = cs_ctx_device(c) & 63;                                                         \
        if ((bytes) > set_[dev_]) {                                                                     \
            cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes));    \
            set_[dev_] = (bytes);                                                                       \
        }                                                                                               \
    } while (0)

int lsd_run(cs_ctx *c, const uint8_t *imgs, bool imgs_on_device, int n_frames, int w, int h, int stride, int channels, float line_length_thres,
            int cap, LsdState &S)
{
    cudaStream_t st = cs_ctx_stream(c);
    const double SCALE = 0.8;
    const int W = (int)std::lrint(w * SCALE), H = (int)std::lrint(h * SCALE);
    if (W < 2 || H < 2) return cs_ctx_fail(c, CS_ERR_INVALID_ARG, "image too small for LSD");
    const size_t px = (size_t)n_frames * w * h, spx = (size_t)n_frames * W * H;
    const int arena_cap = (int)(((size_t)W * H * 2 + 64 + 3) & ~(size_t)3); /* ints per frame */
    int rc;
    const uint8_t *d_img = imgs;
    if (!imgs_on_device) {
        if ((rc = ensure(c, S.img, (size_t)n_frames * h * stride))) return rc;
        if (cudaMemcpyAsync(S.img.p, imgs, (size_t)n_frames * h * stride, cudaMemcpyHostToDevice, st) != cudaSuccess)
            return cs_ctx_fail(c, CS_ERR_CUDA, "H2D copy of frames failed");
        d_img = (const uint8_t *)S.img.p;
    }
    if ((rc = ensure(c, S.tmp, cs_ctx_seq_lines(c) ? px * 8 : 0)) || (rc = ensure(c, S.blur, px * 8)) || (rc = ensure(c, S.scaled, spx * 8)) ||
        (rc = ensure(c, S.modgrad, spx * 8)) || (rc = ensure(c, S.angf, spx * 4)) || (rc = ensure(c, S.pix, spx * 16)) ||
        (rc = ensure(c, S.arena, (size_t)n_frames * arena_cap * 4)) ||
        (rc = ensure(c, S.raw, (size_t)n_frames * cap * 16)) || (rc = ensure(c, S.nraw, (size_t)n_frames * 4)) ||
        (rc = ensure(c, S.out, (size_t)n_frames * cap * 16)) || (rc = ensure(c, S.nout, (size_t)n_frames * 4)) ||
        (rc = ensure(c, S.redo, (size_t)n_frames * 4)) || (rc = ensure(c, S.stats, (size_t)n_frames * 16)) ||
        (rc = ensure(c, S.lgam, (size_t)CS_LGAMMA_TABLE * 8)) || (rc = ensure(c, S.ubits, (size_t)n_frames * (((size_t)W * H + 31) / 32) * 4)) ||
        (rc = ensure(c, S.cand, (size_t)n_frames * LSD_CAND_CAP * sizeof(LsdRect))) || (rc = ensure(c, S.ncand, (size_t)n_frames * 4)) ||
        (rc = ensure(c, S.candline, (size_t)n_frames * LSD_CAND_CAP * 20)) || (rc = ensure(c, S.candstate, (size_t)n_frames * LSD_CAND_CAP * sizeof(LsdCandState))) || (rc = ensure(c, S.err, 16)))
        return rc;
    if (!S.lgam_filled) { /* log_gamma of the integers 1 .. CS_LGAMMA_TABLE - 1, host libm like the reference */
        std::vector<double> t(CS_LGAMMA_TABLE, 0.0);
        for (int i = 1; i < CS_LGAMMA_TABLE; i++) t[i] = cs_lgamma_host((double)i);
        if (cudaMemcpyAsync(S.lgam.p, t.data(), t.size() * 8, cudaMemcpyHostToDevice, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess)
            return cs_ctx_fail(c, CS_ERR_CUDA, "upload of the log_gamma table failed");
        S.lgam_filled = true;
    }
    /* host-side constants of flsd (lsd.cpp:445-447,468-469), evaluated with libm like the reference */
    const double ANG_TH = 22.5, QUANT = 2.0;
    const double prec = LSD_PI * ANG_TH / 180, p = ANG_TH / 180, rho = QUANT / std::sin(prec);
    const double LOG_NT = 5 * (std::log10((double)W) + std::log10((double)H)) / 2 + std::log10(11.0);
    const int min_reg_size = (int)(-LOG_NT / std::log10(p));

    cudaMemsetAsync(S.err.p, 0, 16, st);
    cudaMemsetAsync(S.stats.p, 0, (size_t)n_frames * 16, st);
    const dim3 g_src((w * h + 255) / 256, n_frames), g_dst((W * H + 255) / 256, n_frames);
    if (cs_ctx_seq_lines(c)) { /* A/B: the two-pass kernels */
        k_lsd_hblur<<<g_src, 256, 0, st>>>(d_img, n_frames, w, h, stride, channels, (double *)S.tmp.p);
        k_lsd_vblur<<<g_src, 256, 0, st>>>((const double *)S.tmp.p, n_frames, w, h, (double *)S.blur.p);
    } else
    {
        const dim3 g_tile((w + LSB_TW - 1) / LSB_TW, (h + LSB_TH - 1) / LSB_TH, n_frames);
        CUtensorMap tm;
        /* BGR frames whose rows are a multiple of 16 bytes (640 and 1280 wide are) can be staged by the copy engine */
        if (cs_ctx_use_tma(c) && channels == 3 && stride == 3 * w && cs_make_tmap_bytes(&tm, d_img, 3 * (int64_t)w, (int64_t)n_frames * h, stride, LSB_BOXW, LSB_TH + 6))
            k_lsd_blur<true><<<g_tile, 256, 0, st>>>(tm, d_img, w, h, stride, channels, (double *)S.blur.p, (int32_t *)S.err.p);
        else
            k_lsd_blur<false><<<g_tile, 256, 0, st>>>(tm, d_img, w, h, stride, channels, (double *)S.blur.p, (int32_t *)S.err.p);
    }
    k_lsd_resize<<<g_dst, 256, 0, st>>>((const double *)S.blur.p, n_frames, w, h, W, H, 1. / SCALE, (double *)S.scaled.p);
    k_lsd_grad<<<g_dst, 256, 0, st>>>((const double *)S.scaled.p, n_frames, W, H, rho, (double *)S.modgrad.p, (float *)S.angf.p,
                                                       (uint4 *)S.pix.p);
    LsdGrowArgs A;
    A.W = W;
    A.H = H;
    A.img_w = w;
    A.img_h = h;
    A.pix = (uint4 *)S.pix.p;
    A.angf = (const float *)S.angf.p;
    A.modgrad = (const double *)S.modgrad.p;
    A.list = nullptr;
    A.list_len = nullptr;
    A.st = nullptr;
    A.arena = (int32_t *)S.arena.p;
    A.arena_cap = arena_cap;
    A.spill = nullptr;
    A.LOG_NT = LOG_NT;
    A.min_reg_size = min_reg_size;
    A.prec = prec;
    A.p = p;
    A.scale = SCALE;
    A.line_length_thres = line_length_thres;
    A.raw = (float *)S.raw.p;
    A.n_raw = (int32_t *)S.nraw.p;
    A.out = (float *)S.out.p;
    A.n_out = (int32_t *)S.nout.p;
    A.cap = cap;
    A.lgam = (const double *)S.lgam.p;
    A.ubits = (uint32_t *)S.ubits.p;
    A.cand = (LsdRect *)S.cand.p;
    A.n_cand = (int32_t *)S.ncand.p;
    A.cand_cap = LSD_CAND_CAP;
    A.cand_line = (int32_t *)S.candline.p;
    A.cand_state = (LsdCandState *)S.candstate.p;
    A.err = (int32_t *)S.err.p;
    A.redo = (int32_t *)S.redo.p;
    A.stats = (int32_t *)S.stats.p;
    {
        static const int v_grow = getenv("CS_LSD_GROW_VARIANT") ? atoi(getenv("CS_LSD_GROW_VARIANT")) : 0;
        static const int v_val = getenv("CS_LSD_VAL_VARIANT") ? atoi(getenv("CS_LSD_VAL_VARIANT")) : 0;
        static const int scap = getenv("CS_LSD_SCAP") ? atoi(getenv("CS_LSD_SCAP")) : LSD_SEQ_SCAP;
        const size_t smem = (size_t)scap * 4 + 96 * 8; /* region list + the staging of the ordered sums (scap is even: 8-byte aligned) */
        if (v_grow == 1)
            k_lsd_grow_seq<32><<<n_frames, 32, smem, st>>>(A, scap);
        else if (v_grow == 2)
            k_lsd_grow_seq<25><<<n_frames, 32, smem, st>>>(A, scap);
        else
            k_lsd_grow_seq<21><<<n_frames, 32, smem, st>>>(A, scap);
        /* 256 warps per frame: a warp per candidate for all but the densest frames */
        if (v_val == 1)
            k_lsd_validate<8><<<dim3(64, n_frames), 128, 0, st>>>(A);
        else if (v_val == 2)
            k_lsd_validate<5><<<dim3(64, n_frames), 128, 0, st>>>(A);
        else
            for (int round = 0; round < 6; round++) {
                k_lsd_val_count<<<dim3(64, n_frames), 128, 0, st>>>(A, round);
                k_lsd_val_nfa<<<dim3(64, n_frames), 128, 0, st>>>(A, round);
            }
        k_lsd_emit<<<n_frames, 256, 0, st>>>(A);
    }
    cs_ctx_count_launches(c, cs_lsd_launches_per_run());
    if (cudaGetLastError() != cudaSuccess) return cs_ctx_fail(c, CS_ERR_CUDA, "LSD kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    S.last_frames = n_frames;
    S.last_W = W;
    S.last_H = H;
    S.cap = cap;
    return CS_OK;
}

LsdState *state_of(cs_ctx *c)
{
    void **slot = cs_ctx_lsd_slot(c);
    if (!*slot) *slot = new LsdState();
    return (LsdState *)*slot;
}

}  // namespace

/* device-to-device entry used by the online batch path: frames already in HBM, results stay in HBM */
int cs_lsd_run_device(cs_ctx *c, const uint8_t *d_imgs, int n_frames, int w, int h, int stride, int channels, float line_length_thres, int cap,
                      const float **d_lines, const int32_t **d_counts)
{
    LsdState *S = state_of(c);
    const int rc = lsd_run(c, d_imgs, true, n_frames, w, h, stride, channels, line_length_thres, cap, *S);
    if (rc) return rc;
    *d_lines = (const float *)S->out.p;
    *d_counts = (const int32_t *)S->nout.p;
    return CS_OK;
}

/* host frames in, results stay in HBM (the descriptor path, cs_lbd.cu) */
int cs_lsd_run_host(cs_ctx *c, const uint8_t *imgs, int n_frames, int w, int h, int stride, int channels, float line_length_thres, int cap,
                    const float **d_lines, const int32_t **d_counts, const uint8_t **d_frames)
{
    LsdState *S = state_of(c);
    const int rc = lsd_run(c, imgs, false, n_frames, w, h, stride, channels, line_length_thres, cap, *S);
    if (rc) return rc;
    *d_lines = (const float *)S->out.p;
    *d_counts = (const int32_t *)S->nout.p;
    *d_frames = (const uint8_t *)S->img.p; /* the frames as uploaded (same stride), still in HBM */
    return CS_OK;
}

void cs_lsd_destroy(void *state)
{
    LsdState *S = (LsdState *)state;
    Buf *all[] = {&S->img, &S->tmp, &S->blur, &S->scaled, &S->modgrad, &S->angf, &S->pix, &S->arena, &S->raw, &S->nraw, &S->out, &S->nout, &S->redo, &S->stats, &S->lgam, &S->ubits, &S->cand, &S->ncand, &S->candline, &S->candstate, &S->err};
    for (Buf *b : all)
        if (b->p) cudaFree(b->p);
    delete S;
}

extern "C" {

int cs_detect_lines_batch(cs_ctx *c, const uint8_t *imgs, int n_frames, int width, int height, int stride, int channels,
                          const cs_line_params *params, float *lines_xyxy, int32_t max_lines_per_frame, int32_t *n_lines)
{
    if (!c) return CS_ERR_INVALID_ARG;
    if (!imgs || !params || !lines_xyxy || !n_lines || n_frames <= 0 || width <= 0 || height <= 0 || max_lines_per_frame <= 0)
        return cs_ctx_fail(c, CS_ERR_INVALID_ARG, "null or empty argument");
    if (channels != 1 && channels != 3) return cs_ctx_fail(c, CS_ERR_INVALID_ARG, "channels must be 1 or 3"); /* LSDDetector.cpp:163-164 throws on depth != 0 */
    if (stride < width * channels) return cs_ctx_fail(c, CS_ERR_INVALID_ARG, "stride smaller than a row");
    /* More octaves change nothing here: filter_lines keeps octave 0 only (line_lbd_allclass.cpp:200-207) and octave 0 is detected first and
     * independently of the others in both detectors, so the matrix is the one-octave one (checked against the compiled reference with 2 and
     * 3 octaves, tests/test_oracle_ref_octaves.py). */
    if (params->numoctaves < 1) return cs_ctx_fail(c, CS_ERR_INVALID_ARG, "numoctaves must be at least 1");
    cudaSetDevice(cs_ctx_device(c));
    const float *d_out = nullptr;
    const int32_t *d_nout = nullptr;
    int rc;
    if (params->use_LSD) {
        LsdState *S = state_of(c);
        rc = lsd_run(c, imgs, false, n_frames, width, height, stride, channels, params->line_length_thres, max_lines_per_frame, *S);
        d_out = (const float *)S->out.p;
        d_nout = (const int32_t *)S->nout.p;
    } else
        rc = cs_edl_run(c, imgs, false, n_frames, width, height, stride, channels, params->line_length_thres, max_lines_per_frame, &d_out, &d_nout);
    if (rc) return rc;
    cudaStream_t st = cs_ctx_stream(c);
    std::vector<int32_t> cnt(n_frames);
    if (cudaMemcpyAsync(cnt.data(), d_nout, (size_t)n_frames * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaMemcpyAsync(lines_xyxy, d_out, (size_t)n_frames * max_lines_per_frame * 16, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess)
        return cs_ctx_fail(c, CS_ERR_CUDA, "line result copy failed: %s", cudaGetErrorString(cudaGetLastError()));
    for (int f = 0; f < n_frames; f++) {
        if (cnt[f] > max_lines_per_frame) return cs_ctx_fail(c, CS_ERR_CAPACITY, "frame %d: %d segments exceed max_lines_per_frame", f, cnt[f]);
        n_lines[f] = cnt[f];
    }
    return CS_OK;
}

int cs_detect_lines(cs_ctx *c, const uint8_t *img, int width, int height, int stride, int channels, const cs_line_params *params,
                    float *lines_xyxy, int32_t *n_inout)
{
    if (!c) return CS_ERR_INVALID_ARG;
    if (!n_inout || *n_inout <= 0) return cs_ctx_fail(c, CS_ERR_INVALID_ARG, "n_inout must give the capacity of lines_xyxy");
    int32_t n = 0;
    const int rc = cs_detect_lines_batch(c, img, 1, width, height, stride, channels, params, lines_xyxy, *n_inout, &n);
    if (rc == CS_OK) *n_inout = n;
    return rc;
}

/* inspection of the last run's intermediate images of one frame (tests): any pointer may be NULL */
int cs_debug_lsd(cs_ctx *c, int frame, int32_t *scaled_wh, double *scaled, double *modgrad, double *angles, int32_t *list, int32_t *list_len,
                 float *raw_lines, int32_t *n_raw, int cap_raw)
{
    if (!c) return CS_ERR_INVALID_ARG;
    LsdState *S = state_of(c);
    if (frame < 0 || frame >= S->last_frames) return cs_ctx_fail(c, CS_ERR_INVALID_ARG, "bad frame index");
    cudaSetDevice(cs_ctx_device(c));
    cudaStreamSynchronize(cs_ctx_stream(c));
    const size_t npx = (size_t)S->last_W * S->last_H;
    if (scaled_wh) {
        scaled_wh[0] = S->last_W;
        scaled_wh[1] = S->last_H;
    }
    if (scaled) cudaMemcpy(scaled, (double *)S->scaled.p + frame * npx, npx * 8, cudaMemcpyDeviceToHost);
    if (modgrad) cudaMemcpy(modgrad, (double *)S->modgrad.p + frame * npx, npx * 8, cudaMemcpyDeviceToHost);
    if (angles) {
        /* the level-line angle map as the reference holds it: (double)fastAtan2 * DEG2RAD, NOTDEF = -1024 (the product is IEEE-exact on either side) */
        std::vector<float> deg(npx);
        cudaMemcpy(deg.data(), (float *)S->angf.p + frame * npx, npx * 4, cudaMemcpyDeviceToHost);
        for (size_t i = 0; i < npx; i++) angles[i] = deg[i] < 0.f ? LSD_NOTDEF : (double)deg[i] * LSD_DEG2RAD;
    }
    if (list || list_len) {
        /* the order seeds are visited in: raster order over the pixels with a defined angle (lsd.cpp:478-481) */
        std::vector<float> deg(npx);
        cudaMemcpy(deg.data(), (float *)S->angf.p + frame * npx, npx * 4, cudaMemcpyDeviceToHost);
        int32_t ll = 0;
        for (size_t i = 0; i < npx; i++)
            if (deg[i] >= 0.f) {
                if (list) list[ll] = (int32_t)i;
                ll++;
            }
        if (list_len) *list_len = ll;
    }
    int32_t nr = 0;
    cudaMemcpy(&nr, (int32_t *)S->nraw.p + frame, 4, cudaMemcpyDeviceToHost);
    if (n_raw) *n_raw = nr;
    if (raw_lines) cudaMemcpy(raw_lines, (float *)S->raw.p + (size_t)frame * S->cap * 4, (size_t)std::min(nr, std::min(cap_raw, S->cap)) * 16, cudaMemcpyDeviceToHost);
    return cudaGetLastError() == cudaSuccess ? CS_OK : cs_ctx_fail(c, CS_ERR_CUDA, "debug copy failed");
}

/* cycle counters of the seed loop's phases summed over every warp since the last reset (diagnostics): {region_grow, region2rect, refine,
 * rectangle pixel counts, binomial tails, candidates grown, whole kernel (per CTA), unused} */
int cs_debug_lsd_prof(cs_ctx *c, uint64_t *out16, int reset)
{
    if (!c) return CS_ERR_INVALID_ARG;
    cudaSetDevice(cs_ctx_device(c));
    cudaStreamSynchronize(cs_ctx_stream(c));
    if (out16) cudaMemcpyFromSymbol(out16, g_lsd_prof, 128);
    if (reset) {
        const unsigned long long z[16] = {0};
        cudaMemcpyToSymbol(g_lsd_prof, z, 128);
    }
    return cudaGetLastError() == cudaSuccess ? CS_OK : cs_ctx_fail(c, CS_ERR_CUDA, "debug copy failed");
}

/* diagnostics of the speculative seed loop of the last run: per frame {rounds, candidates processed, refused or lost, re-grown after an
 * override} and whether the sequential kernel redid the frame */
int cs_debug_lsd_stats(cs_ctx *c, int32_t *stats4, int32_t *redo, int n_frames)
{
    if (!c) return CS_ERR_INVALID_ARG;
    LsdState *S = state_of(c);
    if (n_frames > S->last_frames) return cs_ctx_fail(c, CS_ERR_INVALID_ARG, "more frames than the last run held");
    cudaSetDevice(cs_ctx_device(c));
    cudaStreamSynchronize(cs_ctx_stream(c));
    if (stats4) cudaMemcpy(stats4, S->stats.p, (size_t)n_frames * 16, cudaMemcpyDeviceToHost);
    if (redo)
        for (int f = 0; f < n_frames; f++) redo[f] = 1; /* every frame goes through k_lsd_grow_seq (there is no speculative kernel any more) */
    return cudaGetLastError() == cudaSuccess ? CS_OK : cs_ctx_fail(c, CS_ERR_CUDA, "debug copy failed");
}
}
