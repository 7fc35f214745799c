/*
 * cs_lsd.cu -- line_lbd_detect::detect_filter_lines, LSD flavour, for sm_100a (kernel group (i) of the north star).
 *
 * Replaces, for one octave (the only one filter_lines keeps, line_lbd/class/line_lbd_allclass.cpp:200-207):
 *   LSDDetector::detectImpl                line_lbd/libs/LSDDetector.cpp:153-256
 *   LineSegmentDetectorImpl::flsd & co.    line_lbd/libs/lsd.cpp:440-1154   (LSD_REFINE_ADV, default parameters)
 *   filter_lines + keylines_to_mat         line_lbd/class/line_lbd_allclass.cpp:26-36,200-221
 *
 * Streaming stages (one thread per pixel, FP64, evaluation order of OpenCV's C paths, -fmad=false):
 *   k_lsd_hblur   cvtColor + horizontal 7-tap Gaussian (sigma 0.6/0.8)      lsd.cpp:452-457
 *   k_lsd_vblur   vertical 7-tap                                          lsd.cpp:457
 *   k_lsd_resize  cv::resize(x0.8, INTER_LINEAR) on doubles               lsd.cpp:459
 *   k_lsd_grad    2x2 gradient, modulus, fastAtan2 angle, max modulus     lsd.cpp:562-586
 *   k_lsd_hist / k_lsd_scan / k_lsd_scatter   the 1024-bin pseudo-ordering as a STABLE counting sort (bins descending,
 *                 raster order inside a bin == the reference's linked lists)  lsd.cpp:588-634
 * Sequential stage:
 *   k_lsd_grow    the seed loop (region_grow -> region2rect -> refine -> rect_improve/NFA), ONE WARP PER FRAME.  The order in
 *                 which seeds claim pixels defines the result, so a frame is inherently serial; the warp parallelises what is
 *                 order-free inside it (the 3x3 neighbour tests of a region point, list scanning, rectangle pixel counts,
 *                 min/max extents) and keeps every floating-point accumulation in the reference's order.  Throughput comes from
 *                 batching frames (thousands of warps resident per GPU).
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

#define LSD_PI 3.1415926535897932384626433832795
#define LSD_NOTDEF (-1024.0)
#define LSD_DEG2RAD (LSD_PI / 180)
#define LSD_3_2_PI ((3 * LSD_PI) / 2)
#define LSD_2PI (2 * LSD_PI)
#define LSD_LN10 2.30258509299404568402
#define LSD_NBINS 1024
#define LSD_CHUNK_ROWS 8

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
    const int64_t total = (int64_t)n_frames * w * h;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
        const int64_t f = p / ((int64_t)w * h);
        const int r = (int)(p - f * (int64_t)w * h);
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
    const int64_t total = (int64_t)n_frames * w * h;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
        const int64_t f = p / ((int64_t)w * h);
        const int r = (int)(p - f * (int64_t)w * h);
        const int y = r / w, x = r - y * w;
        const double *base = tmp + (size_t)f * w * h;
        double s = c_gauss7[3] * base[(size_t)y * w + x];
#pragma unroll
        for (int k = 1; k <= 3; k++) s += c_gauss7[3 + k] * (base[(size_t)reflect101(y + k, h) * w + x] + base[(size_t)reflect101(y - k, h) * w + x]);
        blur[p] = s;
    }
}

__global__ void __launch_bounds__(256) k_lsd_resize(const double *__restrict__ blur, int n_frames, int sw, int sh, int dw, int dh, double inv_scale,
                                                    double *__restrict__ scaled)
{
    const int64_t total = (int64_t)n_frames * dw * dh;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
        const int64_t f = p / ((int64_t)dw * dh);
        const int r = (int)(p - f * (int64_t)dw * dh);
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

__global__ void __launch_bounds__(256) k_lsd_grad(const double *__restrict__ scaled, int n_frames, int W, int H, double threshold,
                                                  double *__restrict__ modgrad, double *__restrict__ angles, float2 *__restrict__ cs_angle,
                                                  unsigned long long *__restrict__ max_bits)
{
    const int64_t total = (int64_t)n_frames * W * H;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
        const int64_t f = p / ((int64_t)W * H);
        const int addr = (int)(p - f * (int64_t)W * H);
        const int y = addr / W, x = addr - y * W;
        const double *im = scaled + (size_t)f * W * H;
        double norm = 0, ang = LSD_NOTDEF;
        if (x < W - 1 && y < H - 1) {
            const double DA = im[addr + W + 1] - im[addr];
            const double BC = im[addr + 1] - im[addr + W];
            const double gx = DA + BC, gy = DA - BC;
            norm = sqrt((gx * gx + gy * gy) / 4);
            if (!(norm <= threshold)) {
                ang = (double)fast_atan2((float)gx, (float)(-gy)) * LSD_DEG2RAD;
                atomicMax(max_bits + f, (unsigned long long)__double_as_longlong(norm)); /* positive doubles order like integers */
            }
        }
        modgrad[p] = norm;
        angles[p] = ang;
        /* (cos, sin) of float(angle), each the correctly rounded float: what region_grow accumulates (lsd.cpp:680-681) */
        float2 cs = make_float2(0.f, 0.f);
        if (ang != LSD_NOTDEF) {
            const double af = (double)(float)ang;
            cs.x = (float)cos(af);
            cs.y = (float)sin(af);
        }
        cs_angle[p] = cs;
    }
}

__device__ __forceinline__ double bin_coef_of(unsigned long long max_bits)
{
    const double max_grad = max_bits ? __longlong_as_double((long long)max_bits) : -1.0;
    return (max_grad > 0) ? (double)(LSD_NBINS - 1) / max_grad : 0.0;
}

/* per (frame, chunk of rows) histogram of gradient bins */
__global__ void __launch_bounds__(256) k_lsd_hist(const double *__restrict__ modgrad, const double *__restrict__ angles, int W, int H, int n_chunks,
                                                  const unsigned long long *__restrict__ max_bits, int32_t *__restrict__ cnt)
{
    __shared__ int s_h[LSD_NBINS];
    const int f = blockIdx.x / n_chunks, ch = blockIdx.x - f * n_chunks;
    for (int i = threadIdx.x; i < LSD_NBINS; i += 256) s_h[i] = 0;
    __syncthreads();
    const double coef = bin_coef_of(max_bits[f]);
    const int y0 = ch * LSD_CHUNK_ROWS, y1 = min(y0 + LSD_CHUNK_ROWS, H - 1);
    const double *mg = modgrad + (size_t)f * W * H;
    const double *an = angles + (size_t)f * W * H;
    const int npx = (y1 - y0) * (W - 1);
    for (int i = threadIdx.x; i < npx; i += 256) {
        const int y = y0 + i / (W - 1), x = i % (W - 1);
        if (an[(size_t)y * W + x] != LSD_NOTDEF) atomicAdd(&s_h[(int)(mg[(size_t)y * W + x] * coef)], 1);
    }
    __syncthreads();
    int32_t *o = cnt + ((size_t)f * n_chunks + ch) * LSD_NBINS;
    for (int i = threadIdx.x; i < LSD_NBINS; i += 256) o[i] = s_h[i];
}

/* per frame: turn the per-chunk counts into absolute list offsets (bins descending, chunks ascending) */
__global__ void __launch_bounds__(LSD_NBINS) k_lsd_scan(int n_chunks, int32_t *__restrict__ cnt, int32_t *__restrict__ list_len)
{
    __shared__ int s_tot[LSD_NBINS];
    const int f = blockIdx.x, b = threadIdx.x;
    int32_t *c = cnt + (size_t)f * n_chunks * LSD_NBINS;
    int run = 0;
    for (int ch = 0; ch < n_chunks; ch++) {
        const int v = c[(size_t)ch * LSD_NBINS + b];
        c[(size_t)ch * LSD_NBINS + b] = run;
        run += v;
    }
    s_tot[b] = run;
    __syncthreads();
    /* start[b] = sum of totals of the higher bins: inclusive scan over the reversed order */
    const int rb = LSD_NBINS - 1 - b; /* rank in descending order */
    for (int d = 1; d < LSD_NBINS; d <<= 1) {
        int v = 0;
        if (rb >= d) v = s_tot[b + d]; /* element d places earlier in descending order == bin b + d */
        __syncthreads();
        s_tot[b] += v;
        __syncthreads();
    }
    const int start = s_tot[b] - run;
    if (b == 0) list_len[f] = s_tot[0];
    for (int ch = 0; ch < n_chunks; ch++) c[(size_t)ch * LSD_NBINS + b] += start;
}

/* stable scatter: one warp walks its chunk in raster order, 32 pixels per step */
__global__ void __launch_bounds__(32) k_lsd_scatter(const double *__restrict__ modgrad, const double *__restrict__ angles, int W, int H, int n_chunks,
                                                    const unsigned long long *__restrict__ max_bits, const int32_t *__restrict__ cnt,
                                                    int32_t *__restrict__ list)
{
    __shared__ int s_c[LSD_NBINS];
    const int f = blockIdx.x / n_chunks, ch = blockIdx.x - f * n_chunks;
    const int lane = threadIdx.x;
    const int32_t *base = cnt + ((size_t)f * n_chunks + ch) * LSD_NBINS;
    for (int i = lane; i < LSD_NBINS; i += 32) s_c[i] = base[i];
    __syncwarp();
    const double coef = bin_coef_of(max_bits[f]);
    const int y0 = ch * LSD_CHUNK_ROWS, y1 = min(y0 + LSD_CHUNK_ROWS, H - 1);
    const double *mg = modgrad + (size_t)f * W * H;
    const double *an = angles + (size_t)f * W * H;
    int32_t *out = list + (size_t)f * W * H;
    const int npx = (y1 - y0) * (W - 1);
    for (int i0 = 0; i0 < npx; i0 += 32) {
        const int i = i0 + lane;
        bool ok = i < npx;
        int addr = 0, bin = -1 - lane; /* inactive lanes get unique negative keys */
        if (ok) {
            const int y = y0 + i / (W - 1), x = i % (W - 1);
            addr = y * W + x;
            ok = an[addr] != LSD_NOTDEF;
            if (ok) bin = (int)(mg[addr] * coef);
        }
        const unsigned m = __match_any_sync(0xffffffffu, bin);
        const int rank = __popc(m & ((1u << lane) - 1u));
        const int leader = __ffs(m) - 1;
        int old = 0;
        if (ok && lane == leader) {
            old = s_c[bin];
            s_c[bin] = old + __popc(m);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        if (ok) out[old + rank] = addr;
        __syncwarp();
    }
}

/* ---------------------------------------------------------------------------------------- the sequential stage */
struct LsdFrame {
    int W, H;
    const double *angles;
    const double *modgrad;
    const float2 *cs_angle;
    uint8_t *used;
    int32_t *reg;
    double LOG_NT;
};

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
/* lsd.cpp:1138-1154 on an angle value already loaded */
__device__ __forceinline__ bool lsd_aligned_val(double a, double theta, double prec)
{
    if (a == LSD_NOTDEF) return false;
    double n_theta = theta - a;
    if (n_theta < 0) n_theta = -n_theta;
    if (n_theta > LSD_3_2_PI) {
        n_theta -= LSD_2PI;
        if (n_theta < 0) n_theta = -n_theta;
    }
    return n_theta <= prec;
}

__device__ __forceinline__ double lsd_log_gamma(double x)
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

/* lsd.cpp:1100-1136 (warp-uniform scalar code) */
__device__ double lsd_nfa(int n, int k, double p, double LOG_NT)
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

/* lsd.cpp:637-688.  Neighbour tests of one region point run on lanes 0..8; additions stay in the reference's order. */
__device__ void lsd_region_grow(const LsdFrame &F, int s_addr, int &reg_size, double &reg_angle, double prec)
{
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    reg_size = 1;
    reg_angle = F.angles[s_addr];
    float sumdx = (float)cos(reg_angle);
    float sumdy = (float)sin(reg_angle);
    if (lane == 0) {
        F.reg[0] = s_addr;
        F.used[s_addr] = 1;
    }
    __syncwarp();
    /* Three region points per round: lanes 9g..9g+8 fetch the 3x3 neighbourhood of point i+g (used flag, angle, cos/sin) in one
     * go, then the points are consumed strictly in order.  A neighbour claimed while an earlier point of the round is consumed is
     * struck from the later groups, so the sequence of additions is exactly the reference's. */
    const int grp = lane / 9, kk = lane - grp * 9;
    const int ky = kk / 3 - 1, kx = kk - (kk / 3) * 3 - 1; /* (yy, xx) in the reference's loop order */
    for (int i = 0; i < reg_size;) {
        const int navail = min(3, reg_size - i);
        bool cand = false;
        int c_addr = -1;
        double a = LSD_NOTDEF;
        float2 csf = make_float2(0.f, 0.f);
        if (grp < navail) {
            const int pa = F.reg[i + grp];
            const int py = pa / F.W, px = pa - py * F.W;
            const int yy = py + ky, xx = px + kx;
            if (yy >= 0 && yy < F.H && xx >= 0 && xx < F.W) {
                c_addr = yy * F.W + xx;
                if (F.used[c_addr] != 1) {
                    a = F.angles[c_addr];
                    cand = (a != LSD_NOTDEF);
                    if (cand) csf = F.cs_angle[c_addr];
                }
            }
        }
        for (int g = 0; g < navail; g++) {
            unsigned pending = __ballot_sync(FULL, cand && grp == g);
            while (pending) {
                const unsigned ok = __ballot_sync(FULL, cand && ((pending >> lane) & 1u) && lsd_aligned_val(a, reg_angle, prec));
                if (!ok) break;
                const int fl = __ffs(ok) - 1;
                const int addrf = __shfl_sync(FULL, c_addr, fl);
                const float cx = __shfl_sync(FULL, csf.x, fl), cy = __shfl_sync(FULL, csf.y, fl);
                if (lane == 0) {
                    F.used[addrf] = 1;
                    F.reg[reg_size] = addrf;
                }
                ++reg_size;
                /* cos(float(angle)), sin(float(angle)): precomputed per pixel by k_lsd_grad (pinned to the correctly rounded float) */
                sumdx += cx;
                sumdy += cy;
                reg_angle = (double)fast_atan2(sumdy, sumdx) * LSD_DEG2RAD;
                pending &= ~((2u << fl) - 1u);
                if (c_addr == addrf) cand = false; /* the same pixel seen from a later point of this round */
            }
        }
        i += navail;
        __syncwarp();
    }
}

/* ordered accumulation helper: lanes fetch 32 region points at once, every lane then replays them in order */
#define LSD_FOR_REGION_ORDERED(F, reg_size, ...)                                    \
    for (int i0__ = 0; i0__ < (reg_size); i0__ += 32) {                             \
        const int n__ = min(32, (reg_size)-i0__);                                   \
        int my_addr__ = 0;                                                          \
        double my_w__ = 0, my_a__ = 0;                                              \
        if (lane < n__) {                                                           \
            my_addr__ = (F).reg[i0__ + lane];                                       \
            my_w__ = (F).modgrad[my_addr__];                                        \
            my_a__ = (F).angles[my_addr__];                                         \
        }                                                                           \
        for (int j__ = 0; j__ < n__; j__++) {                                       \
            const int addr = __shfl_sync(0xffffffffu, my_addr__, j__);              \
            const double weight = __shfl_sync(0xffffffffu, my_w__, j__);            \
            const double pangle = __shfl_sync(0xffffffffu, my_a__, j__);            \
            const int ry = addr / (F).W, rx = addr - ry * (F).W;                    \
            (void)weight;                                                           \
            (void)pangle;                                                           \
            __VA_ARGS__                                                             \
        }                                                                           \
    }

/* lsd.cpp:690-784 */
__device__ void lsd_region2rect(const LsdFrame &F, int reg_size, double reg_angle, double prec, double p, LsdRect &rec)
{
    const int lane = threadIdx.x & 31;
    double x = 0, y = 0, sum = 0;
    LSD_FOR_REGION_ORDERED(F, reg_size, {
        x += (double)rx * weight;
        y += (double)ry * weight;
        sum += weight;
    })
    x /= sum;
    y /= sum;
    /* get_theta */
    double Ixx = 0.0, Iyy = 0.0, Ixy = 0.0;
    LSD_FOR_REGION_ORDERED(F, reg_size, {
        const double ddx = (double)rx - x, ddy = (double)ry - y;
        Ixx += ddy * ddy * weight;
        Iyy += ddx * ddx * weight;
        Ixy -= ddx * ddy * weight;
    })
    const double lambda = 0.5 * (Ixx + Iyy - sqrt((Ixx - Iyy) * (Ixx - Iyy) + 4.0 * Ixy * Ixy));
    double theta = (fabs(Ixx) > fabs(Iyy)) ? (double)fast_atan2((float)(lambda - Ixx), (float)Ixy) : (double)fast_atan2((float)Ixy, (float)(lambda - Iyy));
    theta *= LSD_DEG2RAD;
    if (fabs(lsd_angle_diff_signed(theta, reg_angle)) > prec) theta += LSD_PI;
    const double dx = cos(theta), dy = sin(theta);
    /* extents: min / max are order-free, so lanes split the region */
    double l_min = 0, l_max = 0, w_min = 0, w_max = 0;
    for (int i = lane; i < reg_size; i += 32) {
        const int addr = F.reg[i];
        const int ry = addr / F.W, rx = addr - ry * F.W;
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

/* lsd.cpp:834-871 (lane 0 replays the reference's in-place compaction; it fixes the order later sums run in) */
__device__ bool lsd_reduce_region_radius(const LsdFrame &F, int &reg_size, double reg_angle, double prec, double p, LsdRect &rec, double density,
                                         double density_th)
{
    const int lane = threadIdx.x & 31;
    const int a0 = F.reg[0];
    const double xc = (double)(a0 % F.W), yc = (double)(a0 / F.W);
    const double radSq1 = lsd_dist_sq(xc, yc, rec.x1, rec.y1), radSq2 = lsd_dist_sq(xc, yc, rec.x2, rec.y2);
    double radSq = radSq1 > radSq2 ? radSq1 : radSq2;
    while (density < density_th) {
        radSq *= 0.75 * 0.75;
        int rs = reg_size;
        if (lane == 0) {
            for (int i = 0; i < rs; ++i) {
                const int addr = F.reg[i];
                if (lsd_dist_sq(xc, yc, (double)(addr % F.W), (double)(addr / F.W)) > radSq) {
                    F.used[addr] = 0;
                    const int last = F.reg[rs - 1];
                    F.reg[rs - 1] = addr;
                    F.reg[i] = last;
                    --rs;
                    --i;
                }
            }
        }
        reg_size = __shfl_sync(0xffffffffu, rs, 0);
        __syncwarp();
        if (reg_size < 2) return false;
        lsd_region2rect(F, reg_size, reg_angle, prec, p, rec);
        density = (double)reg_size / (lsd_dist(rec.x1, rec.y1, rec.x2, rec.y2) * rec.width);
    }
    return true;
}

/* lsd.cpp:786-832 */
__device__ bool lsd_refine(const LsdFrame &F, int &reg_size, double reg_angle, double prec, double p, LsdRect &rec, double density_th)
{
    const int lane = threadIdx.x & 31;
    double density = (double)reg_size / (lsd_dist(rec.x1, rec.y1, rec.x2, rec.y2) * rec.width);
    if (density >= density_th) return true;
    const int a0 = F.reg[0];
    const double xc = (double)(a0 % F.W), yc = (double)(a0 / F.W);
    const double ang_c = F.angles[a0];
    double sum = 0, s_sum = 0;
    int n = 0;
    LSD_FOR_REGION_ORDERED(F, reg_size, {
        if (lsd_dist(xc, yc, (double)rx, (double)ry) < rec.width) {
            const double ang_d = lsd_angle_diff_signed(pangle, ang_c);
            sum += ang_d;
            s_sum += ang_d * ang_d;
            ++n;
        }
    })
    for (int i = lane; i < reg_size; i += 32) F.used[F.reg[i]] = 0;
    __syncwarp();
    const double mean_angle = sum / (double)n;
    const double tau = 2.0 * sqrt((s_sum - 2.0 * mean_angle * sum) / (double)n + mean_angle * mean_angle);
    lsd_region_grow(F, a0, reg_size, reg_angle, tau);
    if (reg_size < 2) return false;
    lsd_region2rect(F, reg_size, reg_angle, prec, p, rec);
    density = (double)reg_size / (lsd_dist(rec.x1, rec.y1, rec.x2, rec.y2) * rec.width);
    if (density < density_th) return lsd_reduce_region_radius(F, reg_size, reg_angle, prec, p, rec, density, density_th);
    return true;
}

/* lsd.cpp:977-1098 with the vendored slips kept; the pixel count of a scan row is split over the lanes */
__device__ double lsd_rect_nfa(const LsdFrame &F, const LsdRect &rec)
{
    const int lane = threadIdx.x & 31;
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
    const double flstep = (my != ly) ? (double)((mx - lx) / (my - ly)) : 0;
    const double slstep = (ly != tx) ? (double)((lx - tx) / (ly - tx)) : 0;
    const double frstep = (my != ry) ? (double)((mx - rx) / (my - ry)) : 0;
    const double srstep = (ry != tx) ? (double)((rx - tx) / (ry - tx)) : 0;
    double lstep = flstep, rstep = frstep;
    double left_x = mx, right_x = mx;
    int total_pts = 0, alg_pts = 0;
    const int max_iter = oy[imax];
    for (int y = my; y <= max_iter; ++y) {
        if (y < 0 || y >= F.H) continue; /* as the reference: rows outside the image also skip the edge stepping */
        const int xa = (int)left_x, xb = (int)right_x;
        const int lo = max(xa, 0), hi = min(xb, F.W - 1);
        for (int x0 = lo; x0 <= hi; x0 += 32) {
            const int x = x0 + lane;
            const bool in = x <= hi;
            bool al = false;
            if (in) al = lsd_aligned_val(F.angles[y * F.W + x], rec.theta, rec.prec);
            total_pts += __popc(__ballot_sync(0xffffffffu, in));
            alg_pts += __popc(__ballot_sync(0xffffffffu, al));
        }
        if (y >= ly) lstep = slstep;
        if (y >= ry) rstep = srstep;
        left_x += lstep;
        right_x += rstep;
    }
    return lsd_nfa(total_pts, alg_pts, rec.p, F.LOG_NT);
}

/* lsd.cpp:873-975 */
__device__ double lsd_rect_improve(const LsdFrame &F, LsdRect &rec)
{
    const double delta = 0.5, delta_2 = delta / 2.0, LOG_EPS = 0;
    double log_nfa = lsd_rect_nfa(F, rec);
    if (log_nfa > LOG_EPS) return log_nfa;
    LsdRect r = rec;
    for (int n = 0; n < 5; ++n) {
        r.p /= 2;
        r.prec = r.p * LSD_PI;
        const double v = lsd_rect_nfa(F, r);
        if (v > log_nfa) {
            log_nfa = v;
            rec = r;
        }
    }
    if (log_nfa > LOG_EPS) return log_nfa;
    r = rec;
    for (int n = 0; n < 5; ++n)
        if ((r.width - delta) >= 0.5) {
            r.width -= delta;
            const double v = lsd_rect_nfa(F, r);
            if (v > log_nfa) {
                rec = r;
                log_nfa = v;
            }
        }
    if (log_nfa > LOG_EPS) return log_nfa;
    r = rec;
    for (int n = 0; n < 5; ++n)
        if ((r.width - delta) >= 0.5) {
            r.x1 += -r.dy * delta_2;
            r.y1 += r.dx * delta_2;
            r.x2 += -r.dy * delta_2;
            r.y2 += r.dx * delta_2;
            r.width -= delta;
            const double v = lsd_rect_nfa(F, r);
            if (v > log_nfa) {
                rec = r;
                log_nfa = v;
            }
        }
    if (log_nfa > LOG_EPS) return log_nfa;
    r = rec;
    for (int n = 0; n < 5; ++n)
        if ((r.width - delta) >= 0.5) {
            r.x1 -= -r.dy * delta_2;
            r.y1 -= r.dx * delta_2;
            r.x2 -= -r.dy * delta_2;
            r.y2 -= r.dx * delta_2;
            r.width -= delta;
            const double v = lsd_rect_nfa(F, r);
            if (v > log_nfa) {
                rec = r;
                log_nfa = v;
            }
        }
    if (log_nfa > LOG_EPS) return log_nfa;
    r = rec;
    for (int n = 0; n < 5; ++n)
        if ((r.width - delta) >= 0.5) {
            r.p /= 2;
            r.prec = r.p * LSD_PI;
            const double v = lsd_rect_nfa(F, r);
            if (v > log_nfa) {
                rec = r;
                log_nfa = v;
            }
        }
    return log_nfa;
}

/* The seed loop of flsd (lsd.cpp:476-535) + the KeyLine filters of LSDDetector::detectImpl (:205-256) and filter_lines.
 * One warp per frame. */
__global__ void __launch_bounds__(32) k_lsd_grow(int W, int H, int img_w, int img_h, const double *__restrict__ angles_all,
                                                 const double *__restrict__ modgrad_all, const float2 *__restrict__ cs_all, uint8_t *__restrict__ used_all,
                                                 int32_t *__restrict__ reg_all, const int32_t *__restrict__ list_all,
                                                 const int32_t *__restrict__ list_len, double LOG_NT, int min_reg_size, double prec, double p,
                                                 double scale, float line_length_thres, float *__restrict__ raw_all, int32_t *__restrict__ n_raw_all,
                                                 float *__restrict__ out_all, int32_t *__restrict__ n_out_all, int cap)
{
    const int f = blockIdx.x, lane = threadIdx.x;
    const size_t npx = (size_t)W * H;
    LsdFrame F;
    F.W = W;
    F.H = H;
    F.angles = angles_all + f * npx;
    F.modgrad = modgrad_all + f * npx;
    F.cs_angle = cs_all + f * npx;
    F.used = used_all + f * npx;
    F.reg = reg_all + f * npx;
    F.LOG_NT = LOG_NT;
    const int32_t *list = list_all + f * npx;
    const int n_list = list_len[f];
    float *raw = raw_all + (size_t)f * cap * 4;
    float *out = out_all + (size_t)f * cap * 4;
    int n_raw = 0, n_out = 0;
    const double DENSITY_TH = 0.7, LOG_EPS = 0;
    const float pre_boundary_thre = 10;

    for (int i0 = 0; i0 < n_list; i0 += 32) {
        /* 32 list entries at a time: seeds whose pixel is already used or has no defined angle are skipped by ballot */
        const int i = i0 + lane;
        int adx = 0;
        bool seed = false;
        if (i < n_list) {
            adx = list[i];
            seed = (F.used[adx] == 0); /* the list holds pixels with a defined angle only */
        }
        unsigned todo = __ballot_sync(0xffffffffu, seed);
        while (todo) {
            const int sl = __ffs(todo) - 1;
            todo &= todo - 1;
            const int s_addr = __shfl_sync(0xffffffffu, adx, sl);
            if (F.used[s_addr] != 0) continue; /* claimed by a region grown since the ballot */
            int reg_size;
            double reg_angle;
            lsd_region_grow(F, s_addr, reg_size, reg_angle, prec);
            if (reg_size < min_reg_size) continue;
            LsdRect rec;
            lsd_region2rect(F, reg_size, reg_angle, prec, p, rec);
            if (!lsd_refine(F, reg_size, reg_angle, prec, p, rec, DENSITY_TH)) continue;
            const double log_nfa = lsd_rect_improve(F, rec);
            if (log_nfa <= LOG_EPS) continue;
            rec.x1 += 0.5;
            rec.y1 += 0.5;
            rec.x2 += 0.5;
            rec.y2 += 0.5;
            rec.x1 /= scale;
            rec.y1 /= scale;
            rec.x2 /= scale;
            rec.y2 /= scale;
            float e[4] = {(float)rec.x1, (float)rec.y1, (float)rec.x2, (float)rec.y2};
            if (lane == 0 && n_raw < cap) {
                raw[4 * n_raw + 0] = e[0];
                raw[4 * n_raw + 1] = e[1];
                raw[4 * n_raw + 2] = e[2];
                raw[4 * n_raw + 3] = e[3];
            }
            n_raw++;
            /* checkLineExtremes + 10-px border rejection + length filter (LSDDetector.cpp:75-101,226-238; filter_lines) */
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
                continue;
            const double ddx = (double)(e[0] - e[2]), ddy = (double)(e[1] - e[3]);
            const float line_length = (float)sqrt(ddx * ddx + ddy * ddy);
            if (!(line_length > line_length_thres)) continue;
            if (lane == 0 && n_out < cap) {
                out[4 * n_out + 0] = sx;
                out[4 * n_out + 1] = sy;
                out[4 * n_out + 2] = ex;
                out[4 * n_out + 3] = ey;
            }
            n_out++;
        }
    }
    if (lane == 0) {
        n_raw_all[f] = n_raw;
        n_out_all[f] = n_out;
    }
}

/* ---------------------------------------------------------------------------------------- host side */
struct Buf {
    void *p = nullptr;
    size_t cap = 0;
};

struct LsdState {
    Buf img, tmp, blur, scaled, modgrad, angles, csang, used, list, reg, maxg, cnt, llen, raw, nraw, out, nout;
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

int lsd_run(cs_ctx *c, const uint8_t *imgs, bool imgs_on_device, int n_frames, int w, int h, int stride, int channels, float line_length_thres,
            int cap, LsdState &S)
{
    cudaStream_t st = cs_ctx_stream(c);
    const double SCALE = 0.8;
    const int W = (int)std::lrint(w * SCALE), H = (int)std::lrint(h * SCALE);
    if (W < 2 || H < 2) return cs_ctx_fail(c, CS_ERR_INVALID_ARG, "image too small for LSD");
    const size_t px = (size_t)n_frames * w * h, spx = (size_t)n_frames * W * H;
    const int n_chunks = (H - 1 + LSD_CHUNK_ROWS - 1) / LSD_CHUNK_ROWS;
    int rc;
    const uint8_t *d_img = imgs;
    if (!imgs_on_device) {
        if ((rc = ensure(c, S.img, (size_t)n_frames * h * stride))) return rc;
        if (cudaMemcpyAsync(S.img.p, imgs, (size_t)n_frames * h * stride, cudaMemcpyHostToDevice, st) != cudaSuccess)
            return cs_ctx_fail(c, CS_ERR_CUDA, "H2D copy of frames failed");
        d_img = (const uint8_t *)S.img.p;
    }
    if ((rc = ensure(c, S.tmp, px * 8)) || (rc = ensure(c, S.blur, px * 8)) || (rc = ensure(c, S.scaled, spx * 8)) ||
        (rc = ensure(c, S.modgrad, spx * 8)) || (rc = ensure(c, S.angles, spx * 8)) || (rc = ensure(c, S.csang, spx * 8)) || (rc = ensure(c, S.used, spx)) ||
        (rc = ensure(c, S.list, spx * 4)) || (rc = ensure(c, S.reg, spx * 4)) || (rc = ensure(c, S.maxg, (size_t)n_frames * 8)) ||
        (rc = ensure(c, S.cnt, (size_t)n_frames * n_chunks * LSD_NBINS * 4)) || (rc = ensure(c, S.llen, (size_t)n_frames * 4)) ||
        (rc = ensure(c, S.raw, (size_t)n_frames * cap * 16)) || (rc = ensure(c, S.nraw, (size_t)n_frames * 4)) ||
        (rc = ensure(c, S.out, (size_t)n_frames * cap * 16)) || (rc = ensure(c, S.nout, (size_t)n_frames * 4)))
        return rc;
    /* host-side constants of flsd (lsd.cpp:445-447,468-469), evaluated with libm like the reference */
    const double ANG_TH = 22.5, QUANT = 2.0;
    const double prec = LSD_PI * ANG_TH / 180, p = ANG_TH / 180, rho = QUANT / std::sin(prec);
    const double LOG_NT = 5 * (std::log10((double)W) + std::log10((double)H)) / 2 + std::log10(11.0);
    const int min_reg_size = (int)(-LOG_NT / std::log10(p));

    cudaMemsetAsync(S.maxg.p, 0, (size_t)n_frames * 8, st);
    cudaMemsetAsync(S.used.p, 0, spx, st);
    k_lsd_hblur<<<grid_for((int64_t)px), 256, 0, st>>>(d_img, n_frames, w, h, stride, channels, (double *)S.tmp.p);
    k_lsd_vblur<<<grid_for((int64_t)px), 256, 0, st>>>((const double *)S.tmp.p, n_frames, w, h, (double *)S.blur.p);
    k_lsd_resize<<<grid_for((int64_t)spx), 256, 0, st>>>((const double *)S.blur.p, n_frames, w, h, W, H, 1. / SCALE, (double *)S.scaled.p);
    k_lsd_grad<<<grid_for((int64_t)spx), 256, 0, st>>>((const double *)S.scaled.p, n_frames, W, H, rho, (double *)S.modgrad.p, (double *)S.angles.p,
                                                       (float2 *)S.csang.p, (unsigned long long *)S.maxg.p);
    k_lsd_hist<<<n_frames * n_chunks, 256, 0, st>>>((const double *)S.modgrad.p, (const double *)S.angles.p, W, H, n_chunks, (const unsigned long long *)S.maxg.p, (int32_t *)S.cnt.p);
    k_lsd_scan<<<n_frames, LSD_NBINS, 0, st>>>(n_chunks, (int32_t *)S.cnt.p, (int32_t *)S.llen.p);
    k_lsd_scatter<<<n_frames * n_chunks, 32, 0, st>>>((const double *)S.modgrad.p, (const double *)S.angles.p, W, H, n_chunks, (const unsigned long long *)S.maxg.p,
                                                      (const int32_t *)S.cnt.p, (int32_t *)S.list.p);
    k_lsd_grow<<<n_frames, 32, 0, st>>>(W, H, w, h, (const double *)S.angles.p, (const double *)S.modgrad.p, (const float2 *)S.csang.p, (uint8_t *)S.used.p,
                                        (int32_t *)S.reg.p,
                                        (const int32_t *)S.list.p, (const int32_t *)S.llen.p, LOG_NT, min_reg_size, prec, p, SCALE, line_length_thres,
                                        (float *)S.raw.p, (int32_t *)S.nraw.p, (float *)S.out.p, (int32_t *)S.nout.p, cap);
    cs_ctx_count_launches(c, 8);
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

void cs_lsd_destroy(void *state)
{
    LsdState *S = (LsdState *)state;
    Buf *all[] = {&S->img, &S->tmp, &S->blur, &S->scaled, &S->modgrad, &S->angles, &S->csang, &S->used, &S->list, &S->reg, &S->maxg, &S->cnt, &S->llen, &S->raw, &S->nraw, &S->out, &S->nout};
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
    if (params->numoctaves != 1) return cs_ctx_fail(c, CS_ERR_UNSUPPORTED, "only one octave is supported (filter_lines keeps octave 0 only)");
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
    if (angles) cudaMemcpy(angles, (double *)S->angles.p + frame * npx, npx * 8, cudaMemcpyDeviceToHost);
    int32_t ll = 0;
    cudaMemcpy(&ll, (int32_t *)S->llen.p + frame, 4, cudaMemcpyDeviceToHost);
    if (list_len) *list_len = ll;
    if (list) cudaMemcpy(list, (int32_t *)S->list.p + frame * npx, (size_t)ll * 4, cudaMemcpyDeviceToHost);
    int32_t nr = 0;
    cudaMemcpy(&nr, (int32_t *)S->nraw.p + frame, 4, cudaMemcpyDeviceToHost);
    if (n_raw) *n_raw = nr;
    if (raw_lines) cudaMemcpy(raw_lines, (float *)S->raw.p + (size_t)frame * S->cap * 4, (size_t)std::min(nr, std::min(cap_raw, S->cap)) * 16, cudaMemcpyDeviceToHost);
    return cudaGetLastError() == cudaSuccess ? CS_OK : cs_ctx_fail(c, CS_ERR_CUDA, "debug copy failed");
}
}
