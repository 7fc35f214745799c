/*
 * cs_edlines.cu -- line_lbd_detect::detect_filter_lines, EDLines flavour (use_LSD = false, the class default), for sm_100a.
 *
 * Replaces, for one octave:
 *   BinaryDescriptor::detectImpl / OctaveKeyLines      line_lbd/libs/binary_descriptor.cpp:483-585,792-886,1060-1141
 *   EDLineDetector::EdgeDrawing                        line_lbd/libs/binary_descriptor.cpp:1579-2377
 *   EDLineDetector::EDline / LeastSquaresLineFit_ / LineValidation_      :2379-2870
 *   nfa / log_gamma                                    line_lbd/include/line_lbd/line_descriptor/descriptor.hpp:680-830
 *
 * Streaming stages (thread per pixel, integer-exact):
 *   k_ed_hblur / k_ed_vblur   cvtColor + cv::GaussianBlur(8U, 5x5, sigma 1) == fixed-point kernel (14,62,104,62,14)/256, one final rounding
 *   k_ed_maps                 Sobel 3x3 (REFLECT_101), |dx|+|dy| thresholded at 81 then /4 (round half to even), direction map
 *   k_ed_anchors              anchors in the reference's column-major scan order (x outer, y inner, stride 2) by an ordered block compaction
 * Sequential stage:
 *   k_ed_route_fit            smart routing of the anchors in order (each chain claims pixels first come, first served), chain assembly,
 *                             least-squares line fitting with extension, NFA validation: one warp per frame, lane 0 walks (the walk is a
 *                             pointer chase through the gradient map; order defines the result).  Frames run in parallel.
 */
#include <cuda_runtime.h>
#include <float.h>
#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <vector>

#include "cs_internal.h"

#define ED_PI 3.14159265358979323846
#define ED_LN10 2.30258509299404568402
#define ED_HORIZONTAL 255
#define ED_UP 1
#define ED_RIGHT 2
#define ED_DOWN 3
#define ED_LEFT 4
#define ED_TRYTIME 6
#define ED_SKIP 2
#define ED_MINLEN 15
#define ED_FITERR 1.6

namespace {

__device__ __forceinline__ int ed_reflect101(int p, int n)
{
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = (p < 0) ? -p : 2 * n - 2 - p;
    return p;
}

__global__ void __launch_bounds__(256) k_ed_hblur(const uint8_t *__restrict__ img, int n_frames, int w, int h, int stride, int channels,
                                                  uint16_t *__restrict__ tmp)
{
    const int64_t total = (int64_t)n_frames * w * h;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
        const int64_t f = p / ((int64_t)w * h);
        const int r = (int)(p - f * (int64_t)w * h);
        const int y = r / w, x = r - y * w;
        const uint8_t *row = img + ((size_t)f * h + y) * stride;
        const int k[5] = {14, 62, 104, 62, 14};
        uint32_t s = 0;
#pragma unroll
        for (int i = 0; i < 5; i++) {
            const int xx = ed_reflect101(x + i - 2, w);
            uint32_t g;
            if (channels == 3) {
                const uint8_t *q = row + 3 * xx;
                g = (q[0] * 3735u + q[1] * 19235u + q[2] * 9798u + (1u << 14)) >> 15;
            } else
                g = row[xx];
            s += k[i] * g;
        }
        tmp[p] = (uint16_t)s; /* <= 255 * 256 */
    }
}

__global__ void __launch_bounds__(256) k_ed_vblur(const uint16_t *__restrict__ tmp, int n_frames, int w, int h, uint8_t *__restrict__ blur)
{
    const int64_t total = (int64_t)n_frames * w * h;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
        const int64_t f = p / ((int64_t)w * h);
        const int r = (int)(p - f * (int64_t)w * h);
        const int y = r / w, x = r - y * w;
        const uint16_t *base = tmp + (size_t)f * w * h;
        const int k[5] = {14, 62, 104, 62, 14};
        uint32_t s = 0;
#pragma unroll
        for (int j = 0; j < 5; j++) s += k[j] * (uint32_t)base[(size_t)ed_reflect101(y + j - 2, h) * w + x];
        blur[p] = (uint8_t)((s + 32768u) >> 16);
    }
}

__global__ void __launch_bounds__(256) k_ed_maps(const uint8_t *__restrict__ blur, int n_frames, int w, int h, int16_t *__restrict__ dxo,
                                                 int16_t *__restrict__ dyo, int16_t *__restrict__ go, uint8_t *__restrict__ diro)
{
    const int64_t total = (int64_t)n_frames * w * h;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
        const int64_t f = p / ((int64_t)w * h);
        const int r = (int)(p - f * (int64_t)w * h);
        const int y = r / w, x = r - y * w;
        const uint8_t *b = blur + (size_t)f * w * h;
        const int ym = ed_reflect101(y - 1, h), yp = ed_reflect101(y + 1, h), xm = ed_reflect101(x - 1, w), xp = ed_reflect101(x + 1, w);
        const int a00 = b[(size_t)ym * w + xm], a01 = b[(size_t)ym * w + x], a02 = b[(size_t)ym * w + xp];
        const int a10 = b[(size_t)y * w + xm], a12 = b[(size_t)y * w + xp];
        const int a20 = b[(size_t)yp * w + xm], a21 = b[(size_t)yp * w + x], a22 = b[(size_t)yp * w + xp];
        const int gx = (a02 + 2 * a12 + a22) - (a00 + 2 * a10 + a20);
        const int gy = (a20 + 2 * a21 + a22) - (a00 + 2 * a01 + a02);
        const int ax = abs(gx), ay = abs(gy), sum = ax + ay;
        const int s = (sum > 81) ? sum : 0; /* threshold(TOZERO, gradienThreshold_ + 1) */
        const int q = s >> 2, rr = s & 3;
        dxo[p] = (int16_t)gx;
        dyo[p] = (int16_t)gy;
        go[p] = (int16_t)(rr < 2 ? q : (rr == 3 ? q + 1 : q + (q & 1))); /* `mat / 4`: round half to even */
        diro[p] = (ax < ay) ? ED_HORIZONTAL : 0;
    }
}

/* anchors, reference scan order: for w = 1, 3, ...: for h = 1, 3, ... (binary_descriptor.cpp:1640-1666) */
__global__ void __launch_bounds__(256) k_ed_anchors(const int16_t *__restrict__ g_all, const uint8_t *__restrict__ dir_all, int w, int h,
                                                    int32_t *__restrict__ anchors_all, int32_t *__restrict__ n_anchors, int cap)
{
    __shared__ int s_warp[8];
    __shared__ int s_total;
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int16_t *g = g_all + (size_t)f * w * h;
    const uint8_t *dir = dir_all + (size_t)f * w * h;
    int32_t *out = anchors_all + (size_t)f * cap;
    const int nw = (w - 2 + 1) / 2, nh = (h - 2 + 1) / 2; /* w = 1,3,..,<= w-2 ; h likewise */
    if (tid == 0) s_total = 0;
    __syncthreads();
    const int total = nw * nh;
    for (int base = 0; base < total; base += 256) {
        const int i = base + tid;
        bool a = false;
        int idx = 0;
        if (i < total) {
            const int wi = i / nh, hi = i - wi * nh;
            const int x = 1 + 2 * wi, y = 1 + 2 * hi;
            idx = y * w + x;
            const int gv = g[idx];
            if (dir[idx] == ED_HORIZONTAL)
                a = (gv >= g[idx - w] + 8) && (gv >= g[idx + w] + 8);
            else
                a = (gv >= g[idx - 1] + 8) && (gv >= g[idx + 1] + 8);
        }
        const unsigned m = __ballot_sync(0xffffffffu, a);
        if (lane == 0) s_warp[wid] = __popc(m);
        __syncthreads();
        int off = s_total;
        for (int k = 0; k < wid; k++) off += s_warp[k];
        if (a) {
            const int slot = off + __popc(m & ((1u << lane) - 1u));
            if (slot < cap) out[slot] = idx;
        }
        __syncthreads();
        if (tid == 0) {
            int t = s_total;
            for (int k = 0; k < 8; k++) t += s_warp[k];
            s_total = t;
        }
        __syncthreads();
    }
    if (tid == 0) n_anchors[f] = s_total;
}

/* ---------------------------------------------------------------------------------------- sequential stage */
struct EdFrame {
    int W, H;
    const int16_t *dx, *dy, *g;
    const uint8_t *dir;
    uint8_t *edge;
    float ATA[4], ATV[2];
    double logNT;
};

__device__ __forceinline__ uint32_t ed_pack(int x, int y) { return (uint32_t)x | ((uint32_t)y << 16); }
__device__ __forceinline__ int ed_x(uint32_t p) { return (int)(p & 0xffffu); }
__device__ __forceinline__ int ed_y(uint32_t p) { return (int)(p >> 16); }

/* one smart-routing walk; appends packed pixels to out[*n] */
__device__ void ed_walk(EdFrame &F, int x, int y, int lastDirection, uint32_t *out, unsigned &n, unsigned cap, int &lastX, int &lastY)
{
    const int W = F.W, H = F.H;
    int idx = y * W + x;
    while (F.g[idx] > 0 && !F.edge[idx]) {
        F.edge[idx] = 1;
        if (n < cap) out[n] = ed_pack(x, y);
        n++;
        int shouldGo = 0;
        if (F.dir[idx] == ED_HORIZONTAL) {
            if (lastDirection == ED_UP || lastDirection == ED_DOWN) shouldGo = (x > lastX) ? ED_RIGHT : ED_LEFT;
            lastX = x;
            lastY = y;
            if (lastDirection == ED_RIGHT || shouldGo == ED_RIGHT) {
                if (x == W - 1 || y == 0 || y == H - 1) break;
                const uint8_t g1 = (uint8_t)F.g[idx - W + 1], g2 = (uint8_t)F.g[idx + 1], g3 = (uint8_t)F.g[idx + W + 1];
                if (g1 >= g2 && g1 >= g3) {
                    x = x + 1;
                    y = y - 1;
                } else if (g3 >= g2 && g3 >= g1) {
                    x = x + 1;
                    y = y + 1;
                } else
                    x = x + 1;
                lastDirection = ED_RIGHT;
            } else if (lastDirection == ED_LEFT || shouldGo == ED_LEFT) {
                if (x == 0 || y == 0 || y == H - 1) break;
                const uint8_t g1 = (uint8_t)F.g[idx - W - 1], g2 = (uint8_t)F.g[idx - 1], g3 = (uint8_t)F.g[idx + W - 1];
                if (g1 >= g2 && g1 >= g3) {
                    x = x - 1;
                    y = y - 1;
                } else if (g3 >= g2 && g3 >= g1) {
                    x = x - 1;
                    y = y + 1;
                } else
                    x = x - 1;
                lastDirection = ED_LEFT;
            }
        } else {
            if (lastDirection == ED_RIGHT || lastDirection == ED_LEFT) shouldGo = (y > lastY) ? ED_DOWN : ED_UP;
            lastX = x;
            lastY = y;
            if (lastDirection == ED_DOWN || shouldGo == ED_DOWN) {
                if (x == 0 || x == W - 1 || y == H - 1) break;
                const uint8_t g1 = (uint8_t)F.g[idx + W + 1], g2 = (uint8_t)F.g[idx + W], g3 = (uint8_t)F.g[idx + W - 1];
                if (g1 >= g2 && g1 >= g3) {
                    x = x + 1;
                    y = y + 1;
                } else if (g3 >= g2 && g3 >= g1) {
                    x = x - 1;
                    y = y + 1;
                } else
                    y = y + 1;
                lastDirection = ED_DOWN;
            } else if (lastDirection == ED_UP || shouldGo == ED_UP) {
                if (x == 0 || x == W - 1 || y == 0) break;
                const uint8_t g1 = (uint8_t)F.g[idx - W + 1], g2 = (uint8_t)F.g[idx - W], g3 = (uint8_t)F.g[idx - W - 1];
                if (g1 >= g2 && g1 >= g3) {
                    x = x + 1;
                    y = y - 1;
                } else if (g3 >= g2 && g3 >= g1) {
                    x = x - 1;
                    y = y - 1;
                } else
                    y = y - 1;
                lastDirection = ED_UP;
            }
        }
        idx = y * W + x;
    }
}

__device__ __forceinline__ bool ed_double_equal(double a, double b)
{
    if (a == b) return true;
    const double abs_diff = fabs(a - b);
    const double aa = fabs(a), bb = fabs(b);
    double abs_max = (aa > bb) ? aa : bb;
    if (abs_max < DBL_MIN) abs_max = DBL_MIN;
    return (abs_diff / abs_max) <= (100.0 * DBL_EPSILON);
}
__device__ double ed_log_gamma(double x)
{
    if (x > 15.0) return 0.918938533204673 + (x - 0.5) * log(x) - x + 0.5 * x * log(x * sinh(1 / x) + 1 / (810.0 * pow(x, 6.0)));
    const double q[7] = {75122.6331530, 80916.6278952, 36308.2951477, 8687.24529705, 1168.92649479, 83.8676043424, 2.50662827511};
    double a = (x + 0.5) * log(x + 5.5) - (x + 5.5);
    double b = 0.0;
    for (int n = 0; n < 7; n++) {
        a -= log(x + (double)n);
        b += q[n] * pow(x, (double)n);
    }
    return a + log(b);
}
__device__ double ed_nfa(int n, int k, double p, double logNT)
{
    const double tolerance = 0.1;
    if (n == 0 || k == 0) return -logNT;
    if (n == k) return -logNT - (double)n * log10(p);
    const double p_term = p / (1.0 - p);
    const double log1term = ed_log_gamma((double)n + 1.0) - ed_log_gamma((double)k + 1.0) - ed_log_gamma((double)(n - k) + 1.0) + (double)k * log(p) +
                            (double)(n - k) * log(1.0 - p);
    double term = exp(log1term);
    if (ed_double_equal(term, 0.0)) {
        if ((double)k > (double)n * p) return -log1term / ED_LN10 - logNT;
        return -logNT;
    }
    double bin_tail = term;
    for (int i = k + 1; i <= n; i++) {
        const double bin_term = (double)(n - i + 1) / (double)i;
        const double mult_term = bin_term * p_term;
        term *= mult_term;
        bin_tail += term;
        if (bin_term < 1.0) {
            const double err = term * ((1.0 - pow(mult_term, (double)(n - i + 1))) / (1.0 - mult_term) - 1.0);
            if (err < tolerance * fabs(-log10(bin_tail) - logNT) * bin_tail) break;
        }
    }
    return -log10(bin_tail) - logNT;
}

/* LeastSquaresLineFit_, first overload (:2628-2714).  Sums of products of small integers are exact in double; OpenCV's float gemm
 * accumulates in double and rounds once to float. */
__device__ double ed_fit_initial(EdFrame &F, const uint32_t *pts, unsigned offsetS, double *eq)
{
    const bool horiz = F.dir[ed_y(pts[offsetS]) * F.W + ed_x(pts[offsetS])] == ED_HORIZONTAL;
    double suu = 0, su = 0, suv = 0, sv = 0;
    for (int i = 0; i < ED_MINLEN; i++) {
        const uint32_t p = pts[offsetS + i];
        const double u = horiz ? ed_x(p) : ed_y(p), v = horiz ? ed_y(p) : ed_x(p);
        suu += u * u;
        su += u;
        suv += u * v;
        sv += v;
    }
    F.ATA[0] = (float)suu;
    F.ATA[1] = (float)su;
    F.ATA[2] = (float)su;
    F.ATA[3] = (float)(double)ED_MINLEN;
    F.ATV[0] = (float)suv;
    F.ATV[1] = (float)sv;
    const double coef = 1.0 / ((double)F.ATA[0] * (double)F.ATA[3] - (double)F.ATA[1] * (double)F.ATA[2]);
    eq[0] = coef * ((double)F.ATA[3] * (double)F.ATV[0] - (double)F.ATA[1] * (double)F.ATV[1]);
    eq[1] = coef * ((double)F.ATA[0] * (double)F.ATV[1] - (double)F.ATA[2] * (double)F.ATV[0]);
    double fitError = 0;
    for (int i = 0; i < ED_MINLEN; i++) {
        const uint32_t p = pts[offsetS + i];
        const double u = horiz ? ed_x(p) : ed_y(p), v = horiz ? ed_y(p) : ed_x(p);
        const double c = v - u * eq[0] - eq[1];
        fitError += c * c;
    }
    return sqrt(fitError);
}

/* second overload (:2716-2787) */
__device__ void ed_fit_update(EdFrame &F, const uint32_t *pts, unsigned offsetS, unsigned newOffsetS, unsigned offsetE, double *eq)
{
    const int length = (int)offsetE - (int)offsetS, newLength = (int)offsetE - (int)newOffsetS;
    if (length <= 0 || newLength <= 0) return;
    const bool horiz = F.dir[ed_y(pts[offsetS]) * F.W + ed_x(pts[offsetS])] == ED_HORIZONTAL;
    double suu = 0, su = 0, suv = 0, sv = 0;
    for (unsigned i = newOffsetS; i < offsetE; i++) {
        const uint32_t p = pts[i];
        const double u = horiz ? ed_x(p) : ed_y(p), v = horiz ? ed_y(p) : ed_x(p);
        suu += u * u;
        su += u;
        suv += u * v;
        sv += v;
    }
    F.ATA[0] = F.ATA[0] + (float)suu;
    F.ATA[1] = F.ATA[1] + (float)su;
    F.ATA[2] = F.ATA[2] + (float)su;
    F.ATA[3] = F.ATA[3] + (float)(double)newLength;
    F.ATV[0] = F.ATV[0] + (float)suv;
    F.ATV[1] = F.ATV[1] + (float)sv;
    const double coef = 1.0 / ((double)F.ATA[0] * (double)F.ATA[3] - (double)F.ATA[1] * (double)F.ATA[2]);
    eq[0] = coef * ((double)F.ATA[3] * (double)F.ATV[0] - (double)F.ATA[1] * (double)F.ATV[1]);
    eq[1] = coef * ((double)F.ATA[0] * (double)F.ATV[1] - (double)F.ATA[2] * (double)F.ATV[0]);
}

/* LineValidation_ (:2789-2870) */
__device__ bool ed_validate(EdFrame &F, const uint32_t *pts, unsigned offsetS, unsigned offsetE, const double *lineEqu, float &direction)
{
    const int n = (int)offsetE - (int)offsetS;
    int meanGradientX = 0, meanGradientY = 0;
    for (int i = 0; i < n; i++) {
        const uint32_t p = pts[offsetS + i];
        const int index = ed_y(p) * F.W + ed_x(p);
        meanGradientX += F.dx[index];
        meanGradientY += F.dy[index];
    }
    const double ddx = fabs(lineEqu[1]), ddy = fabs(lineEqu[0]);
    if (meanGradientX == 0 && meanGradientY == 0) return false;
    if (meanGradientX > 0 && meanGradientY >= 0) direction = (float)atan2(-ddy, ddx);
    if (meanGradientX <= 0 && meanGradientY > 0) direction = (float)atan2(ddy, ddx);
    if (meanGradientX < 0 && meanGradientY <= 0) direction = (float)atan2(ddy, -ddx);
    if (meanGradientX >= 0 && meanGradientY < 0) direction = (float)atan2(-ddy, -ddx);
    if (fabs((double)direction) < 0.15 || ED_PI - fabs((double)direction) < 0.15) {
        if (fabs(lineEqu[2]) < 10 || fabs((double)F.H - fabs(lineEqu[2])) < 10) return false;
    }
    if (fabs(fabs((double)direction) - ED_PI * 0.5) < 0.15) {
        if (fabs(lineEqu[2]) < 10 || fabs((double)F.W - fabs(lineEqu[2])) < 10) return false;
    }
    int k = 0;
    for (int i = 0; i < n; i++) {
        const uint32_t p = pts[offsetS + i];
        const int index = ed_y(p) * F.W + ed_x(p);
        const double pd = atan2(-(double)F.dx[index], (double)F.dy[index]);
        const double dis = fabs((double)direction - pd);
        if (fabs(2 * ED_PI - dis) < 0.392699 || dis < 0.392699) k++;
    }
    return ed_nfa(n, k, 0.125, F.logNT) > 0;
}

/* routing + fitting for one frame by lane 0 of a warp */
__global__ void __launch_bounds__(32) k_ed_route_fit(int W, int H, const int16_t *__restrict__ dx_all, const int16_t *__restrict__ dy_all,
                                                     const int16_t *__restrict__ g_all, const uint8_t *__restrict__ dir_all,
                                                     uint8_t *__restrict__ edge_all, const int32_t *__restrict__ anchors_all,
                                                     const int32_t *__restrict__ n_anchors, int anchor_cap, uint32_t *__restrict__ scratch_all,
                                                     size_t scratch_per_frame, double logNT, float line_length_thres, float *__restrict__ raw_all,
                                                     int32_t *__restrict__ n_raw_all, float *__restrict__ out_all, int32_t *__restrict__ n_out_all,
                                                     int cap, int32_t *__restrict__ err_flag)
{
    if (threadIdx.x != 0) return;
    const int f = blockIdx.x;
    const size_t npx = (size_t)W * H;
    EdFrame F;
    F.W = W;
    F.H = H;
    F.dx = dx_all + f * npx;
    F.dy = dy_all + f * npx;
    F.g = g_all + f * npx;
    F.dir = dir_all + f * npx;
    F.edge = edge_all + f * npx;
    F.logNT = logNT;
    for (int i = 0; i < 4; i++) F.ATA[i] = 0;
    F.ATV[0] = F.ATV[1] = 0;
    const unsigned P = (unsigned)(npx / 5), maxEdges = P / 20;
    uint32_t *scratch = scratch_all + (size_t)f * scratch_per_frame;
    uint32_t *fP = scratch;                 /* first parts   [P]        */
    uint32_t *sP = fP + P;                  /* second parts  [P]        */
    uint32_t *fS = sP + P;                  /* [maxEdges + 2]           */
    uint32_t *sS = fS + maxEdges + 2;       /* [maxEdges + 2]           */
    uint32_t *chain = sS + maxEdges + 2;    /* assembled chains [2P]    */
    uint32_t *sId = chain + 2 * P;          /* [maxEdges + 2]           */
    uint32_t *lpt = sId + maxEdges + 2;     /* line pixels [2P]         */
    float *raw = raw_all + (size_t)f * cap * 4;
    float *out = out_all + (size_t)f * cap * 4;
    n_raw_all[f] = 0;
    n_out_all[f] = 0;

    const int32_t *anchors = anchors_all + (size_t)f * anchor_cap;
    const int na = n_anchors[f];
    if (na > anchor_cap || (unsigned)na > P) { /* reference: "anchor size is larger than its maximal size" -> no lines */
        if (na > anchor_cap) atomicOr(err_flag, 1);
        return;
    }
    unsigned nF = 0, nS = 0, nE = 0;
    int lastX = 0, lastY = 0;
    for (int i = 0; i < na; i++) {
        const int idx = anchors[i];
        if (F.edge[idx]) continue;
        const int y = idx / W, x = idx - y * W;
        const unsigned f0 = nF, s0 = nS;
        if (F.dir[idx] == ED_HORIZONTAL) {
            ed_walk(F, x, y, ED_RIGHT, fP, nF, P, lastX, lastY);
            F.edge[idx] = 0;
            ed_walk(F, x, y, ED_LEFT, sP, nS, P, lastX, lastY);
        } else {
            ed_walk(F, x, y, ED_DOWN, fP, nF, P, lastX, lastY);
            F.edge[idx] = 0;
            ed_walk(F, x, y, ED_UP, sP, nS, P, lastX, lastY);
        }
        if ((int)(nF - f0) + (int)(nS - s0) < ED_MINLEN + 1) {
            nF = f0;
            nS = s0;
        } else {
            if (nE <= maxEdges) {
                fS[nE] = f0;
                sS[nE] = s0;
            }
            nE++;
        }
    }
    if (nE > maxEdges || nF > P || nS > P) return; /* reference prints "Edge drawing Error" and detects nothing */
    fS[nE] = nF;
    sS[nE] = nS;
    /* chain assembly: first part reversed, then the second part without the anchor */
    unsigned nc = 0;
    for (unsigned e = 0; e < nE; e++) {
        sId[e] = nc;
        for (int t = (int)fS[e + 1] - 1; t >= (int)fS[e]; t--) chain[nc++] = fP[t];
        for (int t = (int)sS[e] + 1; t < (int)sS[e + 1]; t++) chain[nc++] = sP[t];
    }
    sId[nE] = nc;
    if (nE == 0) return;

    int n_raw = 0, n_out = 0;
    double lineFitErr = 0, eq[2] = {0, 0};
    unsigned offL = 0, newOffsetS = 0;
    float direction = 0;
    for (unsigned edgeID = 0; edgeID < nE; edgeID++) {
        unsigned S = sId[edgeID];
        const unsigned Eend = sId[edgeID + 1];
        while (Eend > S + ED_MINLEN) {
            while (Eend > S + ED_MINLEN) {
                lineFitErr = ed_fit_initial(F, chain, S, eq);
                if (lineFitErr <= ED_FITERR) break;
                S += ED_SKIP;
            }
            if (lineFitErr > ED_FITERR) break;
            const unsigned lineStart = offL;
            double coef1 = 0;
            bool bExtended = true, bFirstTry = true;
            int numOfOutlier, tryTimes = 0;
            const bool horiz = F.dir[ed_y(chain[S]) * W + ed_x(chain[S])] == ED_HORIZONTAL;
            while (bExtended) {
                tryTimes++;
                if (bFirstTry) {
                    bFirstTry = false;
                    for (int i = 0; i < ED_MINLEN; i++) lpt[offL++] = chain[S++];
                } else
                    ed_fit_update(F, lpt, lineStart, newOffsetS, offL, eq);
                coef1 = horiz ? 1 / sqrt(eq[0] * eq[0] + 1) : 1 / sqrt(1 + eq[0] * eq[0]);
                numOfOutlier = 0;
                newOffsetS = offL;
                while (Eend > S) {
                    const uint32_t p = chain[S];
                    const double d = horiz ? fabs(eq[0] * (double)ed_x(p) - (double)ed_y(p) + eq[1]) * coef1
                                           : fabs((double)ed_x(p) - eq[0] * (double)ed_y(p) - eq[1]) * coef1;
                    lpt[offL++] = p;
                    S++;
                    if (d > ED_FITERR) {
                        numOfOutlier++;
                        if (numOfOutlier > 3) break;
                    } else
                        numOfOutlier = 0;
                }
                offL -= numOfOutlier;
                S -= numOfOutlier;
                if (!(offL - newOffsetS > 0 && tryTimes < ED_TRYTIME)) bExtended = false;
            }
            double lineEqu[3];
            if (horiz) {
                lineEqu[0] = eq[0] * coef1;
                lineEqu[1] = -1 * coef1;
                lineEqu[2] = eq[1] * coef1;
            } else {
                lineEqu[0] = 1 * coef1;
                lineEqu[1] = -eq[0] * coef1;
                lineEqu[2] = -eq[1] * coef1;
            }
            if (ed_validate(F, lpt, lineStart, offL, lineEqu, direction)) {
                const double a1 = lineEqu[1] * lineEqu[1], a2 = lineEqu[0] * lineEqu[0], a3 = lineEqu[0] * lineEqu[1];
                const double a4 = lineEqu[2] * lineEqu[0], a5 = lineEqu[2] * lineEqu[1];
                unsigned Px = ed_x(lpt[lineStart]), Py = ed_y(lpt[lineStart]);
                const float s1 = (float)(a1 * Px - a3 * Py - a4), s2 = (float)(a2 * Py - a3 * Px - a5);
                Px = ed_x(lpt[offL - 1]);
                Py = ed_y(lpt[offL - 1]);
                const float e1 = (float)(a1 * Px - a3 * Py - a4), e2 = (float)(a2 * Py - a3 * Px - a5);
                /* OctaveKeyLines: length and start / end ordering (:862-886,1069-1139) */
                float fdx = fabsf(s1 - e1), fdy = fabsf(s2 - e2);
                const float lineLength = sqrtf(fdx * fdx + fdy * fdy);
                fdx = e1 - s1;
                fdy = e2 - s2;
                bool sw = false;
                const double dr = (double)direction;
                if (dr >= -0.75 * ED_PI && dr < -0.25 * ED_PI && fdy > 0) sw = true;
                if (dr >= -0.25 * ED_PI && dr < 0.25 * ED_PI && fdx < 0) sw = true;
                if (dr >= 0.25 * ED_PI && dr < 0.75 * ED_PI && fdy < 0) sw = true;
                if (((dr >= 0.75 * ED_PI && dr < ED_PI) || (dr >= -ED_PI && dr < -0.75 * ED_PI)) && fdx > 0) sw = true;
                const float sx = sw ? e1 : s1, sy = sw ? e2 : s2, ex = sw ? s1 : e1, ey = sw ? s2 : e2;
                if (n_raw < cap) {
                    raw[4 * n_raw + 0] = sx;
                    raw[4 * n_raw + 1] = sy;
                    raw[4 * n_raw + 2] = ex;
                    raw[4 * n_raw + 3] = ey;
                }
                n_raw++;
                if (lineLength > line_length_thres) {
                    if (n_out < cap) {
                        out[4 * n_out + 0] = sx;
                        out[4 * n_out + 1] = sy;
                        out[4 * n_out + 2] = ex;
                        out[4 * n_out + 3] = ey;
                    }
                    n_out++;
                }
            } else
                offL = lineStart;
        }
    }
    n_raw_all[f] = n_raw;
    n_out_all[f] = n_out;
}

/* ---------------------------------------------------------------------------------------- host side */
struct Buf {
    void *p = nullptr;
    size_t cap = 0;
};
struct EdState {
    Buf img, tmp, blur, dx, dy, g, dir, edge, anchors, nanch, scratch, raw, nraw, out, nout, err;
    int last_frames = 0, last_w = 0, last_h = 0, cap = 0, anchor_cap = 0;
};

int ed_ensure(cs_ctx *c, Buf &b, size_t bytes)
{
    if (bytes <= b.cap) return CS_OK;
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
    const size_t want = bytes + bytes / 16 + 256;
    if (cudaMalloc(&b.p, want) != cudaSuccess) return cs_ctx_fail(c, CS_ERR_CUDA, "cudaMalloc(%zu) failed in the EDLines detector", want);
    b.cap = want;
    return CS_OK;
}
inline int ed_grid(int64_t n) { return (int)std::min<int64_t>((n + 255) / 256, 148 * 32); }

EdState *ed_state_of(cs_ctx *c)
{
    void **slot = cs_ctx_edl_slot(c);
    if (!*slot) *slot = new EdState();
    return (EdState *)*slot;
}

}  // namespace

void cs_edl_destroy(void *state)
{
    EdState *S = (EdState *)state;
    Buf *all[] = {&S->img, &S->tmp, &S->blur, &S->dx, &S->dy, &S->g, &S->dir, &S->edge, &S->anchors, &S->nanch, &S->scratch, &S->raw, &S->nraw, &S->out, &S->nout, &S->err};
    for (Buf *b : all)
        if (b->p) cudaFree(b->p);
    delete S;
}

/* frames in HBM (or host) -> filtered segments in HBM */
int cs_edl_run(cs_ctx *c, const uint8_t *imgs, bool imgs_on_device, int n_frames, int w, int h, int stride, int channels, float line_length_thres,
               int cap, const float **d_lines, const int32_t **d_counts)
{
    EdState &S = *ed_state_of(c);
    cudaStream_t st = cs_ctx_stream(c);
    if (w < 8 || h < 8 || w > 65535 || h > 65535) return cs_ctx_fail(c, CS_ERR_INVALID_ARG, "image size unsupported by EDLines");
    const size_t px = (size_t)n_frames * w * h, npx = (size_t)w * h;
    const unsigned P = (unsigned)(npx / 5), maxEdges = P / 20;
    const size_t scratch_per_frame = (size_t)P * 2 + (size_t)(maxEdges + 2) * 3 + (size_t)P * 4 + 64;
    const int anchor_cap = (int)P + 1;
    int rc;
    const uint8_t *d_img = imgs;
    if (!imgs_on_device) {
        if ((rc = ed_ensure(c, S.img, (size_t)n_frames * h * stride))) return rc;
        if (cudaMemcpyAsync(S.img.p, imgs, (size_t)n_frames * h * stride, cudaMemcpyHostToDevice, st) != cudaSuccess)
            return cs_ctx_fail(c, CS_ERR_CUDA, "H2D copy of frames failed");
        d_img = (const uint8_t *)S.img.p;
    }
    if ((rc = ed_ensure(c, S.tmp, px * 2)) || (rc = ed_ensure(c, S.blur, px)) || (rc = ed_ensure(c, S.dx, px * 2)) || (rc = ed_ensure(c, S.dy, px * 2)) ||
        (rc = ed_ensure(c, S.g, px * 2)) || (rc = ed_ensure(c, S.dir, px)) || (rc = ed_ensure(c, S.edge, px)) ||
        (rc = ed_ensure(c, S.anchors, (size_t)n_frames * anchor_cap * 4)) || (rc = ed_ensure(c, S.nanch, (size_t)n_frames * 4)) ||
        (rc = ed_ensure(c, S.scratch, (size_t)n_frames * scratch_per_frame * 4)) || (rc = ed_ensure(c, S.raw, (size_t)n_frames * cap * 16)) ||
        (rc = ed_ensure(c, S.nraw, (size_t)n_frames * 4)) || (rc = ed_ensure(c, S.out, (size_t)n_frames * cap * 16)) ||
        (rc = ed_ensure(c, S.nout, (size_t)n_frames * 4)) || (rc = ed_ensure(c, S.err, 16)))
        return rc;
    const double logNT = 2.0 * (std::log10((double)w) + std::log10((double)h)); /* :2399, host libm like the reference */
    cudaMemsetAsync(S.edge.p, 0, px, st);
    cudaMemsetAsync(S.err.p, 0, 16, st);
    k_ed_hblur<<<ed_grid((int64_t)px), 256, 0, st>>>(d_img, n_frames, w, h, stride, channels, (uint16_t *)S.tmp.p);
    k_ed_vblur<<<ed_grid((int64_t)px), 256, 0, st>>>((const uint16_t *)S.tmp.p, n_frames, w, h, (uint8_t *)S.blur.p);
    k_ed_maps<<<ed_grid((int64_t)px), 256, 0, st>>>((const uint8_t *)S.blur.p, n_frames, w, h, (int16_t *)S.dx.p, (int16_t *)S.dy.p, (int16_t *)S.g.p,
                                                    (uint8_t *)S.dir.p);
    k_ed_anchors<<<n_frames, 256, 0, st>>>((const int16_t *)S.g.p, (const uint8_t *)S.dir.p, w, h, (int32_t *)S.anchors.p, (int32_t *)S.nanch.p, anchor_cap);
    k_ed_route_fit<<<n_frames, 32, 0, st>>>(w, h, (const int16_t *)S.dx.p, (const int16_t *)S.dy.p, (const int16_t *)S.g.p, (const uint8_t *)S.dir.p,
                                            (uint8_t *)S.edge.p, (const int32_t *)S.anchors.p, (const int32_t *)S.nanch.p, anchor_cap,
                                            (uint32_t *)S.scratch.p, scratch_per_frame, logNT, line_length_thres, (float *)S.raw.p, (int32_t *)S.nraw.p,
                                            (float *)S.out.p, (int32_t *)S.nout.p, cap, (int32_t *)S.err.p);
    cs_ctx_count_launches(c, 5);
    if (cudaGetLastError() != cudaSuccess) return cs_ctx_fail(c, CS_ERR_CUDA, "EDLines kernel launch failed");
    S.last_frames = n_frames;
    S.last_w = w;
    S.last_h = h;
    S.cap = cap;
    S.anchor_cap = anchor_cap;
    if (d_lines) *d_lines = (const float *)S.out.p;
    if (d_counts) *d_counts = (const int32_t *)S.nout.p;
    return CS_OK;
}

extern "C" int cs_debug_edlines(cs_ctx *c, int frame, uint8_t *blur, int16_t *dx, int16_t *dy, int16_t *g, uint8_t *dir, int32_t *anchors,
                                int32_t *n_anchors, uint8_t *edge, float *raw_lines, int32_t *n_raw, int cap_raw)
{
    if (!c) return CS_ERR_INVALID_ARG;
    EdState *S = ed_state_of(c);
    if (frame < 0 || frame >= S->last_frames) return cs_ctx_fail(c, CS_ERR_INVALID_ARG, "bad frame index");
    cudaSetDevice(cs_ctx_device(c));
    cudaStreamSynchronize(cs_ctx_stream(c));
    const size_t npx = (size_t)S->last_w * S->last_h;
    if (blur) cudaMemcpy(blur, (uint8_t *)S->blur.p + frame * npx, npx, cudaMemcpyDeviceToHost);
    if (dx) cudaMemcpy(dx, (int16_t *)S->dx.p + frame * npx, npx * 2, cudaMemcpyDeviceToHost);
    if (dy) cudaMemcpy(dy, (int16_t *)S->dy.p + frame * npx, npx * 2, cudaMemcpyDeviceToHost);
    if (g) cudaMemcpy(g, (int16_t *)S->g.p + frame * npx, npx * 2, cudaMemcpyDeviceToHost);
    if (dir) cudaMemcpy(dir, (uint8_t *)S->dir.p + frame * npx, npx, cudaMemcpyDeviceToHost);
    if (edge) cudaMemcpy(edge, (uint8_t *)S->edge.p + frame * npx, npx, cudaMemcpyDeviceToHost);
    int32_t na = 0, nr = 0;
    cudaMemcpy(&na, (int32_t *)S->nanch.p + frame, 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(&nr, (int32_t *)S->nraw.p + frame, 4, cudaMemcpyDeviceToHost);
    if (n_anchors) *n_anchors = na;
    if (n_raw) *n_raw = nr;
    if (anchors) cudaMemcpy(anchors, (int32_t *)S->anchors.p + (size_t)frame * S->anchor_cap, (size_t)std::min(na, S->anchor_cap) * 4, cudaMemcpyDeviceToHost);
    if (raw_lines) cudaMemcpy(raw_lines, (float *)S->raw.p + (size_t)frame * S->cap * 4, (size_t)std::min(nr, std::min(cap_raw, S->cap)) * 16, cudaMemcpyDeviceToHost);
    return cudaGetLastError() == cudaSuccess ? CS_OK : cs_ctx_fail(c, CS_ERR_CUDA, "debug copy failed");
}
