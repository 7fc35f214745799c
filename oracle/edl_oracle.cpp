/*
 * oracle/edl_oracle.cpp -- CPU ORACLE for the EDLines flavour of line_lbd_detect (use_LSD = false, the class default).
 * TEST INFRASTRUCTURE ONLY.
 *
 * Restates, without OpenCV, for numOfOctave_ = 1 (line_lbd_allclass.cpp:110-123 builds the descriptor with the ctor's octave count;
 * filter_lines keeps octave 0 only):
 *   line_lbd/class/line_lbd_allclass.cpp:143-147,200-221          detect_raw_lines (mask of ones), filter_lines, keylines_to_mat
 *   line_lbd/libs/binary_descriptor.cpp:483-585                   BinaryDescriptor::detectImpl (cvtColor, KeyLine fill)
 *   line_lbd/libs/binary_descriptor.cpp:792-838,862-886,1060-1141 OctaveKeyLines for one octave (5x5 Gaussian, sigma 1; start/end ordering)
 *   line_lbd/libs/binary_descriptor.cpp:1511-1522                 EDLineDetector defaults: gradient threshold 80, anchor threshold 8,
 *                                                                 scan interval 2, min line length 15, fit error 1.6
 *   line_lbd/libs/binary_descriptor.cpp:1579-2377                 EdgeDrawing (Sobel maps, anchors scanned column-major, smart routing with the
 *                                                                 short -> unsigned char neighbour casts, chain assembly)
 *   line_lbd/libs/binary_descriptor.cpp:2379-2626                 EDline (initial fit, extension with <= 3 consecutive outliers, 6 tries)
 *   line_lbd/libs/binary_descriptor.cpp:2628-2787                 LeastSquaresLineFit_ (float normal equations, double solve)
 *   line_lbd/libs/binary_descriptor.cpp:2789-2870                 LineValidation_ (direction, border rejection, NFA with p = 1/8)
 *   line_lbd/include/line_lbd/line_descriptor/descriptor.hpp:680-830   log_gamma, nfa
 * OpenCV calls inside are restated and pinned against cv2 4.13: GaussianBlur(CV_8U, 5x5, sigma 1) == fixed-point kernel
 * (14, 62, 104, 62, 14) / 256 in both directions with one final (+32768) >> 16; Sobel 3x3 with BORDER_REFLECT_101;
 * threshold(TOZERO, 81) on |dx|+|dy|; `mat / 4` == round-half-to-even; compare(CMP_LT).
 *
 * PARITY: PINNED to the reference.  The reference's own binary_descriptor.cpp and headers, compiled from /root/reference into
 * oracle/_ref/libedl_ref.so (oracle/ref/edl_ref.cpp, oracle/Makefile target `ref`) and driven as detect_raw_lines drives them, return
 * byte-identical key lines on every fixture, synthetic and degenerate image tried (tests/test_oracle_ref_edlines.py); counts and
 * checksums are recorded in the committed goldens (tests/test_goldens_sequence.py).  One input class has no reference output to compare
 * with: EDLineDetector::EDline sizes lines.sId as 5 x (number of chains) (binary_descriptor.cpp:2394) and overruns the heap when chains
 * split into more segments than that (AddressSanitizer: heap-buffer-overflow at :2438, e.g. one long curved chain); this restatement
 * keeps every segment.
 */
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "orc_api.h"

extern "C" void orc_bgr2gray(const uint8_t *bgr, int w, int h, int stride, uint8_t *gray, int gstride, int fixed15);

namespace {

const uint8_t kHorizontal = 255; /* |dx| < |dy| */
const int kUp = 1, kRight = 2, kDown = 3, kLeft = 4;
const int kTryTime = 6, kSkipEdgePoint = 2;
const double kLN10 = 2.30258509299404568402;

inline int reflect101(int p, int n)
{
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = (p < 0) ? -p : 2 * n - 2 - p;
    return p;
}

/* cv::GaussianBlur(CV_8U, Size(5,5), 1.0): OpenCV's bit-exact fixed-point smoothing */
void gaussian5_u8(const std::vector<uint8_t> &src, int w, int h, std::vector<uint8_t> &dst)
{
    static const int k[5] = {14, 62, 104, 62, 14};
    std::vector<uint32_t> t((size_t)w * h);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            uint32_t s = 0;
            for (int i = 0; i < 5; i++) s += k[i] * src[(size_t)y * w + reflect101(x + i - 2, w)];
            t[(size_t)y * w + x] = s;
        }
    dst.resize((size_t)w * h);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            uint32_t s = 0;
            for (int j = 0; j < 5; j++) s += k[j] * t[(size_t)reflect101(y + j - 2, h) * w + x];
            dst[(size_t)y * w + x] = (uint8_t)((s + 32768u) >> 16);
        }
}

} // namespace (reopened below)
/* for oracle/ref/minicv.hpp: the one restatement of OpenCV's 8-bit 5 x 5 Gaussian, pinned against cv2 (tests/test_oracle_cv_parity.py) */
extern "C" void edl_orc_gaussian5_u8(const uint8_t *src, int w, int h, uint8_t *dst)
{
    std::vector<uint8_t> in(src, src + (size_t)w * h), out;
    gaussian5_u8(in, w, h, out);
    std::memcpy(dst, out.data(), out.size());
}
namespace {

inline int div4_half_even(int s)
{
    const int q = s >> 2, r = s & 3;
    return r < 2 ? q : (r == 3 ? q + 1 : q + (q & 1));
}

inline bool double_equal(double a, double b)
{
    if (a == b) return true;
    const double abs_diff = std::fabs(a - b);
    const double aa = std::fabs(a), bb = std::fabs(b);
    double abs_max = (aa > bb) ? aa : bb;
    if (abs_max < DBL_MIN) abs_max = DBL_MIN;
    return (abs_diff / abs_max) <= (100.0 * DBL_EPSILON);
}
inline double log_gamma_windschitl(double x) { return 0.918938533204673 + (x - 0.5) * std::log(x) - x + 0.5 * x * std::log(x * std::sinh(1 / x) + 1 / (810.0 * std::pow(x, 6.0))); }
inline double log_gamma_lanczos(double x)
{
    static const double q[7] = {75122.6331530, 80916.6278952, 36308.2951477, 8687.24529705, 1168.92649479, 83.8676043424, 2.50662827511};
    double a = (x + 0.5) * std::log(x + 5.5) - (x + 5.5);
    double b = 0.0;
    for (int n = 0; n < 7; n++) {
        a -= std::log(x + (double)n);
        b += q[n] * std::pow(x, (double)n);
    }
    return a + std::log(b);
}
inline double log_gamma(double x) { return x > 15.0 ? log_gamma_windschitl(x) : log_gamma_lanczos(x); }

/* descriptor.hpp:763-830 */
double nfa(int n, int k, double p, double logNT)
{
    const double tolerance = 0.1;
    if (n == 0 || k == 0) return -logNT;
    if (n == k) return -logNT - (double)n * std::log10(p);
    const double p_term = p / (1.0 - p);
    const double log1term = log_gamma((double)n + 1.0) - log_gamma((double)k + 1.0) - log_gamma((double)(n - k) + 1.0) + (double)k * std::log(p) +
                            (double)(n - k) * std::log(1.0 - p);
    double term = std::exp(log1term);
    if (double_equal(term, 0.0)) {
        if ((double)k > (double)n * p) return -log1term / kLN10 - logNT;
        return -logNT;
    }
    double bin_tail = term;
    for (int i = k + 1; i <= n; i++) {
        const double bin_term = (double)(n - i + 1) / (double)i;
        const double mult_term = bin_term * p_term;
        term *= mult_term;
        bin_tail += term;
        if (bin_term < 1.0) {
            const double err = term * ((1.0 - std::pow(mult_term, (double)(n - i + 1))) / (1.0 - mult_term) - 1.0);
            if (err < tolerance * std::fabs(-std::log10(bin_tail) - logNT) * bin_tail) break;
        }
    }
    return -std::log10(bin_tail) - logNT;
}

struct Ed {
    int W = 0, H = 0;
    std::vector<int16_t> dx, dy, g;
    std::vector<uint8_t> dir, edge;
    /* fit state shared between the two LeastSquaresLineFit_ overloads (ATA, ATV as float matrices) */
    float ATA[4] = {0, 0, 0, 0}, ATV[2] = {0, 0};
    double logNT = 0;
    int minLineLen = 15;
    double fitErrThr = 1.6;

    /* one smart-routing walk (binary_descriptor.cpp:1713-1862 and its three copies) */
    void walk(int x, int y, int lastDirection, std::vector<uint32_t> &outX, std::vector<uint32_t> &outY, int &lastX, int &lastY)
    {
        int idx = y * W + x;
        while (g[idx] > 0 && !edge[idx]) {
            edge[idx] = 1;
            outX.push_back((uint32_t)x);
            outY.push_back((uint32_t)y);
            int shouldGo = 0;
            if (dir[idx] == kHorizontal) {
                if (lastDirection == kUp || lastDirection == kDown) shouldGo = (x > lastX) ? kRight : kLeft;
                lastX = x;
                lastY = y;
                if (lastDirection == kRight || shouldGo == kRight) {
                    if (x == W - 1 || y == 0 || y == H - 1) break;
                    const uint8_t g1 = (uint8_t)g[idx - W + 1], g2 = (uint8_t)g[idx + 1], g3 = (uint8_t)g[idx + W + 1];
                    if (g1 >= g2 && g1 >= g3) {
                        x = x + 1;
                        y = y - 1;
                    } else if (g3 >= g2 && g3 >= g1) {
                        x = x + 1;
                        y = y + 1;
                    } else
                        x = x + 1;
                    lastDirection = kRight;
                } else if (lastDirection == kLeft || shouldGo == kLeft) {
                    if (x == 0 || y == 0 || y == H - 1) break;
                    const uint8_t g1 = (uint8_t)g[idx - W - 1], g2 = (uint8_t)g[idx - 1], g3 = (uint8_t)g[idx + W - 1];
                    if (g1 >= g2 && g1 >= g3) {
                        x = x - 1;
                        y = y - 1;
                    } else if (g3 >= g2 && g3 >= g1) {
                        x = x - 1;
                        y = y + 1;
                    } else
                        x = x - 1;
                    lastDirection = kLeft;
                }
            } else {
                if (lastDirection == kRight || lastDirection == kLeft) shouldGo = (y > lastY) ? kDown : kUp;
                lastX = x;
                lastY = y;
                if (lastDirection == kDown || shouldGo == kDown) {
                    if (x == 0 || x == W - 1 || y == H - 1) break;
                    const uint8_t g1 = (uint8_t)g[idx + W + 1], g2 = (uint8_t)g[idx + W], g3 = (uint8_t)g[idx + W - 1];
                    if (g1 >= g2 && g1 >= g3) {
                        x = x + 1;
                        y = y + 1;
                    } else if (g3 >= g2 && g3 >= g1) {
                        x = x - 1;
                        y = y + 1;
                    } else
                        y = y + 1;
                    lastDirection = kDown;
                } else if (lastDirection == kUp || shouldGo == kUp) {
                    if (x == 0 || x == W - 1 || y == 0) break;
                    const uint8_t g1 = (uint8_t)g[idx - W + 1], g2 = (uint8_t)g[idx - W], g3 = (uint8_t)g[idx - W - 1];
                    if (g1 >= g2 && g1 >= g3) {
                        x = x + 1;
                        y = y - 1;
                    } else if (g3 >= g2 && g3 >= g1) {
                        x = x - 1;
                        y = y - 1;
                    } else
                        y = y - 1;
                    lastDirection = kUp;
                }
            }
            idx = y * W + x;
        }
    }

    /* binary_descriptor.cpp:2628-2714: first fit over minLineLen points starting at offsetS */
    double fit_initial(const std::vector<uint32_t> &xs, const std::vector<uint32_t> &ys, unsigned offsetS, double *eq)
    {
        const bool horiz = dir[ys[offsetS] * W + xs[offsetS]] == kHorizontal;
        /* A = [u_i ; 1], v_i : (u, v) = (x, y) for a horizontal line y = a x + b, (y, x) for a vertical one x = a y + b.
         * The float products of small integers are accumulated exactly (OpenCV's gemm accumulates float data in double). */
        double suu = 0, su = 0, suv = 0, sv = 0;
        for (int i = 0; i < minLineLen; i++) {
            const double u = (double)(float)(horiz ? xs[offsetS + i] : ys[offsetS + i]);
            const double v = (double)(float)(horiz ? ys[offsetS + i] : xs[offsetS + i]);
            suu += u * u;
            su += u;
            suv += u * v;
            sv += v;
        }
        ATA[0] = (float)suu;
        ATA[1] = (float)su;
        ATA[2] = (float)su;
        ATA[3] = (float)(double)minLineLen;
        ATV[0] = (float)suv;
        ATV[1] = (float)sv;
        const double coef = 1.0 / (double(ATA[0]) * double(ATA[3]) - double(ATA[1]) * double(ATA[2]));
        eq[0] = coef * (double(ATA[3]) * double(ATV[0]) - double(ATA[1]) * double(ATV[1]));
        eq[1] = coef * (double(ATA[0]) * double(ATV[1]) - double(ATA[2]) * double(ATV[0]));
        double fitError = 0;
        for (int i = 0; i < minLineLen; i++) {
            const double u = horiz ? xs[offsetS + i] : ys[offsetS + i], v = horiz ? ys[offsetS + i] : xs[offsetS + i];
            const double c = v - u * eq[0] - eq[1];
            fitError += c * c;
        }
        return std::sqrt(fitError);
    }

    /* binary_descriptor.cpp:2716-2787: re-estimate after adding points [newOffsetS, offsetE) */
    void fit_update(const std::vector<uint32_t> &xs, const std::vector<uint32_t> &ys, unsigned offsetS, unsigned newOffsetS, unsigned offsetE, double *eq)
    {
        const int length = (int)offsetE - (int)offsetS, newLength = (int)offsetE - (int)newOffsetS;
        if (length <= 0 || newLength <= 0) return;
        const bool horiz = dir[ys[offsetS] * W + xs[offsetS]] == kHorizontal;
        double suu = 0, su = 0, suv = 0, sv = 0;
        for (unsigned i = newOffsetS; i < offsetE; i++) {
            const double u = (double)(float)(horiz ? xs[i] : ys[i]);
            const double v = (double)(float)(horiz ? ys[i] : xs[i]);
            suu += u * u;
            su += u;
            suv += u * v;
            sv += v;
        }
        const float t[4] = {(float)suu, (float)su, (float)su, (float)(double)newLength};
        const float tv[2] = {(float)suv, (float)sv};
        for (int i = 0; i < 4; i++) ATA[i] = ATA[i] + t[i];
        for (int i = 0; i < 2; i++) ATV[i] = ATV[i] + tv[i];
        const double coef = 1.0 / (double(ATA[0]) * double(ATA[3]) - double(ATA[1]) * double(ATA[2]));
        eq[0] = coef * (double(ATA[3]) * double(ATV[0]) - double(ATA[1]) * double(ATV[1]));
        eq[1] = coef * (double(ATA[0]) * double(ATV[1]) - double(ATA[2]) * double(ATV[0]));
    }

    /* binary_descriptor.cpp:2789-2870 */
    bool validate(const std::vector<uint32_t> &xs, const std::vector<uint32_t> &ys, unsigned offsetS, unsigned offsetE, const double *lineEqu, float &direction)
    {
        const int n = (int)offsetE - (int)offsetS;
        int meanGradientX = 0, meanGradientY = 0;
        std::vector<double> pointDirection;
        pointDirection.reserve(n);
        for (int i = 0; i < n; i++) {
            const int index = ys[offsetS + i] * W + xs[offsetS + i];
            meanGradientX += dx[index];
            meanGradientY += dy[index];
            pointDirection.push_back(std::atan2(-(double)dx[index], (double)dy[index]));
        }
        const double ddx = std::fabs(lineEqu[1]), ddy = std::fabs(lineEqu[0]);
        if (meanGradientX == 0 && meanGradientY == 0) return false;
        /* `direction` is an out-parameter the caller declared once outside its loops: a quadrant test that matches nothing leaves the
         * previous line's value in place; the four tests below are exhaustive once (0,0) is excluded */
        if (meanGradientX > 0 && meanGradientY >= 0) direction = (float)std::atan2(-ddy, ddx);
        if (meanGradientX <= 0 && meanGradientY > 0) direction = (float)std::atan2(ddy, ddx);
        if (meanGradientX < 0 && meanGradientY <= 0) direction = (float)std::atan2(ddy, -ddx);
        if (meanGradientX >= 0 && meanGradientY < 0) direction = (float)std::atan2(-ddy, -ddx);
        if (std::fabs(direction) < 0.15 || M_PI - std::fabs(direction) < 0.15) {
            if (std::fabs(lineEqu[2]) < 10 || std::fabs(H - std::fabs(lineEqu[2])) < 10) return false;
        }
        if (std::fabs(std::fabs(direction) - M_PI * 0.5) < 0.15) {
            if (std::fabs(lineEqu[2]) < 10 || std::fabs(W - std::fabs(lineEqu[2])) < 10) return false;
        }
        int k = 0;
        for (int i = 0; i < n; i++) {
            const double dis = std::fabs(direction - pointDirection[i]);
            if (std::fabs(2 * M_PI - dis) < 0.392699 || dis < 0.392699) k++;
        }
        return nfa(n, k, 0.125, logNT) > 0;
    }
};

}  // namespace

/* line_lbd_detect::detect_filter_lines with use_LSD = false, one octave.  Optional stage outputs for pinning / GPU parity. */
/* kl_out (optional): the key-line fields of every KEPT line that the descriptor side reads (detectImpl :526-540): cap x
 * {direction, lineLength, numOfPixels} as three floats (numOfPixels is a small integer); same order as lines_out */
static int edl_detect_impl(const uint8_t *img, int w, int h, int stride, int channels, float line_length_thres, float *lines_out, int cap,
                           float *raw_lines, int cap_raw, int *n_raw_out, uint8_t *blur_out, int16_t *dx_out, int16_t *dy_out,
                           int16_t *g_out, uint8_t *dir_out, int32_t *anchors_out, int *n_anchors_out, uint8_t *edge_out, float *kl_out)
{
    if (n_raw_out) *n_raw_out = 0;
    if (n_anchors_out) *n_anchors_out = 0;
    /* detectImpl :486-492 */
    std::vector<uint8_t> gray((size_t)w * h);
    if (channels != 1)
        orc_bgr2gray(img, w, h, stride, gray.data(), w, 1);
    else
        for (int y = 0; y < h; y++) std::memcpy(&gray[(size_t)y * w], img + (size_t)y * stride, w);
    /* OctaveKeyLines :811-812: increaseSigma = sqrt(1 - 0) = 1, ksize 5 */
    std::vector<uint8_t> blur;
    gaussian5_u8(gray, w, h, blur);
    if (blur_out) std::memcpy(blur_out, blur.data(), blur.size());

    Ed E;
    E.W = w;
    E.H = h;
    const size_t npx = (size_t)w * h;
    E.dx.resize(npx);
    E.dy.resize(npx);
    E.g.resize(npx);
    E.dir.resize(npx);
    E.edge.assign(npx, 0);
    /* EdgeDrawing :1617-1629 */
    const int gradienThreshold = 80, anchorThreshold = 8, scanIntervals = 2;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            auto px = [&](int yy, int xx) -> int { return blur[(size_t)reflect101(yy, h) * w + reflect101(xx, w)]; };
            const int gx = (px(y - 1, x + 1) + 2 * px(y, x + 1) + px(y + 1, x + 1)) - (px(y - 1, x - 1) + 2 * px(y, x - 1) + px(y + 1, x - 1));
            const int gy = (px(y + 1, x - 1) + 2 * px(y + 1, x) + px(y + 1, x + 1)) - (px(y - 1, x - 1) + 2 * px(y - 1, x) + px(y - 1, x + 1));
            const int ax = std::abs(gx), ay = std::abs(gy), sum = ax + ay;
            const size_t i = (size_t)y * w + x;
            E.dx[i] = (int16_t)gx;
            E.dy[i] = (int16_t)gy;
            E.g[i] = (int16_t)div4_half_even(sum > gradienThreshold + 1 ? sum : 0);
            E.dir[i] = (ax < ay) ? 255 : 0;
        }
    if (dx_out) std::memcpy(dx_out, E.dx.data(), npx * 2);
    if (dy_out) std::memcpy(dy_out, E.dy.data(), npx * 2);
    if (g_out) std::memcpy(g_out, E.g.data(), npx * 2);
    if (dir_out) std::memcpy(dir_out, E.dir.data(), npx);

    /* anchors :1640-1666, scanned column by column */
    std::vector<uint32_t> ancX, ancY;
    for (unsigned ww = 1; ww + 1 < (unsigned)w; ww += scanIntervals)
        for (unsigned hh = 1; hh + 1 < (unsigned)h; hh += scanIntervals) {
            const int idx = hh * w + ww;
            if (E.dir[idx] == kHorizontal) {
                if (E.g[idx] >= E.g[idx - w] + anchorThreshold && E.g[idx] >= E.g[idx + w] + anchorThreshold) {
                    ancX.push_back(ww);
                    ancY.push_back(hh);
                }
            } else {
                if (E.g[idx] >= E.g[idx - 1] + anchorThreshold && E.g[idx] >= E.g[idx + 1] + anchorThreshold) {
                    ancX.push_back(ww);
                    ancY.push_back(hh);
                }
            }
        }
    if (n_anchors_out) *n_anchors_out = (int)ancX.size();
    if (anchors_out)
        for (size_t i = 0; i < ancX.size(); i++) anchors_out[i] = (int32_t)(ancY[i] * w + ancX[i]);

    /* smart routing :1684-2326 */
    const unsigned edgePixelArraySize = (unsigned)(npx / 5), maxNumOfEdge = edgePixelArraySize / 20;
    if (ancX.size() > edgePixelArraySize) return -1;
    std::vector<uint32_t> fX, fY, sX, sY;
    std::vector<unsigned> fS, sS;
    int lastX = 0, lastY = 0;
    for (size_t i = 0; i < ancX.size(); i++) {
        const int x = ancX[i], y = ancY[i];
        const int idx = y * w + x;
        if (E.edge[idx]) continue;
        const unsigned f0 = (unsigned)fX.size(), s0 = (unsigned)sX.size();
        if (E.dir[idx] == kHorizontal) {
            E.walk(x, y, kRight, fX, fY, lastX, lastY);
            E.edge[idx] = 0;
            E.walk(x, y, kLeft, sX, sY, lastX, lastY);
        } else {
            E.walk(x, y, kDown, fX, fY, lastX, lastY);
            E.edge[idx] = 0;
            E.walk(x, y, kUp, sX, sY, lastX, lastY);
        }
        const int lenF = (int)fX.size() - (int)f0, lenS = (int)sX.size() - (int)s0;
        if (lenF + lenS < E.minLineLen + 1) { /* short edge, drop it (its pixels stay marked) */
            fX.resize(f0);
            fY.resize(f0);
            sX.resize(s0);
            sY.resize(s0);
        } else {
            fS.push_back(f0);
            sS.push_back(s0);
        }
    }
    if (edge_out) std::memcpy(edge_out, E.edge.data(), npx);
    const unsigned numEdges = (unsigned)fS.size();
    fS.push_back((unsigned)fX.size());
    sS.push_back((unsigned)sX.size());
    if (numEdges > maxNumOfEdge || fX.size() > edgePixelArraySize || sX.size() > edgePixelArraySize) return -1;

    /* chain assembly :2340-2373: first part reversed, then the second part without its first pixel (the anchor) */
    std::vector<uint32_t> cx, cy;
    std::vector<unsigned> sId;
    for (unsigned e = 0; e < numEdges; e++) {
        sId.push_back((unsigned)cx.size());
        for (int t = (int)fS[e + 1] - 1; t >= (int)fS[e]; t--) {
            cx.push_back(fX[t]);
            cy.push_back(fY[t]);
        }
        for (int t = (int)sS[e] + 1; t < (int)sS[e + 1]; t++) {
            cx.push_back(sX[t]);
            cy.push_back(sY[t]);
        }
    }
    sId.push_back((unsigned)cx.size());

    /* EDline :2379-2626 */
    std::vector<float> endpoints; /* 4 per line */
    std::vector<float> directions;
    std::vector<int> npixels; /* lines_.sId[k + 1] - lines_.sId[k] (:1077): the pixels of the fitted line */
    if (numEdges > 0) {
        std::vector<uint32_t> lx(cx.size()), ly(cx.size());
        E.logNT = 2.0 * (std::log10((double)w) + std::log10((double)h));
        double lineFitErr = 0, eq[2] = {0, 0};
        unsigned offsetInLineArray = 0, newOffsetS = 0;
        float direction = 0; /* see validate(): declared once, like the reference */
        for (unsigned edgeID = 0; edgeID < numEdges; edgeID++) {
            unsigned S = sId[edgeID];
            const unsigned Eend = sId[edgeID + 1];
            while (Eend > S + E.minLineLen) {
                while (Eend > S + E.minLineLen) {
                    lineFitErr = E.fit_initial(cx, cy, S, eq);
                    if (lineFitErr <= E.fitErrThr) break;
                    S += kSkipEdgePoint;
                }
                if (lineFitErr > E.fitErrThr) break;
                const unsigned lineStart = offsetInLineArray;
                double coef1 = 0;
                bool bExtended = true, bFirstTry = true;
                int numOfOutlier, tryTimes = 0;
                const bool horiz = E.dir[cy[S] * w + cx[S]] == kHorizontal;
                while (bExtended) {
                    tryTimes++;
                    if (bFirstTry) {
                        bFirstTry = false;
                        for (int i = 0; i < E.minLineLen; i++) {
                            lx[offsetInLineArray] = cx[S];
                            ly[offsetInLineArray++] = cy[S++];
                        }
                    } else
                        E.fit_update(lx, ly, lineStart, newOffsetS, offsetInLineArray, eq);
                    coef1 = horiz ? 1 / std::sqrt(eq[0] * eq[0] + 1) : 1 / std::sqrt(1 + eq[0] * eq[0]);
                    numOfOutlier = 0;
                    newOffsetS = offsetInLineArray;
                    while (Eend > S) {
                        const double d = horiz ? std::fabs(eq[0] * cx[S] - cy[S] + eq[1]) * coef1 : std::fabs(cx[S] - eq[0] * cy[S] - eq[1]) * coef1;
                        lx[offsetInLineArray] = cx[S];
                        ly[offsetInLineArray++] = cy[S++];
                        if (d > E.fitErrThr) {
                            numOfOutlier++;
                            if (numOfOutlier > 3) break;
                        } else
                            numOfOutlier = 0;
                    }
                    offsetInLineArray -= numOfOutlier;
                    S -= numOfOutlier;
                    if (!(offsetInLineArray - newOffsetS > 0 && tryTimes < kTryTime)) bExtended = false;
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
                if (E.validate(lx, ly, lineStart, offsetInLineArray, lineEqu, direction)) {
                    const double a1 = lineEqu[1] * lineEqu[1], a2 = lineEqu[0] * lineEqu[0], a3 = lineEqu[0] * lineEqu[1];
                    const double a4 = lineEqu[2] * lineEqu[0], a5 = lineEqu[2] * lineEqu[1];
                    unsigned Px = lx[lineStart], Py = ly[lineStart];
                    endpoints.push_back((float)(a1 * Px - a3 * Py - a4));
                    endpoints.push_back((float)(a2 * Py - a3 * Px - a5));
                    Px = lx[offsetInLineArray - 1];
                    Py = ly[offsetInLineArray - 1];
                    endpoints.push_back((float)(a1 * Px - a3 * Py - a4));
                    endpoints.push_back((float)(a2 * Py - a3 * Px - a5));
                    directions.push_back(direction);
                    npixels.push_back((int)(offsetInLineArray - lineStart));
                } else
                    offsetInLineArray = lineStart;
            }
        }
    }
    const int n_raw = (int)directions.size();
    if (n_raw_out) *n_raw_out = n_raw;

    /* OctaveKeyLines :862-886,1069-1139 (start / end ordering) + detectImpl KeyLine fill + filter_lines + keylines_to_mat */
    int n_out = 0;
    for (int k = 0; k < n_raw; k++) {
        const float s1 = endpoints[4 * k], s2 = endpoints[4 * k + 1], e1 = endpoints[4 * k + 2], e2 = endpoints[4 * k + 3];
        float ddx = std::fabs(s1 - e1), ddy = std::fabs(s2 - e2);
        const float lineLength = std::sqrt(ddx * ddx + ddy * ddy);
        const float direction = directions[k];
        ddx = e1 - s1;
        ddy = e2 - s2;
        bool shouldChange = false;
        if (direction >= -0.75 * M_PI && direction < -0.25 * M_PI && ddy > 0) shouldChange = true;
        if (direction >= -0.25 * M_PI && direction < 0.25 * M_PI && ddx < 0) shouldChange = true;
        if (direction >= 0.25 * M_PI && direction < 0.75 * M_PI && ddy < 0) shouldChange = true;
        if (((direction >= 0.75 * M_PI && direction < M_PI) || (direction >= -M_PI && direction < -0.75 * M_PI)) && ddx > 0) shouldChange = true;
        const float scale0 = 1;
        float sx, sy, ex, ey;
        if (shouldChange) {
            sx = scale0 * e1;
            sy = scale0 * e2;
            ex = scale0 * s1;
            ey = scale0 * s2;
        } else {
            sx = scale0 * s1;
            sy = scale0 * s2;
            ex = scale0 * e1;
            ey = scale0 * e2;
        }
        if (raw_lines && k < cap_raw) {
            raw_lines[4 * k + 0] = sx;
            raw_lines[4 * k + 1] = sy;
            raw_lines[4 * k + 2] = ex;
            raw_lines[4 * k + 3] = ey;
        }
        if (!(lineLength > line_length_thres)) continue;
        if (n_out < cap) {
            lines_out[4 * n_out + 0] = sx;
            lines_out[4 * n_out + 1] = sy;
            lines_out[4 * n_out + 2] = ex;
            lines_out[4 * n_out + 3] = ey;
            if (kl_out) {
                kl_out[3 * n_out + 0] = direction;
                kl_out[3 * n_out + 1] = lineLength;
                kl_out[3 * n_out + 2] = (float)npixels[k];
            }
        }
        n_out++;
    }
    return n_out;
}

extern "C" int edl_orc_detect(const uint8_t *img, int w, int h, int stride, int channels, float line_length_thres, float *lines_out, int cap,
                              float *raw_lines, int cap_raw, int *n_raw_out, uint8_t *blur_out, int16_t *dx_out, int16_t *dy_out,
                              int16_t *g_out, uint8_t *dir_out, int32_t *anchors_out, int *n_anchors_out, uint8_t *edge_out)
{
    return edl_detect_impl(img, w, h, stride, channels, line_length_thres, lines_out, cap, raw_lines, cap_raw, n_raw_out, blur_out, dx_out, dy_out, g_out,
                           dir_out, anchors_out, n_anchors_out, edge_out, nullptr);
}

/* the same, returning per kept line {direction, lineLength, numOfPixels} as well, and the Sobel maps the descriptor reads */
extern "C" int edl_orc_detect_keylines(const uint8_t *img, int w, int h, int stride, int channels, float line_length_thres, float *lines_out, float *kl_out,
                                       int cap, int16_t *dx_out, int16_t *dy_out)
{
    return edl_detect_impl(img, w, h, stride, channels, line_length_thres, lines_out, cap, nullptr, 0, nullptr, nullptr, dx_out, dy_out, nullptr, nullptr,
                           nullptr, nullptr, nullptr, kl_out);
}
