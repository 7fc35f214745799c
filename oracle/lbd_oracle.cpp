/*
 * oracle/lbd_oracle.cpp -- CPU ORACLE for the descriptor / matcher half of line_lbd_detect.
 * TEST INFRASTRUCTURE ONLY.
 *
 * Restates, without OpenCV, for one octave (the class keeps octave 0 only, line_lbd_allclass.cpp:200-207):
 *   line_lbd/class/line_lbd_allclass.cpp:224-272            detect_descrip_lines (both overloads): detect_raw_lines, lbd->compute, octave / length filter
 *   line_lbd/class/line_lbd_allclass.cpp:285-339            detect_descrip_lines_octaves: start / end swap, angle normalisation (lbd_orc_order_keylines)
 *   line_lbd/class/line_lbd_allclass.cpp:341-356            match_line_descrip
 *   line_lbd/libs/LSDDetector.cpp:226-250                   the KeyLine fields of the LSD flavour (length, LineIterator count, angle, size, response)
 *   line_lbd/libs/binary_descriptor.cpp:526-545             the KeyLine fields of the EDLines flavour (direction, numOfPixels from the fitted chain)
 *   line_lbd/libs/binary_descriptor.cpp:140-179             BinaryDescriptor(): the Gaussian weights F_g (63 rows) and F_l (3 x 7 rows)
 *   line_lbd/libs/binary_descriptor.cpp:352-402             computeGaussianPyramid + computeSobel, octave 0: GaussianBlur(5 x 5, sigma 1), Sobel 3 x 3 to 16S
 *   line_lbd/libs/binary_descriptor.cpp:603-790             computeImpl: one descriptor row per key line, binaryConversion over the 32 band pairs
 *   line_lbd/libs/binary_descriptor.cpp:1146-1509           computeLBD
 *   line_lbd/libs/binary_descriptor_matcher.cpp:196-262     BinaryDescriptorMatcher::match(query, train): Mihasher(256, 32), K = 1
 *   line_lbd/libs/binary_descriptor_matcher.cpp:598-756     Mihasher::batchquery / query (multi-index hashing: which of several nearest codes comes first)
 *
 * Overload resolution that decides last bits (the reference writes unqualified sqrt / cos / round inside namespace cv::line_descriptor):
 *   sqrt -> std::sqrt(float): OpenCV's cvstd.hpp puts `using std::sqrt` (and exp, pow, log, abs, min, max, swap) into namespace cv, so
 *           `1 / sqrt(tempM)` is a float division by a float root (oracle/ref/minicv.hpp carries the same using-declarations);
 *   cos, sin, round, atan2, fabs -> no such name in namespace cv, so the global ones: with <math.h> included anywhere in the translation unit
 *           (libstdc++'s <math.h> adds `using std::cos` ... to the global namespace) a float argument picks the FLOAT overload -- cosf, sinf,
 *           roundf, atan2f.  That is how the stand-in build of the reference resolves them (oracle/ref/linelbd_ref.cpp, which says so) and what
 *           this file and the product's host side follow; a build of the reference in which only <cmath> is visible would call the double
 *           functions and round the result to float, which differs from cosf / sinf in the last bit for a small share of the angles.
 *
 * PARITY: PINNED to the reference.  oracle/_ref/liblinelbd_ref.so is the reference's own line_lbd_allclass.cpp + binary_descriptor.cpp +
 * binary_descriptor_matcher.cpp (+ lsd.cpp, LSDDetector.cpp), compiled from /root/reference; tests/test_oracle_ref_lbd.py requires identical
 * key lines, identical 32-byte descriptors, identical 72-float descriptors (==) and identical matches on fixtures, synthetic frames and
 * random codes with planted ties.
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

extern "C" void orc_bgr2gray(const uint8_t *bgr, int w, int h, int stride, uint8_t *gray, int gstride, int fixed15);
extern "C" void edl_orc_gaussian5_u8(const uint8_t *src, int w, int h, uint8_t *dst);
extern "C" int lsd_orc_detect(const uint8_t *img, int w, int h, int stride, int channels, float line_length_thres, float *lines_out, int cap,
                              float *raw_lines, int cap_raw, int *n_raw_out, double *scaled_out, double *modgrad_out, double *angles_out,
                              int32_t *list_out, int *list_len, int refine_mode);
extern "C" int edl_orc_detect_keylines(const uint8_t *img, int w, int h, int stride, int channels, float line_length_thres, float *lines_out, float *kl_out,
                                       int cap, int16_t *dx_out, int16_t *dy_out);

/* the KeyLine fields anything downstream reads (descriptor.hpp:104-172), octave 0: startPoint == sPointInOctave */
struct lbd_keyline {
    float sx, sy, ex, ey;
    float angle;       /* KeyLine::angle */
    float line_length; /* KeyLine::lineLength */
    float response;
    float size;
    int32_t num_pixels; /* KeyLine::numOfPixels */
    int32_t class_id;
};

namespace {

const int kBands = 9, kBandWidth = 7, kRows = kBands * kBandWidth, kDesc = kBands * 8;

/* binary_descriptor.cpp:74-107 */
const int kCombinations[32][2] = {{0, 1}, {0, 2}, {0, 3}, {0, 4}, {0, 5}, {0, 6}, {1, 2}, {1, 3}, {1, 4}, {1, 5}, {1, 6}, {2, 3}, {2, 4}, {2, 5}, {2, 6}, {2, 7},
                                  {2, 8}, {3, 4}, {3, 5}, {3, 6}, {3, 7}, {3, 8}, {4, 5}, {4, 6}, {4, 7}, {4, 8}, {5, 6}, {5, 7}, {5, 8}, {6, 7}, {6, 8}, {7, 8}};

inline int reflect101(int p, int n)
{
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = (p < 0) ? -p : 2 * n - 2 - p;
    return p;
}

/* :140-179.  Integer divisions as written: u = (21 - 1) / 2 = 10, sigma = (14 + 1) / 2 = 7; u = (63 - 1) / 2 = 31 = sigma */
void gauss_tables(double *G, double *L)
{
    double u = (kBandWidth * 3 - 1) / 2;
    double sigma = (kBandWidth * 2 + 1) / 2;
    double invsigma2 = -1 / (2 * sigma * sigma);
    for (int i = 0; i < kBandWidth * 3; i++) {
        const double dis = i - u;
        L[i] = std::exp(dis * dis * invsigma2);
    }
    u = (kBands * kBandWidth - 1) / 2;
    sigma = u;
    invsigma2 = -1 / (2 * sigma * sigma);
    for (int i = 0; i < kRows; i++) {
        const double dis = i - u;
        G[i] = std::exp(dis * dis * invsigma2);
    }
}

/* computeSobel :374-398 for octave 0 */
void sobel_maps(const uint8_t *img, int w, int h, int stride, int channels, std::vector<int16_t> &dx, std::vector<int16_t> &dy)
{
    std::vector<uint8_t> gray((size_t)w * h), blur((size_t)w * h);
    if (channels != 1)
        orc_bgr2gray(img, w, h, stride, gray.data(), w, 1);
    else
        for (int y = 0; y < h; y++) std::memcpy(&gray[(size_t)y * w], img + (size_t)y * stride, w);
    edl_orc_gaussian5_u8(gray.data(), w, h, blur.data());
    dx.resize((size_t)w * h);
    dy.resize((size_t)w * h);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            auto px = [&](int yy, int xx) -> int { return blur[(size_t)reflect101(yy, h) * w + reflect101(xx, w)]; };
            dx[(size_t)y * w + x] = (int16_t)((px(y - 1, x + 1) + 2 * px(y, x + 1) + px(y + 1, x + 1)) - (px(y - 1, x - 1) + 2 * px(y, x - 1) + px(y + 1, x - 1)));
            dy[(size_t)y * w + x] = (int16_t)((px(y + 1, x - 1) + 2 * px(y + 1, x) + px(y + 1, x + 1)) - (px(y - 1, x - 1) + 2 * px(y - 1, x) + px(y - 1, x + 1)));
        }
}

/* computeLBD :1146-1509 for one line; desc: 72 floats */
void lbd_one_line(const int16_t *pdxImg, const int16_t *pdyImg, int w, int h, const lbd_keyline &kl, const double *gaussCoefG, const double *gaussCoefL,
                  float *desVec)
{
    const short heightOfLSP = (short)(kBandWidth * kBands);
    const short descriptor_size = kDesc;
    float pgdLRowSum, ngdLRowSum, pgdL2RowSum, ngdL2RowSum, pgdORowSum, ngdORowSum, pgdO2RowSum, ngdO2RowSum;
    float pgdLBandSum[kBands] = {0}, ngdLBandSum[kBands] = {0}, pgdL2BandSum[kBands] = {0}, ngdL2BandSum[kBands] = {0};
    float pgdOBandSum[kBands] = {0}, ngdOBandSum[kBands] = {0}, pgdO2BandSum[kBands] = {0}, ngdO2BandSum[kBands] = {0};
    const short halfHeight = (heightOfLSP - 1) / 2;
    const short realWidth = (short)w, imageWidth = realWidth - 1, imageHeight = (short)(h - 1);
    const short lengthOfLSP = (short)kl.num_pixels;
    const short halfWidth = (lengthOfLSP - 1) / 2;
    const float lineMiddlePointX = (float)(0.5 * (kl.sx + kl.ex));
    const float lineMiddlePointY = (float)(0.5 * (kl.sy + kl.ey));
    float dL[2], dO[2];
    dL[0] = std::cos(kl.angle); /* cosf / sinf: see the header */
    dL[1] = std::sin(kl.angle);
    dO[0] = -dL[1];
    dO[1] = dL[0];
    float sCorX0 = -dL[0] * halfWidth + dL[1] * halfHeight + lineMiddlePointX;
    float sCorY0 = -dL[1] * halfWidth - dL[0] * halfHeight + lineMiddlePointY;
    for (short hID = 0; hID < heightOfLSP; hID++) {
        float sCorX = sCorX0, sCorY = sCorY0;
        pgdLRowSum = ngdLRowSum = pgdORowSum = ngdORowSum = 0;
        for (short wID = 0; wID < lengthOfLSP; wID++) {
            short tempCor = (short)std::round(sCorX);
            const short xCor = (tempCor < 0) ? 0 : (tempCor > imageWidth) ? imageWidth : tempCor;
            tempCor = (short)std::round(sCorY);
            const short yCor = (tempCor < 0) ? 0 : (tempCor > imageHeight) ? imageHeight : tempCor;
            const short dx = pdxImg[yCor * realWidth + xCor], dy = pdyImg[yCor * realWidth + xCor];
            const float gDL = dx * dL[0] + dy * dL[1];
            const float gDO = dx * dO[0] + dy * dO[1];
            if (gDL > 0)
                pgdLRowSum += gDL;
            else
                ngdLRowSum -= gDL;
            if (gDO > 0)
                pgdORowSum += gDO;
            else
                ngdORowSum -= gDO;
            sCorX += dL[0];
            sCorY += dL[1];
        }
        sCorX0 -= dL[1];
        sCorY0 += dL[0];
        float coefInGaussion = (float)gaussCoefG[hID];
        pgdLRowSum = coefInGaussion * pgdLRowSum;
        ngdLRowSum = coefInGaussion * ngdLRowSum;
        pgdL2RowSum = pgdLRowSum * pgdLRowSum;
        ngdL2RowSum = ngdLRowSum * ngdLRowSum;
        pgdORowSum = coefInGaussion * pgdORowSum;
        ngdORowSum = coefInGaussion * ngdORowSum;
        pgdO2RowSum = pgdORowSum * pgdORowSum;
        ngdO2RowSum = ngdORowSum * ngdORowSum;
        auto add = [&](short bandID, float c) {
            pgdLBandSum[bandID] += c * pgdLRowSum;
            ngdLBandSum[bandID] += c * ngdLRowSum;
            pgdL2BandSum[bandID] += c * c * pgdL2RowSum;
            ngdL2BandSum[bandID] += c * c * ngdL2RowSum;
            pgdOBandSum[bandID] += c * pgdORowSum;
            ngdOBandSum[bandID] += c * ngdORowSum;
            pgdO2BandSum[bandID] += c * c * pgdO2RowSum;
            ngdO2BandSum[bandID] += c * c * ngdO2RowSum;
        };
        short bandID = (short)(hID / kBandWidth);
        add(bandID, (float)gaussCoefL[hID % kBandWidth + kBandWidth]);
        bandID--;
        if (bandID >= 0) add(bandID, (float)gaussCoefL[hID % kBandWidth + 2 * kBandWidth]);
        bandID = bandID + 2;
        if (bandID < kBands) add(bandID, (float)gaussCoefL[hID % kBandWidth]);
    }
    const float invN2 = (float)(1.0 / (kBandWidth * 2.0)), invN3 = (float)(1.0 / (kBandWidth * 3.0));
    for (short bandID = 0; bandID < kBands; bandID++) {
        const float invN = (bandID == 0 || bandID == kBands - 1) ? invN2 : invN3;
        const short desID = bandID * 8;
        float temp = pgdLBandSum[bandID] * invN;
        desVec[desID] = temp;
        desVec[desID + 4] = std::sqrt(pgdL2BandSum[bandID] * invN - temp * temp);
        temp = ngdLBandSum[bandID] * invN;
        desVec[desID + 1] = temp;
        desVec[desID + 5] = std::sqrt(ngdL2BandSum[bandID] * invN - temp * temp);
        temp = pgdOBandSum[bandID] * invN;
        desVec[desID + 2] = temp;
        desVec[desID + 6] = std::sqrt(pgdO2BandSum[bandID] * invN - temp * temp);
        temp = ngdOBandSum[bandID] * invN;
        desVec[desID + 3] = temp;
        desVec[desID + 7] = std::sqrt(ngdO2BandSum[bandID] * invN - temp * temp);
    }
    float tempM = 0, tempS = 0;
    for (int i = 0; i < kDesc; i += 8) {
        tempM += desVec[i] * desVec[i];
        tempM += desVec[i + 1] * desVec[i + 1];
        tempM += desVec[i + 2] * desVec[i + 2];
        tempM += desVec[i + 3] * desVec[i + 3];
        tempS += desVec[i + 4] * desVec[i + 4];
        tempS += desVec[i + 5] * desVec[i + 5];
        tempS += desVec[i + 6] * desVec[i + 6];
        tempS += desVec[i + 7] * desVec[i + 7];
    }
    tempM = 1 / std::sqrt(tempM);
    tempS = 1 / std::sqrt(tempS);
    for (int i = 0; i < kDesc; i += 8) {
        for (int j = 0; j < 4; j++) desVec[i + j] = desVec[i + j] * tempM;
        for (int j = 4; j < 8; j++) desVec[i + j] = desVec[i + j] * tempS;
    }
    for (short i = 0; i < descriptor_size; i++)
        if (desVec[i] > 0.4) desVec[i] = (float)0.4;
    float temp = 0;
    for (short i = 0; i < descriptor_size; i++) temp += desVec[i] * desVec[i];
    temp = 1 / std::sqrt(temp);
    for (short i = 0; i < descriptor_size; i++) desVec[i] = desVec[i] * temp;
}

/* Mihasher::query's enumeration of the bit strings with s ones among 8 (:670-735), as the order they are looked up in */
void mih_pattern_rank(int rank[256])
{
    for (int i = 0; i < 256; i++) rank[i] = -1;
    const int curb = 8;
    for (int s = 0; s <= 4; s++) {
        int power[16];
        uint64_t bitstr = 0;
        for (int i = 0; i < s; i++) power[i] = i;
        power[s] = curb + 1;
        int bit = s - 1, r = 0;
        while (true) {
            if (bit != -1) {
                bitstr ^= (power[bit] == bit) ? (uint64_t)1 << power[bit] : (uint64_t)3 << (power[bit] - 1);
                power[bit]++;
                bit--;
            } else {
                if (bitstr < 256 && rank[bitstr] < 0) rank[bitstr] = r;
                r++;
                while (++bit < s && power[bit] == power[bit + 1] - 1) {
                    bitstr ^= (uint64_t)1 << (power[bit] - 1);
                    power[bit] = bit;
                }
                if (bit == s) break;
            }
        }
    }
}

}  // namespace

/* number of pixels cv::LineIterator(img, Point2f, Point2f) reports for two points inside the image: Point2f -> Point rounds half to even
 * (saturate_cast<int>(float) == cvRound), 8-connected: max(|dx|, |dy|) + 1 */
static int line_iterator_count(float x1, float y1, float x2, float y2, int w, int h)
{
    auto clampi = [](long v, int n) { return (int)(v < 0 ? 0 : (v >= n ? n - 1 : v)); };
    const int ix1 = clampi(std::lrint(x1), w), iy1 = clampi(std::lrint(y1), h), ix2 = clampi(std::lrint(x2), w), iy2 = clampi(std::lrint(y2), h);
    return std::max(std::abs(ix2 - ix1), std::abs(iy2 - iy1)) + 1;
}

/* LSDDetector::detectImpl :226-250 from the rows detect_filter_lines returns (the clamped extremes of the kept lines, octave scale 1) */
extern "C" void lbd_orc_keylines_from_lsd(const float *lines, int n, int w, int h, lbd_keyline *out)
{
    for (int k = 0; k < n; k++) {
        const float *e = lines + 4 * k;
        lbd_keyline &kl = out[k];
        kl.sx = e[0];
        kl.sy = e[1];
        kl.ex = e[2];
        kl.ey = e[3];
        kl.line_length = (float)std::sqrt(std::pow(e[0] - e[2], 2) + std::pow(e[1] - e[3], 2));
        kl.num_pixels = line_iterator_count(e[0], e[1], e[2], e[3], w, h);
        kl.angle = std::atan2(kl.ey - kl.sy, kl.ex - kl.sx); /* atan2f: see the header */
        kl.size = (kl.ex - kl.sx) * (kl.ey - kl.sy);
        kl.response = kl.line_length / std::max(w, h);
        kl.class_id = k;
    }
}

/* detect_descrip_lines(gray_img, keylines_out, line_descrips) (line_lbd_allclass.cpp:253-272): the kept key lines, in order.
 * Descriptors are per line, so filtering before describing returns the rows the reference keeps. */
extern "C" int lbd_orc_detect_keylines(const uint8_t *img, int w, int h, int stride, int channels, int use_lsd, float line_length_thres, lbd_keyline *out, int cap)
{
    std::vector<float> lines((size_t)cap * 4);
    if (use_lsd) {
        const int n = lsd_orc_detect(img, w, h, stride, channels, line_length_thres, lines.data(), cap, nullptr, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 2);
        if (n < 0 || n > cap) return n < 0 ? n : -3;
        lbd_orc_keylines_from_lsd(lines.data(), n, w, h, out);
        /* class_id counts every line that passed the border test, kept or not (LSDDetector.cpp:245); the descriptor only needs it unique */
        return n;
    }
    std::vector<float> kl((size_t)cap * 3);
    const int n = edl_orc_detect_keylines(img, w, h, stride, channels, line_length_thres, lines.data(), kl.data(), cap, nullptr, nullptr);
    if (n < 0 || n > cap) return n < 0 ? n : -3;
    for (int k = 0; k < n; k++) {
        lbd_keyline &o = out[k];
        o.sx = lines[4 * k];
        o.sy = lines[4 * k + 1];
        o.ex = lines[4 * k + 2];
        o.ey = lines[4 * k + 3];
        o.angle = kl[3 * k];
        o.line_length = kl[3 * k + 1];
        o.num_pixels = (int32_t)kl[3 * k + 2];
        o.size = (o.ex - o.sx) * (o.ey - o.sy);            /* binary_descriptor.cpp:543 */
        o.response = o.line_length / std::max(w, h);        /* :544 */
        o.class_id = k;
    }
    return n;
}

/* detect_descrip_lines_octaves' reordering (line_lbd_allclass.cpp:321-330), one octave: start x <= end x, angle folded into [-pi/2, pi/2] */
extern "C" void lbd_orc_order_keylines(lbd_keyline *kl, int n)
{
    const double PI = 3.14159265; /* line_lbd_allclass.cpp:19, a double: the comparison and the fold are done in double (:272-281) */
    for (int i = 0; i < n; i++)
        if (kl[i].sx > kl[i].ex) {
            std::swap(kl[i].sx, kl[i].ex);
            std::swap(kl[i].sy, kl[i].ey);
            const float a = kl[i].angle;
            if (a > PI / 2)
                kl[i].angle = (float)(a - PI);
            else if (a < -PI / 2)
                kl[i].angle = (float)(a + PI);
        }
}

/* BinaryDescriptor::compute(image, keylines, descriptors [, returnFloatDescr]) (:587-592, 603-790): desc n x 32 bytes, fdesc (optional) n x 72 floats */
extern "C" int lbd_orc_compute(const uint8_t *img, int w, int h, int stride, int channels, const lbd_keyline *kl, int n, uint8_t *desc, float *fdesc)
{
    if (n <= 0) return 0; /* "Error: keypoint list is empty" */
    std::vector<int16_t> dx, dy;
    sobel_maps(img, w, h, stride, channels, dx, dy);
    double G[kRows], L[kBandWidth * 3];
    gauss_tables(G, L);
    for (int i = 0; i < n; i++) {
        float d[kDesc];
        lbd_one_line(dx.data(), dy.data(), w, h, kl[i], G, L, d);
        if (fdesc) std::memcpy(fdesc + (size_t)i * kDesc, d, sizeof d);
        if (desc)
            for (int comb = 0; comb < 32; comb++) { /* binaryConversion :405-416 */
                const float *f1 = &d[8 * kCombinations[comb][0]], *f2 = &d[8 * kCombinations[comb][1]];
                uint8_t r = 0;
                for (int b = 0; b < 8; b++)
                    if (f1[b] > f2[b]) r += (uint8_t)(1 << b);
                desc[(size_t)i * 32 + comb] = r;
            }
    }
    return n;
}

/* the weights, for the product's host side to be checked against (it builds its own with libm) */
extern "C" void lbd_orc_gauss_tables(double *G63, double *L21) { gauss_tables(G63, L21); }
extern "C" void lbd_orc_pattern_rank(int32_t *rank256)
{
    int r[256];
    mih_pattern_rank(r);
    for (int i = 0; i < 256; i++) rank256[i] = r[i];
}

/* line_lbd_detect::match_line_descrip (:341-356) over BinaryDescriptorMatcher::match(query, train) (:196-262).
 * Multi-index hashing with m = 32 one-byte substrings and K = 1 returns, of the codes at the smallest Hamming distance, the one its
 * search meets first: radius s = 0 .. 4 outermost, then substring k = 0 .. 31, then the s-bit patterns in the order query() flips them,
 * then bucket order (= train index, populate() appends).  A code none of whose bytes is within 4 bits of the query's is never met; a
 * query that meets nothing yields no DMatch; a nearest code further than D = 128 leaves trainIdx uninitialised in the reference
 * (results[] is never written) -- reported as -1 here.  Returns the number of matches with distance < thres, in query order. */
extern "C" int lbd_orc_match(const uint8_t *q, int nq, const uint8_t *t, int nt, float thres, int32_t *query_idx, int32_t *train_idx, float *dist)
{
    if (nq <= 0 || nt <= 0) return 0; /* "descriptors matrices cannot be void" */
    int rank[256];
    mih_pattern_rank(rank);
    int n_out = 0;
    for (int i = 0; i < nq; i++) {
        uint64_t best = ~0ull;
        for (int j = 0; j < nt; j++) {
            int d = 0, smin = 9, kmin = 0;
            for (int k = 0; k < 32; k++) {
                const int x = q[(size_t)i * 32 + k] ^ t[(size_t)j * 32 + k], s = __builtin_popcount(x);
                d += s;
                if (s < smin) {
                    smin = s;
                    kmin = k;
                }
            }
            if (smin > 4) continue; /* never looked up */
            const int x = q[(size_t)i * 32 + kmin] ^ t[(size_t)j * 32 + kmin];
            const uint64_t key = ((uint64_t)d << 47) | ((uint64_t)smin << 44) | ((uint64_t)kmin << 39) | ((uint64_t)rank[x] << 32) | (uint64_t)j;
            best = std::min(best, key);
        }
        if (best == ~0ull) continue;
        const int d = (int)(best >> 47);
        if (!((float)d < thres)) continue;
        query_idx[n_out] = i;
        train_idx[n_out] = d <= 128 ? (int32_t)(best & 0xffffffffu) : -1;
        dist[n_out] = (float)d;
        n_out++;
    }
    return n_out;
}
