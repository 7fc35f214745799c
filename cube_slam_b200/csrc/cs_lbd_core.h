/* cs_lbd_core.h -- the arithmetic of the LBD line descriptor and of the descriptor matcher, written once for the device kernels
 * (cs_lbd.cu) and for a host build of the same functions (tests/lbd_core_host.cpp: the CPU test suite runs these very functions, thread
 * index by thread index, against the oracle, so the kernels' arithmetic is checked where there is no GPU).
 *
 * Replaces BinaryDescriptor::computeLBD (line_lbd/libs/binary_descriptor.cpp:1146-1509), binaryConversion (:405-416) and the nearest-code
 * search of BinaryDescriptorMatcher::match (binary_descriptor_matcher.cpp:196-262, Mihasher::query :637-756).
 *
 * Bit-exactness: every float operation is an IEEE + - * / sqrt in the reference's order (nvcc -fmad=false -prec-div=true
 * -prec-sqrt=true, g++ -ffp-contract=off); cos / sin of the line direction and the Gaussian weights come from the host's libm, as the
 * reference computes them (cs_lbd.cu).  The work of one line is cut where the reference's loops are independent:
 *   rows    the 63 rows of the support region are independent of each other (each walks the line's length, one gather per step) except
 *           for the start point, a chain of hID float subtractions that every row replays on its own;
 *   bands   each of the 9 x 8 band sums receives its rows in increasing row order, whatever the other sums do;
 *   finish  the two normalisations are short sequential sums over the 72 values (one thread);
 *   bits    32 bytes, one comparison byte each.
 */
#ifndef CS_LBD_CORE_H
#define CS_LBD_CORE_H

#include <stdint.h>

#if defined(__CUDACC__)
#define CS_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#define CS_HD inline
#endif

#define CS_LBD_BANDS 9
#define CS_LBD_BAND_WIDTH 7
#define CS_LBD_ROWS 63  /* heightOfLSP */
#define CS_LBD_DESC 72  /* NUM_OF_BANDS * 8 */
#define CS_LBD_BYTES 32

/* one key line, prepared on the host (cs_lbd.cu: lbd_prepare) */
struct CsLbdLine {
    float mid_x, mid_y; /* lineMiddlePointX / Y (:1237-1238) */
    float dl_x, dl_y;   /* dL = (cos, sin) of KeyLine::angle (:1251-1252) */
    int32_t length;     /* lengthOfLSP = (short) numOfPixels (:1233) */
    int32_t frame;      /* which frame's Sobel maps */
};

/* round() of the reference (:1282,1284): half away from zero.  x - trunc(x) is exact, so this is the same integer as roundf / round. */
CS_HD int cs_lbd_round(float x)
{
    const float t = truncf(x);
    const float fr = x - t;
    if (fr >= 0.5f) return (int)t + 1;
    if (fr <= -0.5f) return (int)t - 1;
    return (int)t;
}

/* Row hID of the support region (:1262-1331 up to the band sums): out = {pgdLRowSum, ngdLRowSum, pgdORowSum, ngdORowSum} after the global
 * weight F_g[hID].  dx / dy: the frame's Sobel maps, w x h int16. */
CS_HD void cs_lbd_row(const CsLbdLine &L, int hID, const int16_t *dx, const int16_t *dy, int w, int h, float coefG, float out[4])
{
    const short halfHeight = (CS_LBD_ROWS - 1) / 2;
    const short lengthOfLSP = (short)L.length;
    const short halfWidth = (short)((lengthOfLSP - 1) / 2);
    const float dL0 = L.dl_x, dL1 = L.dl_y, dO0 = -dL1, dO1 = dL0;
    float sCorX0 = -dL0 * halfWidth + dL1 * halfHeight + L.mid_x;
    float sCorY0 = -dL1 * halfWidth - dL0 * halfHeight + L.mid_y;
    for (int r = 0; r < hID; r++) { /* the start of row hID: hID steps along the normal, added one at a time like the reference */
        sCorX0 -= dL1;
        sCorY0 += dL0;
    }
    float sCorX = sCorX0, sCorY = sCorY0;
    float pgdL = 0, ngdL = 0, pgdO = 0, ngdO = 0;
    const int imageWidth = w - 1, imageHeight = h - 1;
    for (short wID = 0; wID < lengthOfLSP; wID++) {
        int xCor = cs_lbd_round(sCorX);
        xCor = xCor < 0 ? 0 : (xCor > imageWidth ? imageWidth : xCor);
        int yCor = cs_lbd_round(sCorY);
        yCor = yCor < 0 ? 0 : (yCor > imageHeight ? imageHeight : yCor);
        const float gx = (float)dx[(size_t)yCor * w + xCor], gy = (float)dy[(size_t)yCor * w + xCor];
        const float gDL = gx * dL0 + gy * dL1;
        const float gDO = gx * dO0 + gy * dO1;
        if (gDL > 0)
            pgdL += gDL;
        else
            ngdL -= gDL;
        if (gDO > 0)
            pgdO += gDO;
        else
            ngdO -= gDO;
        sCorX += dL0;
        sCorY += dL1;
    }
    out[0] = coefG * pgdL;
    out[1] = coefG * ngdL;
    out[2] = coefG * pgdO;
    out[3] = coefG * ngdO;
}

/* Band sum t = band * 8 + q over its rows in increasing row order (:1336-1377).  q: 0 pgdL, 1 ngdL, 2 pgdL2, 3 ngdL2, 4 pgdO, 5 ngdO,
 * 6 pgdO2, 7 ngdO2.  rows: 63 x 4 from cs_lbd_row.  coefL: F_l, 21 floats. */
CS_HD float cs_lbd_band_sum(int t, const float *rows, const float *coefL)
{
    const int band = t >> 3, q = t & 7;
    const int src = (q & 1) | ((q & 4) >> 1); /* 0 pgdL, 1 ngdL, 2 pgdO, 3 ngdO */
    const bool squared = (q & 2) != 0;
    int h0 = (band - 1) * CS_LBD_BAND_WIDTH, h1 = (band + 2) * CS_LBD_BAND_WIDTH;
    if (h0 < 0) h0 = 0;
    if (h1 > CS_LBD_ROWS) h1 = CS_LBD_ROWS;
    float acc = 0;
    for (int hID = h0; hID < h1; hID++) {
        const int rb = hID / CS_LBD_BAND_WIDTH, m = hID % CS_LBD_BAND_WIDTH;
        /* the row's own band takes F_l[m + 7]; the band above it (rb - 1) F_l[m + 14]; the band below it (rb + 1) F_l[m] */
        const float c = coefL[rb == band ? m + CS_LBD_BAND_WIDTH : (rb == band + 1 ? m + 2 * CS_LBD_BAND_WIDTH : m)];
        const float v = rows[hID * 4 + src];
        if (squared)
            acc += c * c * (v * v);
        else
            acc += c * v;
    }
    return acc;
}

/* mean / standard deviation of one band from its eight sums (:1392-1416) into des[band * 8 .. + 8) */
CS_HD void cs_lbd_band_stats(int band, const float *sums /* 72 */, float *des /* 72 */)
{
    const float invN2 = (float)(1.0 / (CS_LBD_BAND_WIDTH * 2.0)), invN3 = (float)(1.0 / (CS_LBD_BAND_WIDTH * 3.0));
    const float invN = (band == 0 || band == CS_LBD_BANDS - 1) ? invN2 : invN3;
    const float *s = sums + band * 8;
    float *d = des + band * 8;
    float temp = s[0] * invN;
    d[0] = temp;
    d[4] = sqrtf(s[2] * invN - temp * temp);
    temp = s[1] * invN;
    d[1] = temp;
    d[5] = sqrtf(s[3] * invN - temp * temp);
    temp = s[4] * invN;
    d[2] = temp;
    d[6] = sqrtf(s[6] * invN - temp * temp);
    temp = s[5] * invN;
    d[3] = temp;
    d[7] = sqrtf(s[7] * invN - temp * temp);
}

/* the two normalisations and the 0.4 clamp between them (:1418-1484), in place */
CS_HD void cs_lbd_finish(float *des /* 72 */)
{
    float tempM = 0, tempS = 0;
    for (int i = 0; i < CS_LBD_DESC; i += 8) {
        tempM += des[i] * des[i];
        tempM += des[i + 1] * des[i + 1];
        tempM += des[i + 2] * des[i + 2];
        tempM += des[i + 3] * des[i + 3];
        tempS += des[i + 4] * des[i + 4];
        tempS += des[i + 5] * des[i + 5];
        tempS += des[i + 6] * des[i + 6];
        tempS += des[i + 7] * des[i + 7];
    }
    tempM = 1 / sqrtf(tempM);
    tempS = 1 / sqrtf(tempS);
    for (int i = 0; i < CS_LBD_DESC; i += 8) {
        des[i] = des[i] * tempM;
        des[i + 1] = des[i + 1] * tempM;
        des[i + 2] = des[i + 2] * tempM;
        des[i + 3] = des[i + 3] * tempM;
        des[i + 4] = des[i + 4] * tempS;
        des[i + 5] = des[i + 5] * tempS;
        des[i + 6] = des[i + 6] * tempS;
        des[i + 7] = des[i + 7] * tempS;
    }
    for (int i = 0; i < CS_LBD_DESC; i++)
        if ((double)des[i] > 0.4) des[i] = (float)0.4;
    float temp = 0;
    for (int i = 0; i < CS_LBD_DESC; i++) temp += des[i] * des[i];
    temp = 1 / sqrtf(temp);
    for (int i = 0; i < CS_LBD_DESC; i++) des[i] = des[i] * temp;
}

/* byte `comb` of the binary descriptor (computeImpl :757-770 over binaryConversion): bands (a, b) of the table at :74-107 */
CS_HD uint8_t cs_lbd_byte(int comb, const float *des)
{
    /* the pairs (i, j), i < j, of {0..8} that the reference lists: packed as i * 16 + j */
    const uint8_t pairs[CS_LBD_BYTES] = {0x01, 0x02, 0x03, 0x04, 0x05, 0x06, 0x12, 0x13, 0x14, 0x15, 0x16, 0x23, 0x24, 0x25, 0x26, 0x27,
                                         0x28, 0x34, 0x35, 0x36, 0x37, 0x38, 0x45, 0x46, 0x47, 0x48, 0x56, 0x57, 0x58, 0x67, 0x68, 0x78};
    const float *f1 = des + 8 * (pairs[comb] >> 4), *f2 = des + 8 * (pairs[comb] & 15);
    unsigned r = 0;
    for (int i = 0; i < 8; i++)
        if (f1[i] > f2[i]) r += 1u << i;
    return (uint8_t)r;
}

/* ---- matcher.  Multi-index hashing over 32 one-byte substrings with K = 1 returns, of the train codes at the smallest Hamming distance,
 * the one its search meets first: search radius s = 0 .. 4 outermost, then substring k = 0 .. 31, then the s-bit xor patterns in the order
 * Mihasher::query flips them -- which is increasing numeric value of the pattern (its enumeration, :675-735, moves the lowest movable one
 * first; tests/test_lbd_host_core.py checks the order against a run of that enumeration) -- then bucket order (= train index, populate()
 * appends).  A code none of whose bytes is within 4 bits of the query's is never met.  The key orders
 * (distance, s, k, pattern, train index); ~0 = never met. */
#define CS_LBD_KEY_DIST(key) ((int)((key) >> 48))
#define CS_LBD_KEY_TRAIN(key) ((uint32_t)((key)&0xffffffffull))
CS_HD uint64_t cs_lbd_match_key(const uint32_t *q /* 8 words */, const uint32_t *t /* 8 words */, uint32_t train_index)
{
    unsigned d = 0, smin = 9, kmin = 0, xmin = 0;
    for (int wd = 0; wd < 8; wd++) {
        const uint32_t x = q[wd] ^ t[wd];
        for (int b = 0; b < 4; b++) {
            const unsigned xb = (x >> (8 * b)) & 255u;
            unsigned s = xb - ((xb >> 1) & 0x55u);
            s = (s & 0x33u) + ((s >> 2) & 0x33u);
            s = (s + (s >> 4)) & 0x0fu;
            d += s;
            if (s < smin) {
                smin = s;
                kmin = (unsigned)(wd * 4 + b);
                xmin = xb;
            }
        }
    }
    if (smin > 4) return ~0ull;
    return ((uint64_t)d << 48) | ((uint64_t)smin << 45) | ((uint64_t)kmin << 40) | ((uint64_t)xmin << 32) | (uint64_t)train_index;
}

#endif /* CS_LBD_CORE_H */
