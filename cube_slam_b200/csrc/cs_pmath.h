/*
 * cs_pmath.h -- atan2 from IEEE-754 double + - * / and comparisons only, so that it rounds the same on the host and on the device
 * (host: -ffp-contract=off, device: -fmad=false; no fused multiply-add on either side).
 *
 * Why: the reference takes atan2 from libm (object_3d_util.cpp:167-172,321,392,480).  Its results feed fuse_normalize_scores_v2, which sorts
 * the angle errors of all proposals of a box and cuts the sorted list at 2/3 (object_3d_util.cpp:504-520).  Proposals come in mirror pairs
 * whose angle errors are mathematically equal and differ in the last bits, so WHICH member of a pair lands inside the cut -- and with it
 * the kept set, its min / max and every normalised score of the box -- hangs on the last bit of atan2: CUDA's atan2 (<= 2 ulp) and
 * glibc's (< 1 ulp) split such pairs differently on a per cent or so of the boxes.  With one arithmetic definition of atan2 on both sides
 * the CUDA path and its CPU oracle agree bit for bit, and the remaining difference to a particular libm is stated (and measured in
 * tests/test_pmath.py: at most 1 ulp from glibc's) instead of hidden in a tolerance.
 *
 * The algorithm is the classic table-free reduction used by fdlibm-style libraries: |t| is reduced with the breakpoints 7/16, 11/16,
 * 19/16, 39/16 to atan(c) + atan((t - c) / (1 + t c)), c in {0.5, 1, 1.5, inf}, followed by an odd polynomial of degree 23 on the
 * reduced argument; atan2 adds the quadrant with pi split in two doubles.
 */
#ifndef CS_PMATH_H
#define CS_PMATH_H

#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define CS_PM_HD __host__ __device__ __forceinline__
#else
#define CS_PM_HD static inline
#endif

CS_PM_HD uint64_t cs_pm_bits(double v)
{
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(v);
#else
    uint64_t u;
    memcpy(&u, &v, 8);
    return u;
#endif
}

/* atan of a finite, non-negative argument */
CS_PM_HD double cs_pm_atan_pos(double x)
{
    const double hi[4] = {4.63647609000806093515e-01, 7.85398163397448278999e-01, 9.82793723247329054082e-01, 1.57079632679489655800e+00};
    const double lo[4] = {2.26987774529616870924e-17, 3.06161699786838301793e-17, 1.39033110312309984516e-17, 6.12323399573676603587e-17};
    const double a0 = 3.33333333333329318027e-01, a1 = -1.99999999998764832476e-01, a2 = 1.42857142725034663711e-01,
                 a3 = -1.11111104054623557880e-01, a4 = 9.09088713343650656196e-02, a5 = -7.69187620504482999495e-02,
                 a6 = 6.66107313738753120669e-02, a7 = -5.83357013379057348645e-02, a8 = 4.97687799461593236017e-02,
                 a9 = -3.65315727442169155270e-02, a10 = 1.62858201153657823623e-02;
    if (x >= 7.378697629483820646e19) return hi[3] + lo[3]; /* 2^66 */
    int id;
    if (x < 0.4375) {
        if (x < 1.862645149230957e-09) return x; /* 2^-29 */
        id = -1;
    } else if (x < 1.1875) {
        if (x < 0.6875) {
            id = 0;
            x = (2.0 * x - 1.0) / (2.0 + x);
        } else {
            id = 1;
            x = (x - 1.0) / (x + 1.0);
        }
    } else if (x < 2.4375) {
        id = 2;
        x = (x - 1.5) / (1.0 + 1.5 * x);
    } else {
        id = 3;
        x = -1.0 / x;
    }
    const double z = x * x, w = z * z;
    const double s1 = z * (a0 + w * (a2 + w * (a4 + w * (a6 + w * (a8 + w * a10)))));
    const double s2 = w * (a1 + w * (a3 + w * (a5 + w * (a7 + w * a9))));
    if (id < 0) return x - x * (s1 + s2);
    return hi[id] - ((x * (s1 + s2) - lo[id]) - x);
}

CS_PM_HD double cs_pm_atan2(double y, double x)
{
    const double pi = 3.1415926535897931160e+00, pi_lo = 1.2246467991473531772e-16, pi_o_2 = 1.5707963267948965580e+00,
                 pi_o_4 = 7.8539816339744827900e-01;
    if (x != x || y != y) return x + y;
    const uint64_t bx = cs_pm_bits(x), by = cs_pm_bits(y);
    const int m = (int)(by >> 63) | ((int)(bx >> 63) << 1); /* 1: y negative, 2: x negative (signed zeros count) */
    const uint64_t ax = bx & 0x7fffffffffffffffull, ay = by & 0x7fffffffffffffffull, inf = 0x7ff0000000000000ull;
    if (ay == 0) return (m & 2) ? ((m & 1) ? -pi : pi) : y;
    if (ax == 0) return (m & 1) ? -pi_o_2 : pi_o_2;
    if (ax == inf) {
        if (ay == inf) return (m & 2) ? ((m & 1) ? -3.0 * pi_o_4 : 3.0 * pi_o_4) : ((m & 1) ? -pi_o_4 : pi_o_4);
        return (m & 2) ? ((m & 1) ? -pi : pi) : ((m & 1) ? -0.0 : 0.0);
    }
    if (ay == inf) return (m & 1) ? -pi_o_2 : pi_o_2;
    const int k = (int)(ay >> 52) - (int)(ax >> 52); /* exponent of y / x, roughly */
    double z;
    if (k > 60)
        z = pi_o_2 + 0.5 * pi_lo;
    else if ((m & 2) && k < -60)
        z = 0.0;
    else {
        double q = y / x;
        if (q < 0) q = -q;
        z = cs_pm_atan_pos(q);
    }
    switch (m) {
        case 0: return z;
        case 1: return -z;
        case 2: return pi - (z - pi_lo);
        default: return (z - pi_lo) - pi;
    }
}

#endif /* CS_PMATH_H */
