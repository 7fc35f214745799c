/*
 * cs_geom.cuh -- FP64 device geometry of the cuboid sweep.  Compiled with -fmad=false so that
 * + - * / sqrt are IEEE-exact and match the CPU evaluation order of the reference
 * (detect_3d_cuboid/src/object_3d_util.cpp, box_proposal_detail.cpp).
 */
#ifndef CS_GEOM_CUH
#define CS_GEOM_CUH

#include "cs_pmath.h"
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "cs_internal.h"

#define CS_PI 3.14159265358979323846

struct D2 {
    double x, y;
};

__device__ __forceinline__ double g_min(double a, double b) { return (b < a) ? b : a; } /* std::min */
__device__ __forceinline__ double g_max(double a, double b) { return (a < b) ? b : a; } /* std::max */
__device__ __forceinline__ double g_norm(D2 a, D2 b)
{
    const double dx = a.x - b.x, dy = a.y - b.y;
    return sqrt(dx * dx + dy * dy);
}

/* matrix_utils.cpp:327-335 */
__device__ __forceinline__ double g_normalize_to_pi(double angle)
{
    if (angle > CS_PI / 2) return angle - CS_PI;
    if (angle < -CS_PI / 2) return angle + CS_PI;
    return angle;
}

/* object_3d_util.cpp:141-144 */
__device__ __forceinline__ bool g_inside(D2 p, double l, double t, double r, double b)
{
    return l <= p.x && p.x <= r && t <= p.y && p.y <= b;
}

/* object_3d_util.cpp:194-230, boundary is a vertical segment x = bx, y in [by0, by1] */
__device__ __forceinline__ D2 g_hit_vertical(D2 s, D2 e, double bx, double by0, double by1)
{
    D2 hit = {-1, -1};
    const double dx = e.x - s.x, dy = e.y - s.y;
    if (by0 == by1) { /* degenerate: also a horizontal edge */
        const double lambd = (by0 - s.y) / dy;
        if (lambd >= 0) {
            const double tx = s.x + lambd * dx, ty = s.y + lambd * dy;
            if (bx <= tx && tx <= bx) {
                hit.x = tx;
                hit.y = by0;
            }
            (void)ty;
        }
    }
    const double lambd = (bx - s.x) / dx;
    if (lambd >= 0) {
        const double tx = s.x + lambd * dx, ty = s.y + lambd * dy;
        if (by0 <= ty && ty <= by1) {
            hit.x = bx;
            hit.y = ty;
        }
        (void)tx;
    }
    return hit;
}

/* object_3d_util.cpp:194-230, boundary is a horizontal segment y = by, x in [bx0, bx1] */
__device__ __forceinline__ D2 g_hit_horizontal(D2 s, D2 e, double bx0, double bx1, double by)
{
    D2 hit = {-1, -1};
    const double dx = e.x - s.x, dy = e.y - s.y;
    {
        const double lambd = (by - s.y) / dy;
        if (lambd >= 0) {
            const double tx = s.x + lambd * dx;
            if (bx0 <= tx && tx <= bx1) {
                hit.x = tx;
                hit.y = by;
            }
        }
    }
    if (bx0 == bx1) { /* degenerate: also a vertical edge */
        const double lambd = (bx0 - s.x) / dx;
        if (lambd >= 0) {
            const double ty = s.y + lambd * dy;
            if (by <= ty && ty <= by) {
                hit.x = bx0;
                hit.y = ty;
            }
        }
    }
    return hit;
}

/* object_3d_util.cpp:233-252 with infinite_line = true */
__device__ __forceinline__ D2 g_intersect(D2 p1s, D2 p1e, D2 p2s, D2 p2e)
{
    const double X2_X1 = p1e.x - p1s.x, Y2_Y1 = p1e.y - p1s.y;
    const double X4_X3 = p2e.x - p2s.x, Y4_Y3 = p2e.y - p2s.y;
    const double X1_X3 = p1s.x - p2s.x, Y1_Y3 = p1s.y - p2s.y;
    const double u_a = (X4_X3 * Y1_Y3 - Y4_Y3 * X1_X3) / (Y4_Y3 * X2_X1 - X4_X3 * Y2_Y1);
    D2 r;
    r.x = p1s.x + X2_X1 * u_a;
    r.y = p1s.y + Y2_Y1 * u_a;
    return r;
}

/* getVanishingPoints (object_3d_util.cpp:602-607): vps[0..2] */
/* ycs: one entry of the yaw table = {yaw, cos(yaw), sin(yaw)}, the cosine and sine taken on the host with libm like the reference does */
__device__ __forceinline__ void g_vanishing_points(const double *KinvR, const double *ycs, D2 *vps)
{
    const double c = ycs[1], s = ycs[2];
    double h0, h1, h2;
    h0 = (KinvR[0] * c + KinvR[1] * s) + KinvR[2] * 0.0;
    h1 = (KinvR[3] * c + KinvR[4] * s) + KinvR[5] * 0.0;
    h2 = (KinvR[6] * c + KinvR[7] * s) + KinvR[8] * 0.0;
    vps[0].x = h0 / h2;
    vps[0].y = h1 / h2;
    h0 = (KinvR[0] * (-s) + KinvR[1] * c) + KinvR[2] * 0.0;
    h1 = (KinvR[3] * (-s) + KinvR[4] * c) + KinvR[5] * 0.0;
    h2 = (KinvR[6] * (-s) + KinvR[7] * c) + KinvR[8] * 0.0;
    vps[1].x = h0 / h2;
    vps[1].y = h1 / h2;
    h0 = (KinvR[0] * 0.0 + KinvR[1] * 0.0) + KinvR[2] * 1.0;
    h1 = (KinvR[3] * 0.0 + KinvR[4] * 0.0) + KinvR[5] * 1.0;
    h2 = (KinvR[6] * 0.0 + KinvR[7] * 0.0) + KinvR[8] * 1.0;
    vps[2].x = h0 / h2;
    vps[2].y = h1 / h2;
}

/* top-x sample i of a job: linespace<int>(left+5, right-5, step) (box_proposal_detail.cpp:144-146) */
__device__ __forceinline__ int g_top_x(const CsJob &jb, int i)
{
    if (jb.top_override) return jb.top_lo + (int)(((int64_t)(jb.top_hi - jb.top_lo) * i) / max(jb.n_top - 1, 1));
    return jb.top_lo + i * jb.top_step;
}

/* The corner chain of one candidate (box_proposal_detail.cpp:257-418).  Returns false at the first
 * failed test, exactly where the reference `continue`s.  c[0..7] = corners 1..8; vp1pos in {1,2}. */
__device__ __forceinline__ bool g_build_corners(const CsJob &jb, const D2 *vps, int top_x, int config_id, double shorted, D2 *c, int &vp1pos)
{
    const double left = jb.left, right = jb.right, top = jb.top, down = jb.down_y_expan;
    const D2 vp_1 = vps[0], vp_2 = vps[1], vp_3 = vps[2];
    D2 c1;
    c1.x = top_x;
    c1.y = top;
    vp1pos = 0;
    D2 c2 = g_hit_vertical(vp_1, c1, right, top, down);
    if (c2.x == -1) {
        c2 = g_hit_vertical(vp_1, c1, left, top, down);
        if (c2.x != -1) vp1pos = 2;
    } else
        vp1pos = 1;
    if (!(vp1pos > 0)) return false;
    if (g_norm(c1, c2) < shorted) return false;
    D2 c3, c4;
    if (config_id == 1) {
        c4 = (vp1pos == 1) ? g_hit_vertical(vp_2, c1, left, top, down) : g_hit_vertical(vp_2, c1, right, top, down);
        if (c4.y == -1) return false;
        if (g_norm(c1, c4) < shorted) return false;
        c3 = g_intersect(vp_2, c2, vp_1, c4);
        if (!g_inside(c3, left, top, right, down)) return false;
        if ((g_norm(c3, c4) < shorted) || (g_norm(c3, c2) < shorted)) return false;
    } else {
        c3 = (vp1pos == 1) ? g_hit_vertical(vp_2, c2, left, top, down) : g_hit_vertical(vp_2, c2, right, top, down);
        if (c3.y == -1) return false;
        if (g_norm(c2, c3) < shorted) return false;
        c4 = g_intersect(vp_1, c3, vp_2, c1);
        if (!g_inside(c4, left, (double)jb.roi_t, right, (double)jb.roi_b)) return false;
        if ((g_norm(c3, c4) < shorted) || (g_norm(c4, c1) < shorted)) return false;
    }
    const D2 c5 = g_hit_horizontal(vp_3, c3, left, right, down);
    if (c5.y == -1) return false;
    if (g_norm(c3, c5) < shorted) return false;
    const double el = jb.roi_l, et = jb.roi_t, er = jb.roi_r, eb = jb.roi_b;
    const D2 c6 = g_intersect(vp_2, c5, vp_3, c2);
    if (!g_inside(c6, el, et, er, eb)) return false;
    if ((g_norm(c6, c2) < shorted) || (g_norm(c6, c5) < shorted)) return false;
    const D2 c7 = g_intersect(vp_1, c6, vp_3, c1);
    if (!g_inside(c7, el, et, er, eb)) return false;
    if ((g_norm(c7, c1) < shorted) || (g_norm(c7, c6) < shorted)) return false;
    const D2 c8 = g_intersect(vp_1, c5, vp_2, c7);
    if (!g_inside(c8, el, et, er, eb)) return false;
    if ((g_norm(c8, c4) < shorted) || (g_norm(c8, c5) < shorted) || (g_norm(c8, c7) < shorted)) return false;
    c[0] = c1;
    c[1] = c2;
    c[2] = c3;
    c[3] = c4;
    c[4] = c5;
    c[5] = c6;
    c[6] = c7;
    c[7] = c8;
    return true;
}

/* visible edges and VP edge ids (box_proposal_detail.cpp:429-447), 0-based */
__constant__ int8_t c_vis1[9][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 0}, {1, 5}, {2, 4}, {3, 7}, {4, 7}, {4, 5}};
__constant__ int8_t c_vis2[7][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 0}, {1, 5}, {2, 4}, {4, 5}};
__constant__ int8_t c_vpe1[3][4] = {{0, 1, 7, 4}, {3, 0, 4, 5}, {3, 7, 1, 5}};
__constant__ int8_t c_vpe2[3][4] = {{0, 1, 2, 3}, {3, 0, 4, 5}, {2, 4, 1, 5}};

/* box_edge_sum_dists (object_3d_util.cpp:427-453): 11 samples per visible edge, float32 running sum in
 * the reference's order.  Out-of-range indices (the reference's latent inclusive-box UB) are clamped. */
__device__ __forceinline__ double g_edge_sum_dists(const float *__restrict__ dist, int pitch, int dw, int dh, const D2 *c, double off_x, double off_y,
                                                   int config_id, bool reweight)
{
    float sum_dist = 0;
    const int n_edges = (config_id == 1) ? 9 : 7;
    const bool rw = reweight && (config_id == 2);
    for (int e = 0; e < n_edges; e++) {
        const int i0 = (config_id == 1) ? c_vis1[e][0] : c_vis2[e][0];
        const int i1 = (config_id == 1) ? c_vis1[e][1] : c_vis2[e][1];
        const double ax = c[i0].x - off_x, ay = c[i0].y - off_y;
        const double bx = c[i1].x - off_x, by = c[i1].y - off_y;
#pragma unroll
        for (int s = 0; s < 11; s++) {
            const double t = (double)s / 10.0, u = 1 - (double)s / 10.0;
            const double px = t * ax + u * bx;
            const double py = t * ay + u * by;
            int ix = (int)px, iy = (int)py;
            ix = min(max(ix, 0), dw - 1);
            iy = min(max(iy, 0), dh - 1);
            float d1 = __ldg(dist + (size_t)iy * pitch + ix);
            if (rw) {
                if (4 <= e && e <= 5) d1 = (float)((double)d1 * 3.0 / 2.0);
                if (6 == e) d1 = (float)((double)d1 * 2.0);
            }
            sum_dist = sum_dist + d1;
        }
    }
    return (double)sum_dist;
}

/* box_edge_alignment_angle_error (object_3d_util.cpp:455-492); vp_angles 3x2, NaN = no support */
__device__ __forceinline__ double g_angle_error(const double *vp_angles, int config_id, const D2 *c)
{
    double total = 0;
    const double not_found_penalty = 30.0 / 180.0 * CS_PI * 2;
    for (int vp_id = 0; vp_id < 3; vp_id++) {
        double valid[2];
        int nv = 0;
        for (int i = 0; i < 2; i++) {
            const double a = vp_angles[vp_id * 2 + i];
            if (!isnan(a)) valid[nv++] = a;
        }
        if (nv > 0) {
            for (int ee = 0; ee < 2; ee++) {
                const int i0 = (config_id == 1) ? c_vpe1[vp_id][2 * ee] : c_vpe2[vp_id][2 * ee];
                const int i1 = (config_id == 1) ? c_vpe1[vp_id][2 * ee + 1] : c_vpe2[vp_id][2 * ee + 1];
                const double box_edge_angle = g_normalize_to_pi(cs_pm_atan2(c[i1].y - c[i0].y, c[i1].x - c[i0].x));
                double best = 100;
                for (int i = 0; i < nv; i++) {
                    double t = fabs(box_edge_angle - valid[i]);
                    t = g_min(t, CS_PI - t);
                    if (t < best) best = t;
                }
                total = total + best;
            }
        } else
            total = total + not_found_penalty;
    }
    return total;
}

/* plane_hits_3d for one pixel (object_3d_util.cpp:568-585) */
__device__ __forceinline__ void g_plane_hit(const double *T, const double *invK, const double *plane, D2 px, double *world)
{
    double ray[3];
    for (int i = 0; i < 3; i++) ray[i] = (invK[i * 3 + 0] * px.x + invK[i * 3 + 1] * px.y) + invK[i * 3 + 2] * 1.0;
    const double frac = -plane[3] / ((plane[0] * ray[0] + plane[1] * ray[1]) + plane[2] * ray[2]);
    double s[3];
    for (int i = 0; i < 3; i++) s[i] = frac * ray[i];
    double h[4];
    for (int i = 0; i < 4; i++) h[i] = ((T[i * 4 + 0] * s[0] + T[i * 4 + 1] * s[1]) + T[i * 4 + 2] * s[2]) + T[i * 4 + 3] * 1.0;
    for (int i = 0; i < 3; i++) world[i] = h[i] / h[3];
}

/* change_2d_corner_to_3d_object (object_3d_util.cpp:610-648) */
__device__ inline void g_lift_to_3d(const D2 *c, double config_id, double vp1pos, const double *ycs, const double *ground, const double *T,
                                    const double *invK, cs_cuboid_rec &o)
{
    const double yaw = ycs[0]; /* {yaw, cos(yaw), sin(yaw)} from the host-built yaw table */
    double g[4][3];
    for (int i = 0; i < 4; i++) g_plane_hit(T, invK, ground, c[4 + i], g[i]);
    double dx = g[0][0] - g[3][0], dy = g[0][1] - g[3][1], dz = g[0][2] - g[3][2];
    const double length_half = sqrt((dx * dx + dy * dy) + dz * dz) / 2;
    dx = g[0][0] - g[1][0];
    dy = g[0][1] - g[1][1];
    dz = g[0][2] - g[1][2];
    const double width_half = sqrt((dx * dx + dy * dy) + dz * dz) / 2;
    double nrm[3] = {dy * 1.0 - dz * 0.0, dz * 0.0 - dx * 1.0, dx * 0.0 - dy * 0.0};
    const double nn = sqrt((nrm[0] * nrm[0] + nrm[1] * nrm[1]) + nrm[2] * nrm[2]);
    for (int i = 0; i < 3; i++) nrm[i] /= nn;
    const double dist = -((nrm[0] * g[0][0] + nrm[1] * g[0][1]) + nrm[2] * g[0][2]);
    double pw[4] = {nrm[0], nrm[1], nrm[2], dist};
    if (dist < 0)
        for (int i = 0; i < 4; i++) pw[i] = -pw[i];
    double ps[4];
    for (int i = 0; i < 4; i++) ps[i] = ((T[0 * 4 + i] * pw[0] + T[1 * 4 + i] * pw[1]) + T[2 * 4 + i] * pw[2]) + T[3 * 4 + i] * pw[3];
    double top[3];
    g_plane_hit(T, invK, ps, c[1], top);
    const double height_half = top[2] / 2;
    const double mean_x = (((g[0][0] + g[1][0]) + g[2][0]) + g[3][0]) / 4;
    const double mean_y = (((g[0][1] + g[1][1]) + g[2][1]) + g[3][1]) / 4;
    o.pos[0] = mean_x;
    o.pos[1] = mean_y;
    o.pos[2] = height_half;
    o.rotY = yaw;
    o.scale[0] = length_half;
    o.scale[1] = width_half;
    o.scale[2] = height_half;
    o.box_config_type[0] = config_id;
    o.box_config_type[1] = vp1pos;
    const int perm1[8] = {6, 5, 8, 7, 2, 3, 4, 1};
    const int perm2[8] = {5, 6, 7, 8, 3, 2, 1, 4};
    for (int i = 0; i < 8; i++) {
        const int src = ((vp1pos == 1) ? perm1[i] : perm2[i]) - 1;
        o.box_corners_2d[0 * 8 + i] = (int)c[src].x;
        o.box_corners_2d[1 * 8 + i] = (int)c[src].y;
    }
    const double body[3][8] = {{1, 1, -1, -1, 1, 1, -1, -1}, {1, -1, -1, 1, 1, -1, -1, 1}, {-1, -1, -1, -1, 1, 1, 1, 1}};
    const double cy = ycs[1], sy = ycs[2];
    const double rot[9] = {cy, -sy, 0, sy, cy, 0, 0, 0, 1};
    double S[16];
    for (int i = 0; i < 16; i++) S[i] = 0;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double acc = 0;
            for (int k = 0; k < 3; k++) {
                const double term = rot[i * 3 + k] * (k == j ? o.scale[j] : 0.0);
                acc = (k == 0) ? term : acc + term;
            }
            S[i * 4 + j] = acc;
        }
    S[3] = o.pos[0];
    S[7] = o.pos[1];
    S[11] = o.pos[2];
    S[15] = 1;
    for (int k = 0; k < 8; k++) {
        double h[4];
        for (int i = 0; i < 4; i++) h[i] = ((S[i * 4 + 0] * body[0][k] + S[i * 4 + 1] * body[1][k]) + S[i * 4 + 2] * body[2][k]) + S[i * 4 + 3] * 1.0;
        for (int i = 0; i < 3; i++) o.box_corners_3d_world[i * 8 + k] = h[i] / h[3];
    }
}

#endif
