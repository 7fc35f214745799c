/*
 * cs_host_pose.cpp -- host-side camera-pose tables and sample grids of detect_cuboid.
 *
 * These are a handful of FP64 values per frame (set_cam_pose, the roll/pitch pose hypotheses, the yaw
 * sample list).  They are evaluated on the host with libm -- like the reference does -- so that the
 * sample grids the kernels sweep are the reference's own numbers, and uploaded as small tables.
 *   set_calibration / set_cam_pose        detect_3d_cuboid/src/box_proposal_detail.cpp:36-54
 *   quat_to_euler_zyx, euler_zyx_to_rot   detect_3d_cuboid/src/matrix_utils.cpp:36-46,75-89
 *   linespace                             detect_3d_cuboid/src/matrix_utils.cpp:350-363
 * Compiled with -ffp-contract=off.
 */
#include "cs_host_pose.h"

#include <cmath>
#include <cstring>

namespace cshost {

static inline double cof(const double *m, int i, int j)
{
    const int a = (i + 1) % 3, b = (i + 2) % 3, c = (j + 1) % 3, d = (j + 2) % 3;
    return m[a * 3 + c] * m[b * 3 + d] - m[a * 3 + d] * m[b * 3 + c];
}

/* fixed-size 3x3 inverse, cofactor form (the evaluation order of Eigen's 3x3 inverse) */
void invert3(const double *m, double *out)
{
    const double k0 = cof(m, 0, 0), k1 = cof(m, 1, 0), k2 = cof(m, 2, 0);
    const double det = (k0 * m[0] + k1 * m[3]) + k2 * m[6];
    const double s = 1.0 / det;
    out[0] = k0 * s;
    out[1] = k1 * s;
    out[2] = k2 * s;
    for (int r = 1; r < 3; r++)
        for (int c = 0; c < 3; c++) out[r * 3 + c] = cof(m, c, r) * s;
}

static void matmul3(const double *a, const double *b, double *c)
{
    for (int r = 0; r < 3; r++)
        for (int k = 0; k < 3; k++) c[r * 3 + k] = (a[r * 3] * b[k] + a[r * 3 + 1] * b[3 + k]) + a[r * 3 + 2] * b[6 + k];
}

void euler_to_rot(double roll, double pitch, double yaw, double *R)
{
    const double cp = std::cos(pitch), sp = std::sin(pitch), sr = std::sin(roll), cr = std::cos(roll), sy = std::sin(yaw), cy = std::cos(yaw);
    R[0] = cp * cy;
    R[1] = (sr * sp * cy) - (cr * sy);
    R[2] = (cr * sp * cy) + (sr * sy);
    R[3] = cp * sy;
    R[4] = (sr * sp * sy) + (cr * cy);
    R[5] = (cr * sp * sy) - (sr * cy);
    R[6] = -sp;
    R[7] = sr * cp;
    R[8] = cr * cp;
}

/* Eigen::Quaterniond(R) followed by quat_to_euler_zyx */
static void euler_from_rot(const double *R, double *e)
{
    double x, y, z, w;
    const double tr = R[0] + R[4] + R[8];
    if (tr > 0.0) {
        double t = std::sqrt(tr + 1.0);
        w = 0.5 * t;
        t = 0.5 / t;
        x = (R[7] - R[5]) * t;
        y = (R[2] - R[6]) * t;
        z = (R[3] - R[1]) * t;
    } else {
        int i = 0;
        if (R[4] > R[0]) i = 1;
        if (R[8] > R[i * 4]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        double q[3];
        double t = std::sqrt(R[i * 4] - R[j * 4] - R[k * 4] + 1.0);
        q[i] = 0.5 * t;
        t = 0.5 / t;
        w = (R[k * 3 + j] - R[j * 3 + k]) * t;
        q[j] = (R[j * 3 + i] + R[i * 3 + j]) * t;
        q[k] = (R[k * 3 + i] + R[i * 3 + k]) * t;
        x = q[0];
        y = q[1];
        z = q[2];
    }
    e[0] = std::atan2(2 * (w * x + y * z), 1 - 2 * (x * x + y * y));
    e[1] = std::asin(2 * (w * y - z * x));
    e[2] = std::atan2(2 * (w * z + x * y), 1 - 2 * (y * y + z * z));
}

void make_pose(const double *K, const double *T, CsPose *pose, double *euler_out)
{
    double R[9], invR[9], e[3];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) R[r * 3 + c] = T[r * 4 + c];
    euler_from_rot(R, e);
    invert3(R, invR);
    matmul3(K, invR, pose->KinvR);
    std::memcpy(pose->T, T, sizeof(double) * 16);
    /* ground plane (0,0,1,0) seen from the sensor: T^T * plane */
    const double pw[4] = {0, 0, 1, 0};
    for (int c = 0; c < 4; c++) pose->ground[c] = ((T[c] * pw[0] + T[4 + c] * pw[1]) + T[8 + c] * pw[2]) + T[12 + c] * pw[3];
    pose->roll = e[0];
    pose->pitch = e[1];
    pose->camera_yaw = e[2];
    if (euler_out) std::memcpy(euler_out, e, sizeof(e));
}

void linespace_d(double start, double end, double step, std::vector<double> &out)
{
    while (start <= end) {
        out.push_back(start);
        start += step;
        if (out.size() > 1000) break;
    }
}

int linespace_count_i(int start, int end, int step)
{
    int n = 0;
    while (start <= end) {
        n++;
        start += step;
        if (n > 1000) break;
    }
    return n;
}


/* ---- cuboid -> g2o measurement (object_slam/src/main_obj.cpp:455-473,505; g2o_Object.h:36-41,127-133; Thirdparty/g2o/g2o/types/se3quat.h)
 * Quaternions are (x, y, z, w) like Eigen's coeffs().  Each step is the g2o / Eigen operation it names. */
namespace {
struct Q {
    double x, y, z, w;
};
/* SE3Quat::normalizeRotation: w >= 0, unit norm */
void normalize_rotation(Q &q)
{
    if (q.w < 0) {
        q.x = -q.x;
        q.y = -q.y;
        q.z = -q.z;
        q.w = -q.w;
    }
    const double n = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    q.x /= n;
    q.y /= n;
    q.z /= n;
    q.w /= n;
}
/* Eigen quaternion product a * b */
Q qmul(const Q &a, const Q &b)
{
    Q r;
    r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
    r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
    return r;
}
/* Eigen QuaternionBase::_transformVector: v + w * (2 q x v) + q x (2 q x v) */
void qrot(const Q &q, const double *v, double *out)
{
    double uv[3] = {q.y * v[2] - q.z * v[1], q.z * v[0] - q.x * v[2], q.x * v[1] - q.y * v[0]};
    uv[0] += uv[0];
    uv[1] += uv[1];
    uv[2] += uv[2];
    out[0] = v[0] + q.w * uv[0] + (q.y * uv[2] - q.z * uv[1]);
    out[1] = v[1] + q.w * uv[1] + (q.z * uv[0] - q.x * uv[2]);
    out[2] = v[2] + q.w * uv[2] + (q.x * uv[1] - q.y * uv[0]);
}
/* Eigen::Quaterniond(Matrix3d): Shepperd's branches, as in euler_from_rot above */
Q quat_from_rot(const double *R)
{
    Q q;
    const double tr = R[0] + R[4] + R[8];
    if (tr > 0.0) {
        double t = std::sqrt(tr + 1.0);
        q.w = 0.5 * t;
        t = 0.5 / t;
        q.x = (R[7] - R[5]) * t;
        q.y = (R[2] - R[6]) * t;
        q.z = (R[3] - R[1]) * t;
    } else {
        int i = 0;
        if (R[4] > R[0]) i = 1;
        if (R[8] > R[i * 4]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        double v[3];
        double t = std::sqrt(R[i * 4] - R[j * 4] - R[k * 4] + 1.0);
        v[i] = 0.5 * t;
        t = 0.5 / t;
        q.w = (R[k * 3 + j] - R[j * 3 + k]) * t;
        v[j] = (R[j * 3 + i] + R[i * 3 + j]) * t;
        v[k] = (R[k * 3 + i] + R[i * 3 + k]) * t;
        q.x = v[0];
        q.y = v[1];
        q.z = v[2];
    }
    return q;
}
}  // namespace

void quat_of_rotation(const double *R, double *q_xyzw)
{
    const Q q = quat_from_rot(R);
    q_xyzw[0] = q.x;
    q_xyzw[1] = q.y;
    q_xyzw[2] = q.z;
    q_xyzw[3] = q.w;
}

void cuboid_measurement(const double *pos, double rotY, const double *cam_t, const double *cam_q_xyzw, const double *cam_euler_new,
                        double *meas_t, double *meas_q_xyzw)
{
    /* cube_ground_value.fromMinimalVector: pose = SE3Quat(zyx_euler_to_quat(0, 0, rotY), pos) */
    Q qo;
    {
        const double sy = std::sin(rotY * 0.5), cy = std::cos(rotY * 0.5);
        const double sp = std::sin(0.0), cp = std::cos(0.0), sr = std::sin(0.0), cr = std::cos(0.0);
        qo.w = cr * cp * cy + sr * sp * sy;
        qo.x = sr * cp * cy - cr * sp * sy;
        qo.y = cr * sp * cy + sr * cp * sy;
        qo.z = cr * cp * sy - sr * sp * cy;
        normalize_rotation(qo);
    }
    /* the camera pose: SE3Quat(Vector7d) normalises; with sampled roll / pitch SE3Quat(euler_zyx_to_rot(new eulers), t) */
    Q qc;
    if (cam_euler_new) {
        double R[9];
        euler_to_rot(cam_euler_new[0], cam_euler_new[1], cam_euler_new[2], R);
        qc = quat_from_rot(R);
    } else {
        qc.x = cam_q_xyzw[0];
        qc.y = cam_q_xyzw[1];
        qc.z = cam_q_xyzw[2];
        qc.w = cam_q_xyzw[3];
    }
    normalize_rotation(qc);
    /* transform_to: Twc.inverse() * pose.  inverse(): r = conj, t = r * (-t); operator*: t += r * t2, r *= r2, normalise */
    Q qi = {-qc.x, -qc.y, -qc.z, qc.w};
    const double nt[3] = {cam_t[0] * -1., cam_t[1] * -1., cam_t[2] * -1.};
    double ti[3], rp[3];
    qrot(qi, nt, ti);
    qrot(qi, pos, rp);
    meas_t[0] = ti[0] + rp[0];
    meas_t[1] = ti[1] + rp[1];
    meas_t[2] = ti[2] + rp[2];
    Q qm = qmul(qi, qo);
    normalize_rotation(qm);
    meas_q_xyzw[0] = qm.x;
    meas_q_xyzw[1] = qm.y;
    meas_q_xyzw[2] = qm.z;
    meas_q_xyzw[3] = qm.w;
}

}  // namespace cshost


/* get_cuboid_draw_edge_markers(final_universal_object = true) + plot_image_with_cuboid_edges' marker table
 * (detect_3d_cuboid/src/object_3d_util.cpp:56-69,80-92,109-112): the 12 edges of a detected cuboid as the reference draws them. */
extern "C" int cs_cuboid_draw_edges(const cs_cuboid_rec *rec, int32_t edges[12][8])
{
    if (!rec || !edges) return CS_ERR_INVALID_ARG;
    /* rows: start corner, end corner (1-based, as the reference writes them), marker type (1-based) */
    static const int cfg1_left[12][3] = {{3, 4, 4}, {4, 1, 2}, {4, 8, 6}, {1, 2, 3}, {2, 3, 1}, {2, 6, 5}, {1, 5, 5}, {3, 7, 5}, {5, 6, 3}, {6, 7, 1}, {7, 8, 3}, {8, 5, 1}};
    static const int cfg1_right[12][3] = {{2, 3, 2}, {3, 4, 4}, {3, 7, 6}, {1, 2, 3}, {1, 4, 1}, {2, 6, 5}, {1, 5, 5}, {4, 8, 5}, {5, 6, 3}, {6, 7, 1}, {7, 8, 3}, {8, 5, 1}};
    static const int cfg2[12][3] = {{2, 3, 2}, {3, 4, 4}, {4, 1, 2}, {3, 7, 6}, {4, 8, 6}, {1, 2, 3}, {2, 6, 5}, {1, 5, 5}, {5, 6, 3}, {6, 7, 1}, {7, 8, 3}, {8, 5, 1}};
    /* line_markers: B, G, R, thickness */
    static const int markers[6][4] = {{0, 0, 255, 2}, {0, 0, 255, 1}, {0, 255, 0, 2}, {0, 255, 0, 1}, {255, 0, 0, 2}, {255, 0, 0, 1}};
    const int(*tab)[3] = (rec->box_config_type[0] == 1) ? ((rec->box_config_type[1] == 1) ? cfg1_left : cfg1_right) : cfg2;
    for (int e = 0; e < 12; e++) {
        const int a = tab[e][0] - 1, b = tab[e][1] - 1, m = tab[e][2] - 1;
        edges[e][0] = rec->box_corners_2d[0 * 8 + a];
        edges[e][1] = rec->box_corners_2d[1 * 8 + a];
        edges[e][2] = rec->box_corners_2d[0 * 8 + b];
        edges[e][3] = rec->box_corners_2d[1 * 8 + b];
        for (int k = 0; k < 4; k++) edges[e][4 + k] = markers[m][k];
    }
    return CS_OK;
}
