#ifndef CS_HOST_POSE_H
#define CS_HOST_POSE_H
#include <vector>

#include "cs_internal.h"

namespace cshost {
void invert3(const double *m, double *out);
void euler_to_rot(double roll, double pitch, double yaw, double *R);
/* set_cam_pose for one transform: fills KinvR, T, ground, roll/pitch/camera_yaw */
void make_pose(const double *K, const double *T, CsPose *pose, double *euler_out);
void linespace_d(double start, double end, double step, std::vector<double> &out);
int linespace_count_i(int start, int end, int step);
/* object -> camera-frame measurement of object_slam (main_obj.cpp:455-473); cam_euler_new NULL unless roll / pitch were sampled */
/* Eigen::Quaterniond(Matrix3d) of a row-major 3x3 rotation, (x, y, z, w) */
void quat_of_rotation(const double *R, double *q_xyzw);
void cuboid_measurement(const double *pos, double rotY, const double *cam_t, const double *cam_q_xyzw, const double *cam_euler_new, double *meas_t,
                        double *meas_q_xyzw);
}  // namespace cshost
#endif
