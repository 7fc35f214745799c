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
}  // namespace cshost
#endif
