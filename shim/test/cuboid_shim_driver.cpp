/*
 * shim/test/cuboid_shim_driver.cpp -- TEST HARNESS for shim/detect_3d_cuboid_b200.cpp (not product code).
 *
 * Builds the shim the way a maintainer would -- INSTEAD OF detect_3d_cuboid/src/box_proposal_detail.cpp, next to the reference's own
 * object_3d_util.cpp and matrix_utils.cpp (included here from /root/reference) and against the reference's class header
 * (detect_3d_cuboid/include/detect_3d_cuboid/detect_3d_cuboid.h:53-80) -- and calls it the way object_slam/src/main_obj.cpp:354-361,449
 * does: construct detect_3d_cuboid, set_calibration, the mode flags, detect_cuboid(image, pose, boxes, lines, out).  This image has
 * neither Eigen nor OpenCV C++ headers: Eigen::Matrix is oracle/ref/minieigen.hpp and cv::Mat oracle/ref/minicv.hpp (stand-ins, see those
 * files); the class, the shim and libcubeslam_b200.so are the real things.
 *
 * Built by oracle/Makefile (only where the reference checkout exists) into oracle/_ref/libshim_cuboid.so; tests/test_gpu_shim_runs.py
 * loads it on the GPU box.
 */
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <exception>
#include <vector>

#include "/root/reference/detect_3d_cuboid/src/matrix_utils.cpp"
#include "/root/reference/detect_3d_cuboid/src/object_3d_util.cpp"

/* same flat record as oracle/ref/cuboid_ref.cpp: 60 doubles per cuboid */
extern "C" int shim_detect_cuboid(const uint8_t *img, int w, int h, int channels, const double *K9, const double *T16, const double *boxes, int n_boxes,
                                  const double *lines, int n_lines, const int *flags, double nominal_skew_ratio, int max_cuboid_num, double *out,
                                  int cap_per_box, int *counts, double *euler_out)
{
    try {
        detect_3d_cuboid det;
        det.whether_plot_detail_images = false;
        det.whether_plot_final_images = false;
        det.whether_save_final_images = false;
        det.print_details = false;
        Eigen::Matrix3d K;
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) K(i, j) = K9[i * 3 + j];
        det.set_calibration(K);
        det.consider_config_1 = flags[0] != 0;
        det.consider_config_2 = flags[1] != 0;
        det.whether_sample_cam_roll_pitch = flags[2] != 0;
        det.whether_sample_bbox_height = flags[3] != 0;
        det.nominal_skew_ratio = nominal_skew_ratio;
        det.max_cuboid_num = max_cuboid_num;
        Eigen::Matrix4d T;
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++) T(i, j) = T16[i * 4 + j];
        Eigen::MatrixXd bb(n_boxes, 5), ed(n_lines, 4);
        for (int i = 0; i < n_boxes; i++)
            for (int j = 0; j < 5; j++) bb(i, j) = boxes[i * 5 + j];
        for (int i = 0; i < n_lines; i++)
            for (int j = 0; j < 4; j++) ed(i, j) = lines[i * 4 + j];
        cv::Mat image(h, w, channels == 3 ? CV_8UC3 : CV_8UC1);
        std::memcpy(image.data, img, (size_t)w * h * channels);
        std::vector<ObjectSet> all;
        det.detect_cuboid(image, T, bb, ed, all);
        for (int i = 0; i < 3; i++) euler_out[i] = det.cam_pose_raw.euler_angle(i); /* what main_obj.cpp:465 reads back */
        for (int b = 0; b < n_boxes; b++) {
            const int n = b < (int)all.size() ? (int)all[b].size() : 0;
            counts[b] = n;
            for (int k = 0; k < n && k < cap_per_box; k++) {
                const cuboid *c = all[b][k];
                double *o = out + ((size_t)b * cap_per_box + k) * 60;
                int q = 0;
                for (int i = 0; i < 3; i++) o[q++] = c->pos(i);
                o[q++] = c->rotY;
                for (int i = 0; i < 3; i++) o[q++] = c->scale(i);
                for (int i = 0; i < 2; i++) o[q++] = c->box_config_type(i);
                for (int i = 0; i < 2; i++)
                    for (int j = 0; j < 8; j++) o[q++] = c->box_corners_2d(i, j);
                for (int i = 0; i < 3; i++)
                    for (int j = 0; j < 8; j++) o[q++] = c->box_corners_3d_world(i, j);
                for (int i = 0; i < 4; i++) o[q++] = c->rect_detect_2d(i);
                o[q++] = c->edge_distance_error;
                o[q++] = c->edge_angle_error;
                o[q++] = c->normalized_error;
                o[q++] = c->skew_ratio;
                o[q++] = c->down_expand_height;
                o[q++] = c->camera_roll_delta;
                o[q++] = c->camera_pitch_delta;
            }
        }
        for (auto &set : all)
            for (cuboid *c : set) delete c;
        return 0;
    } catch (const std::exception &e) {
        fprintf(stderr, "shim_detect_cuboid: %s\n", e.what());
        return -1;
    }
}
