/*
 * oracle/ref/cuboid_ref.cpp -- CPU ORACLE, TEST INFRASTRUCTURE ONLY: the reference's own detect_3d_cuboid::detect_cuboid, compiled from
 * /root/reference.
 *
 * One translation unit made of the reference's three sources, included from where they lie:
 *   detect_3d_cuboid/src/matrix_utils.cpp, object_3d_util.cpp, box_proposal_detail.cpp
 * with the reference's own headers, against stand-ins for what this image lacks: oracle/ref/minieigen.hpp for Eigen, oracle/ref/minicv.hpp
 * for OpenCV (Canny / distanceTransform / cvtColor forwarded to the oracle's cv2-pinned restatements), empty ROS / profiler headers
 * (oracle/ref/fakeros).  No reference source is copied.  The sampling loops, the configuration logic, every threshold and index, the
 * scoring and ranking are the reference's code; the linear algebra underneath is the stand-in's (see minieigen.hpp for what that means for
 * the last bits of continuous outputs).
 *
 * Entry point: what object_slam/src/main_obj.cpp:354-361,449 does -- construct detect_3d_cuboid, set_calibration, the mode flags,
 * detect_cuboid(image, pose, boxes, lines, out) -- flattened to plain arrays.
 */
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <exception>
#include <vector>

#include "/root/reference/detect_3d_cuboid/src/matrix_utils.cpp"
#include "/root/reference/detect_3d_cuboid/src/object_3d_util.cpp"
#include "/root/reference/detect_3d_cuboid/src/box_proposal_detail.cpp"

/* flat record per returned cuboid (doubles): pos[3], rotY, scale[3], box_config_type[2], box_corners_2d[16] (2 x 8, row-major),
 * box_corners_3d_world[24] (3 x 8, row-major), rect_detect_2d[4], edge_distance_error, edge_angle_error, normalized_error, skew_ratio,
 * down_expand_height, camera_roll_delta, camera_pitch_delta  = 60 doubles */
enum { REF_CUBOID_DOUBLES = 60 };

extern "C" int ref_detect_cuboid(const uint8_t *img, int w, int h, int channels, const double *K9, const double *T16, const double *boxes, int n_boxes,
                                 const double *lines, int n_lines, const int *flags /* consider_config_1, _2, sample_roll_pitch, sample_bbox_height */,
                                 double nominal_skew_ratio, int max_cuboid_num, double *out, int cap_per_box, int *counts)
{
    try {
        detect_3d_cuboid det;
        det.whether_plot_detail_images = false;
        det.whether_plot_final_images = false;
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
        for (int b = 0; b < n_boxes; b++) {
            const int n = b < (int)all.size() ? (int)all[b].size() : 0;
            counts[b] = n;
            for (int k = 0; k < n && k < cap_per_box; k++) {
                const cuboid *c = all[b][k];
                double *o = out + ((size_t)b * cap_per_box + k) * REF_CUBOID_DOUBLES;
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
        fprintf(stderr, "ref_detect_cuboid: %s\n", e.what());
        return -1;
    }
}
