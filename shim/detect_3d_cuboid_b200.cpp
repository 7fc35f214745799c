/*
 * shim/detect_3d_cuboid_b200.cpp -- class detect_3d_cuboid (detect_3d_cuboid/include/detect_3d_cuboid/detect_3d_cuboid.h:53-80) implemented
 * on libcubeslam_b200.so.  Compile this translation unit INSTEAD OF detect_3d_cuboid/src/box_proposal_detail.cpp inside the reference's
 * detect_3d_cuboid package (its header, object_3d_util.cpp and matrix_utils.cpp stay as they are; object_slam/src/main_obj.cpp:354-366,449
 * and orb_object_slam/src/Tracking.cc:242-244,1625 call it unchanged), add include/ to the include path and link -lcubeslam_b200.
 *
 * The image this repository is built in has neither Eigen nor OpenCV's C++ headers, so the body is guarded: without them the file is an
 * empty translation unit (tests/test_shim_compiles.py compiles it either way).
 */
#if defined(__has_include)
#if __has_include(<Eigen/Core>) && __has_include(<opencv2/core/core.hpp>) && __has_include("detect_3d_cuboid/detect_3d_cuboid.h")
#define CS_SHIM_ENABLED 1
#endif
#endif

#ifdef CS_SHIM_ENABLED
#include <cstdio>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <unordered_map>
#include <vector>

#include "cube_slam_b200.h"
#include "detect_3d_cuboid/detect_3d_cuboid.h"
#include "detect_3d_cuboid/object_3d_util.h" /* plot_image_with_cuboid: the drawing stays reference code (object_3d_util.cpp:54-131) */

#include <opencv2/highgui/highgui.hpp>

namespace {

/* One context per detector object (the reference object is stateful and used from one thread): the class layout
 * cannot grow a member, so the contexts live in a side table keyed by the object's address.  Contexts are created on first use and
 * live until the process ends (the reference never destroys its detectors either). */
struct CtxTable {
    std::mutex mu;
    std::unordered_map<const detect_3d_cuboid *, cs_ctx *> map;
    cs_ctx *get(const detect_3d_cuboid *self)
    {
        std::lock_guard<std::mutex> g(mu);
        auto it = map.find(self);
        if (it != map.end()) return it->second;
        cs_ctx *c = cs_create(/*device*/ 0, /*max_width*/ 2048, /*max_height*/ 2048, /*max_frames*/ 1, /*max_boxes*/ 64, /*max_lines*/ 8192);
        if (!c) throw std::runtime_error("cube_slam_b200: cs_create failed (no CUDA device?)");
        /* a drop-in has the reference's semantics to the letter: with whether_sample_cam_roll_pitch, later boxes of a frame start from the
         * cam_pose the earlier ones left (box_proposal_detail.cpp:126-128 after :237,485) -- bit 10, one pass per box rank; a frame with a
         * single box (all object_slam ever passes) takes the ordinary one-pass path either way */
        cs_set_profiling(c, 1024);
        map.emplace(self, c);
        return c;
    }
};
CtxTable &table()
{
    static CtxTable t;
    return t;
}

/* the ABI takes row-major arrays; Eigen's default storage is column-major */
template <typename M>
std::vector<double> row_major(const M &m, int cols)
{
    std::vector<double> v((size_t)m.rows() * cols);
    for (int i = 0; i < (int)m.rows(); i++)
        for (int j = 0; j < cols; j++) v[(size_t)i * cols + j] = m(i, j);
    return v;
}

}  // namespace

/* box_proposal_detail.cpp:36-40 */
void detect_3d_cuboid::set_calibration(const Eigen::Matrix3d &Kalib)
{
    cam_pose.Kalib = Kalib;
    cam_pose.invK = Kalib.inverse();
    cs_set_calibration(table().get(this), row_major(Kalib, 3).data());
}

/* box_proposal_detail.cpp:42-54: callers read cam_pose_raw.euler_angle (main_obj.cpp:465) */
void detect_3d_cuboid::set_cam_pose(const Eigen::Matrix4d &transToWolrd)
{
    double e[3], kr[9];
    cs_cam_pose(row_major(cam_pose.Kalib, 3).data(), row_major(transToWolrd, 4).data(), e, kr);
    cam_pose.transToWolrd = transToWolrd;
    cam_pose.rotationToWorld = transToWolrd.topLeftCorner<3, 3>();
    cam_pose.euler_angle = Eigen::Vector3d(e[0], e[1], e[2]);
    cam_pose.invR = cam_pose.rotationToWorld.inverse();
    cam_pose.projectionMatrix = cam_pose.Kalib * transToWolrd.inverse().topRows<3>();
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) cam_pose.KinvR(i, j) = kr[i * 3 + j];
    cam_pose.camera_yaw = e[2];
}

/* box_proposal_detail.cpp:56-557 */
void detect_3d_cuboid::detect_cuboid(const cv::Mat &rgb_img, const Eigen::Matrix4d &transToWolrd, const Eigen::MatrixXd &obj_bbox_coors,
                                     Eigen::MatrixXd all_lines_raw, std::vector<ObjectSet> &all_object_cuboids)
{
    cs_ctx *ctx = table().get(this);
    set_cam_pose(transToWolrd);
    cam_pose_raw = cam_pose;
    const int N = (int)obj_bbox_coors.rows();
    all_object_cuboids.resize(N); /* :71-72 */
    if (N == 0) return;
    cs_cuboid_params p;
    cs_default_cuboid_params(&p);
    p.consider_config_1 = consider_config_1;
    p.consider_config_2 = consider_config_2;
    p.whether_sample_cam_roll_pitch = whether_sample_cam_roll_pitch;
    p.whether_sample_bbox_height = whether_sample_bbox_height;
    p.max_cuboid_num = max_cuboid_num;
    p.nominal_skew_ratio = nominal_skew_ratio;
    p.max_cut_skew = max_cut_skew;
    const std::vector<double> boxes = row_major(obj_bbox_coors, 5), lines = row_major(all_lines_raw, 4), T = row_major(transToWolrd, 4);
    const cv::Mat img = rgb_img.isContinuous() ? rgb_img : rgb_img.clone();
    const int topk = max_cuboid_num > 0 ? max_cuboid_num : 1;
    std::vector<cs_cuboid_rec> recs((size_t)N * topk);
    std::vector<int32_t> counts(N);
    const int rc = cs_detect_cuboids(ctx, img.data, img.cols, img.rows, (int)img.step, img.channels(), T.data(), boxes.data(), N, lines.data(),
                                     (int)all_lines_raw.rows(), &p, recs.data(), counts.data());
    if (rc != CS_OK) { /* the reference prints and carries on; it never throws from detect_cuboid */
        std::printf("detect_cuboid: %s\n", cs_last_error(ctx));
        return;
    }
    for (int i = 0; i < N; i++)
        for (int k = 0; k < counts[i]; k++) {
            /* heap cuboids the caller keeps for good (box_proposal_detail.cpp:489,535; callers never delete them) */
            const cs_cuboid_rec &r = recs[(size_t)i * topk + k];
            cuboid *c = new cuboid();
            c->pos = Eigen::Vector3d(r.pos[0], r.pos[1], r.pos[2]);
            c->scale = Eigen::Vector3d(r.scale[0], r.scale[1], r.scale[2]);
            c->rotY = r.rotY;
            c->box_config_type = Eigen::Vector2d(r.box_config_type[0], r.box_config_type[1]);
            c->box_corners_2d.resize(2, 8);
            c->box_corners_3d_world.resize(3, 8);
            for (int j = 0; j < 8; j++) {
                for (int a = 0; a < 2; a++) c->box_corners_2d(a, j) = r.box_corners_2d[a * 8 + j];
                for (int a = 0; a < 3; a++) c->box_corners_3d_world(a, j) = r.box_corners_3d_world[a * 8 + j];
            }
            c->rect_detect_2d = Eigen::Vector4d(r.rect_detect_2d[0], r.rect_detect_2d[1], r.rect_detect_2d[2], r.rect_detect_2d[3]);
            c->edge_distance_error = r.edge_distance_error;
            c->edge_angle_error = r.edge_angle_error;
            c->normalized_error = r.normalized_error;
            c->skew_ratio = r.skew_ratio;
            c->down_expand_height = r.down_expand_height;
            c->camera_roll_delta = r.camera_roll_delta;
            c->camera_pitch_delta = r.camera_pitch_delta;
            all_object_cuboids[i].push_back(c);
        }
    if (whether_plot_final_images || whether_save_final_images) { /* :541-556 */
        cv::Mat frame_all_cubes_img = rgb_img.clone();
        for (size_t i = 0; i < all_object_cuboids.size(); i++)
            if (!all_object_cuboids[i].empty()) plot_image_with_cuboid(frame_all_cubes_img, all_object_cuboids[i][0]);
        if (whether_save_final_images) cuboids_2d_img = frame_all_cubes_img;
        if (whether_plot_final_images) {
            cv::imshow("frame_all_cubes_img", frame_all_cubes_img);
            cv::waitKey(0);
        }
    }
}
#endif /* CS_SHIM_ENABLED */
