/*
 * oracle/orc_api.h -- C interface of the CPU ORACLE (test infrastructure, NOT product code).
 *
 * The oracle is a dependency-free restatement (no Eigen / OpenCV / ROS) of the reference's
 * per-frame cuboid proposal path.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product library
 * (cube_slam_b200/lib/libcubeslam_b200.so) never links or calls anything in oracle/.
 *
 * PARITY STATUS, per stage:
 *   line detection (stage i, both flavours): PINNED -- the reference's own lsd.cpp and binary_descriptor.cpp compile from
 *     /root/reference against a small OpenCV stand-in (oracle/ref/, oracle/_ref/) and the restatements are byte-identical to them
 *     (tests/test_oracle_ref_lsd.py, tests/test_oracle_ref_edlines.py, checksums in tests/golden/); the shipped LSD segment file is
 *     reproduced whole (tests/test_oracle_lines.py).
 *   cuboid proposals (stage ii): PINNED -- detect_3d_cuboid's three sources compile from /root/reference against small stand-ins for
 *     Eigen and OpenCV (oracle/ref/minieigen.hpp, minicv.hpp; oracle/_ref/libcuboid_ref.so) and every field of every cuboid they return
 *     equals this oracle's record, == on doubles, in every mode tried (tests/test_oracle_ref_detect_cuboid.py).  The loops, thresholds,
 *     indices and branches are the reference's code; the linear algebra under them is the stand-in's, written as Eigen documents it.
 *     The OpenCV stages inside (cvtColor, Canny, distanceTransform) are pinned bit for bit against the in-container cv2 4.13
 *     (tests/test_oracle_cv_parity.py); a soft cross-check against the MATLAB cuboids the authors ship for the object_slam sequence
 *     remains (tests/test_oracle_matlab_crosscheck.py).
 */
#ifndef ORC_API_H
#define ORC_API_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* mirrors detect_3d_cuboid's public mode members (detect_3d_cuboid.h:65-79) and the
 * hard-coded locals of detect_cuboid (box_proposal_detail.cpp:79-87,177-179,197,126-128,144). */
typedef struct orc_params {
    int consider_config_1;              /* detect_3d_cuboid.h:72 */
    int consider_config_2;              /* :73 */
    int whether_sample_cam_roll_pitch;  /* :74 */
    int whether_sample_bbox_height;     /* :75 */
    int max_cuboid_num;                 /* :77 */
    double nominal_skew_ratio;          /* :78 */
    double max_cut_skew;                /* :79 */
    double vp12_edge_angle_thre;        /* box_proposal_detail.cpp:79  (15) */
    double vp3_edge_angle_thre;         /* :80 (10) */
    double shorted_edge_thre;           /* :81 (20) */
    int reweight_edge_distance;         /* :82 (true) */
    int whether_normalize_two_errors;   /* :85 (true) */
    double weight_vp_angle;             /* :86 (0.8) */
    double weight_skew_error;           /* :87 (1.5) */
    double pre_merge_dist_thre;         /* :177 (20) */
    double pre_merge_angle_thre;        /* :178 (5 deg) */
    double edge_length_threshold;       /* :179 (30) */
    double canny_low;                   /* :197 (80) */
    double canny_high;                  /* :197 (200) */
    double yaw_half_range_deg;          /* :128 (45) */
    double yaw_step_deg;                /* :128 (6) */
    int top_sample_count_override;      /* 0 = reference formula (:144); >0 = BASELINE C5 dense sweep */
} orc_params;

/* POD mirror of class cuboid (detect_3d_cuboid.h:15-36) + bookkeeping */
typedef struct orc_cuboid {
    double pos[3];
    double scale[3];
    double rotY;
    double box_config_type[2];
    int32_t box_corners_2d[16];          /* row-major 2x8 */
    double box_corners_3d_world[24];     /* row-major 3x8 */
    double rect_detect_2d[4];
    double edge_distance_error;
    double edge_angle_error;
    double normalized_error;
    double skew_ratio;
    double down_expand_height;
    double camera_roll_delta;
    double camera_pitch_delta;
    double combined_score;               /* box_proposal_detail.cpp:526 */
    int32_t proposal_index;              /* row in the reference's valid-proposal list of its height sample */
    int32_t height_sample_id;
    int32_t valid;
    int32_t pad_;
} orc_cuboid;

/* optional trace of one (object, height-sample) ROI job; every pointer may be NULL */
typedef struct orc_trace {
    int want_object;            /* which bbox to trace */
    int want_height_sample;     /* which height sample */
    /* outputs */
    int roi[4];                 /* left, top, width, height of the dist-map ROI */
    int n_lines_roi;            /* lines inside ROI before merge */
    int n_lines_merged;         /* after merge + length filter */
    double *merged_lines;       /* cap_lines x 4 */
    int cap_lines;
    uint8_t *canny;             /* cap_px */
    float *dist;                /* cap_px */
    int cap_px;
    int n_candidates;           /* enumerated (roll,pitch,yaw,top,config) tuples */
    int n_valid;
    double *rows;               /* cap_valid x 9  (all_configs_error_one_objH) */
    double *corners;            /* cap_valid x 16 (row-major 2x8) */
    int32_t *cand_index;        /* cap_valid: enumeration index of each valid proposal */
    int cap_valid;
    int n_kept;
    int32_t *kept_ids;          /* cap_valid */
    double *kept_scores;        /* cap_valid */
} orc_trace;

void orc_default_params(orc_params *p);

/* cv::cvtColor(BGR2GRAY) restated; fixed15 != 0 -> OpenCV 4.x 15-bit coefficients, else 2.4/3.x 14-bit */
void orc_bgr2gray(const uint8_t *bgr, int w, int h, int stride, uint8_t *gray, int gstride, int fixed15);
/* cv::Canny(src, dst, low, high) aperture 3, L1 gradient, isolated ROI (BORDER_REPLICATE) */
void orc_canny(const uint8_t *src, int w, int h, int stride, double low, double high, uint8_t *dst);
/* cv::distanceTransform(255 - edges, DIST_L2, 3): 3x3 chamfer, 16.16 fixed point (non-IPP C path) */
void orc_chamfer_dt(const uint8_t *edges, int w, int h, float *dist);

/* merge_break_lines (object_3d_util.cpp:300-376); lines n x 4 in, out n_out x 4 (cap n); returns n_out */
int orc_merge_break_lines(const double *lines, int n, double dist_thre, double angle_thre_deg,
                          double len_thre, double *out);

/* which atan2 the angle-error chain uses: 1 (default) the arithmetic definition shared with the CUDA path (pmath.h), 0 libm's */
void orc_set_portable_atan2(int on);
double orc_atan2_portable(double y, double x);

/* knife-edge bookkeeping of fuse_normalize_scores_v2's angle-cut comparison (see cuboid_oracle.cpp) */
double orc_last_cut_margin(int box); /* smallest relative gap of that comparison for box `box` (< 64) of this thread's last orc_detect_cuboid */
void orc_set_cut_flip(int box);      /* take the other branch for that box where the gap is < 1e-13; -1 = off */

/* detect_3d_cuboid::detect_cuboid (box_proposal_detail.cpp:56-557) for one frame.
 * boxes N x 5 [x y w h prob] 0-based; lines M x 4; out N x topk_cap; out_counts N. */
int orc_detect_cuboid(const uint8_t *img, int w, int h, int stride, int channels,
                      const double *K, const double *T_wc, const double *boxes, int N,
                      const double *lines, int M, const orc_params *p,
                      int topk_cap, orc_cuboid *out, int *out_counts,
                      int64_t *n_candidates_total, int64_t *n_valid_total, orc_trace *trace);

/* set_cam_pose (box_proposal_detail.cpp:42-54): out[0..2]=euler, out[3..11]=KinvR, out[12]=camera_yaw */
void orc_cam_pose(const double *K, const double *T_wc, double *out13);

#ifdef __cplusplus
}
#endif
#endif
