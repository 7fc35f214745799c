/* tests/host_core/carried_emu.cpp -- the product's pass structure for several boxes in a roll / pitch-sampled frame
 * (cube_slam_b200/csrc/cs_carried.h, what cs_detect_cuboids_batch runs with cs_set_profiling bit 10) with the CPU ORACLE standing in for
 * the device: a pass hands every frame's rank-r box to orc_detect_cuboid as a one-box call, telling it through orc_set_first_box_cam_yaw
 * which camera yaw the previous box left, and reads back the yaw this box leaves (orc_cam_yaw_left).  If the records that come out equal
 * the oracle's ordinary all-boxes-at-once call -- the reference's own loop -- then (a) the camera yaw is all a box inherits from its
 * predecessor and (b) the CSR slicing, scattering and hand-over of cs_carried_passes are right.  Test infrastructure, never shipped. */
#include <cmath>
#include <cstdint>

#include "../../cube_slam_b200/csrc/cs_carried.h"
#include "../../oracle/orc_api.h"

extern "C" void orc_set_first_box_cam_yaw(double yaw);
extern "C" double orc_cam_yaw_left(void);
static_assert(sizeof(orc_cuboid) == sizeof(cs_cuboid_rec), "the oracle's record and the ABI's record share one layout");

extern "C" int emu_carried(const uint8_t *imgs, int n_frames, int w, int h, int stride, int channels, const double *K, const double *T_wc, const double *boxes,
                           const int32_t *box_off, const double *lines, const int32_t *line_off, const orc_params *p, cs_cuboid_rec *out, int32_t *out_counts,
                           int32_t *n_passes)
{
    const int topk = p->max_cuboid_num;
    *n_passes = 0;
    auto run_pass = [&](const double *sub_boxes, const int32_t *sub_off, const std::vector<double> &cam_yaw, cs_cuboid_rec *recs, int32_t *counts,
                        std::vector<double> *yaw_left) -> int {
        ++*n_passes;
        int rc = 0;
        for (int f = 0; f < n_frames && rc == 0; f++) {
            if (sub_off[f + 1] == sub_off[f]) continue;
            if (sub_off[f + 1] - sub_off[f] != 1) return -100; /* a pass carries at most one box per frame */
            orc_set_first_box_cam_yaw(cam_yaw[f]);
            int cnt = 0;
            rc = orc_detect_cuboid(imgs + (size_t)f * h * stride, w, h, stride, channels, K, T_wc + (size_t)f * 16, sub_boxes + (size_t)sub_off[f] * 5, 1,
                                   lines + (size_t)line_off[f] * 4, line_off[f + 1] - line_off[f], p, topk, (orc_cuboid *)&recs[(size_t)sub_off[f] * topk], &cnt, nullptr,
                                   nullptr, nullptr);
            counts[sub_off[f]] = cnt;
            if (yaw_left) (*yaw_left)[f] = orc_cam_yaw_left();
        }
        orc_set_first_box_cam_yaw(std::nan(""));
        return rc;
    };
    return cs_carried_passes(n_frames, boxes, box_off, topk, out, out_counts, run_pass);
}
