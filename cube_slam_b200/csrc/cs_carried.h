/* cs_carried.h -- the pass structure of cs_detect_cuboids_batch in carried-pose mode (cs_set_profiling bit 10), host-only and free of
 * CUDA so that the CPU test suite can run it with the oracle standing in for the device (tests/host_core/carried_emu.cpp).
 *
 * With whether_sample_cam_roll_pitch the reference works through the boxes of a frame in order and derives box k + 1's yaw samples from
 * the cam_pose box k left behind (box_proposal_detail.cpp:126-128 after :237,485).  One pass per box rank: pass r takes the rank-r box of
 * every frame that has one (a CSR with at most one box per frame), runs it with the camera yaw pass r - 1 left for that frame (NaN: the
 * raw pose's), and reports the camera yaw each box leaves.  Frames, and the boxes of one rank, run side by side inside a pass. */
#ifndef CS_CARRIED_H
#define CS_CARRIED_H

#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/cube_slam_b200.h"

/* run_pass(sub_boxes, sub_off, cam_yaw, recs, counts, yaw_left) -> status, where
 *   sub_off   n_frames + 1 offsets, at most one box per frame;        cam_yaw   n_frames camera yaws, NaN = the raw pose's;
 *   recs      n_sub x topk records, counts n_sub;                     yaw_left  n_frames slots, to fill for the frames that have a box in
 *                                                                               this pass; nullptr when nobody will read them (last pass) */
template <class RunPass>
int cs_carried_passes(int n_frames, const double *boxes, const int32_t *box_offsets, int topk, cs_cuboid_rec *out, int32_t *out_counts, RunPass &&run_pass)
{
    int max_rank = 0;
    for (int f = 0; f < n_frames; f++) max_rank = std::max(max_rank, box_offsets[f + 1] - box_offsets[f]);
    std::vector<double> cam_yaw((size_t)n_frames, std::nan("")), yaw_left((size_t)n_frames);
    std::vector<double> sub_boxes;
    std::vector<int32_t> sub_off((size_t)n_frames + 1);
    std::vector<cs_cuboid_rec> recs;
    std::vector<int32_t> counts;
    for (int r = 0; r < max_rank; r++) {
        sub_boxes.clear();
        sub_off[0] = 0;
        for (int f = 0; f < n_frames; f++) {
            const bool has = box_offsets[f + 1] - box_offsets[f] > r;
            if (has) sub_boxes.insert(sub_boxes.end(), boxes + (size_t)(box_offsets[f] + r) * 5, boxes + (size_t)(box_offsets[f] + r) * 5 + 5);
            sub_off[f + 1] = sub_off[f] + (has ? 1 : 0);
        }
        const int n_sub = sub_off[n_frames];
        recs.assign((size_t)n_sub * topk, cs_cuboid_rec());
        counts.assign((size_t)n_sub, 0);
        const bool last = r + 1 == max_rank;
        std::fill(yaw_left.begin(), yaw_left.end(), std::nan(""));
        const int rc = run_pass(sub_boxes.data(), sub_off.data(), cam_yaw, recs.data(), counts.data(), last ? nullptr : &yaw_left);
        if (rc) return rc;
        for (int f = 0; f < n_frames; f++)
            if (sub_off[f + 1] > sub_off[f]) {
                const size_t dst = (size_t)(box_offsets[f] + r), src = (size_t)sub_off[f];
                std::memcpy(&out[dst * topk], &recs[src * topk], sizeof(cs_cuboid_rec) * topk);
                out_counts[dst] = counts[src];
            }
        cam_yaw = yaw_left; /* frames without a box in this pass have none in the next either */
    }
    return CS_OK;
}

#endif /* CS_CARRIED_H */
