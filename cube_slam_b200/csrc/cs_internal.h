/* cs_internal.h -- structures shared by the host orchestration and the sm_100a kernels. */
#ifndef CS_INTERNAL_H
#define CS_INTERNAL_H

#include <stdint.h>

#include "../../include/cube_slam_b200.h"

#define CS_MAX_YAW 512        /* yaw samples per object (reference default 16; dense sweep 181) */
#define CS_MAX_POSE 32        /* roll x pitch samples (reference 5 x 5) */
#define CS_MAX_TOPK 32        /* upper bound of max_cuboid_num handled on device */
#define CS_LINE_CAP 1024      /* lines of one frame that fall inside one ROI (shared-memory resident) */
#define CS_MAXL_OUT 256       /* merged lines kept per ROI */

/* One camera pose hypothesis: detect_3d_cuboid::cam_pose after set_cam_pose
 * (box_proposal_detail.cpp:42-54) plus the ground plane in the sensor frame (:100,238). */
struct CsPose {
    double KinvR[9];
    double T[16];
    double ground[4];
    double roll, pitch;
    double camera_yaw; /* cam_pose.camera_yaw re-derived through the quaternion */
};

/* Per-frame constants */
struct CsFrame {
    double invK[9];
    double euler_raw[3];
    int32_t pose_off;  /* first CsPose of this frame; pose 0 == raw pose when sampling is off */
    int32_t n_pose;
    int32_t yaw_off;   /* first yaw sample of this frame in the yaw table */
    int32_t n_yaw;
    int32_t line_off;  /* CSR into the batch line array */
    int32_t n_lines;
};

/* One (2D box, height sample) ROI job: box_proposal_detail.cpp:107-163 evaluated on the host. */
struct CsJob {
    int32_t frame;
    int32_t obj;        /* global object index */
    int32_t hs;         /* height-sample id */
    int32_t left, top, right;           /* left_x_raw, top_y_raw, right_x_raw */
    int32_t width_raw, height_raw;
    int32_t down_expand, down_y_expan;
    int32_t roi_l, roi_t, roi_r, roi_b; /* dist-map ROI corners (inclusive box test uses these) */
    int32_t roi_w, roi_h;               /* width_expan_distmap, height_expan_distmap */
    int32_t n_top, top_lo, top_hi, top_step, top_override;
    int32_t n_cand;                     /* n_pose * n_yaw * n_top * 2 */
    int32_t tile_off;                   /* first canny tile of this job (prefix over jobs) */
    int32_t tiles_x;
    int32_t bw;                         /* 32-bit words per bit-plane row, ceil(roi_w / 32) */
    int32_t dpitch;                     /* row pitch of the dist map in floats (roi_w rounded up to 4) */
    int64_t bit_off;                    /* offset (words) of this ROI's two bordered bit planes */
    int64_t px_off;                     /* offset (floats) of this ROI's dist map, multiple of 16 */
    int64_t cand_off;                   /* offset into the candidate record arenas */
    double diag;                        /* obj_diaglength_expan */
};

struct CsObj {
    int32_t frame;
    int32_t job_off, n_jobs;
    int32_t left, top, width_raw, height_raw;
};


/* narrow view of the context for the other translation units (the struct itself lives in cs_context.cu) */
#ifdef __CUDACC__
#include <cuda_runtime.h>
cudaStream_t cs_ctx_stream(cs_ctx *c);
#endif
int cs_ctx_device(cs_ctx *c);
int cs_ctx_fail(cs_ctx *c, int code, const char *fmt, ...);
void **cs_ctx_lsd_slot(cs_ctx *c);          /* owned by cs_lsd.cu */
void cs_lsd_destroy(void *state);           /* called from cs_destroy */
int cs_lsd_run_device(cs_ctx *c, const uint8_t *d_imgs, int n_frames, int w, int h, int stride, int channels, float line_length_thres, int cap,
                      const float **d_lines, const int32_t **d_counts);
void cs_ctx_count_launches(cs_ctx *c, int64_t n);
int cs_ctx_seq_lines(cs_ctx *c);
int cs_ctx_use_tma(cs_ctx *c);              /* cs_set_profiling bit 8 clear */            /* cs_set_profiling bit 7 */
void **cs_ctx_edl_slot(cs_ctx *c);          /* owned by cs_edlines.cu */
void cs_edl_destroy(void *state);           /* called from cs_destroy */
/* EDLines flavour of detect_filter_lines: frames on the device (or host, copied in) -> filtered float32 segments + counts in HBM */
int cs_edl_run(cs_ctx *c, const uint8_t *imgs, bool imgs_on_device, int n_frames, int w, int h, int stride, int channels, float line_length_thres,
               int cap, const float **d_lines, const int32_t **d_counts);

/* the same run, also keeping per kept segment what the descriptor needs of its key line (binary_descriptor.cpp:526-540): d_extra holds
 * cap x 2 floats per frame, {KeyLine::angle (lineDirection_), KeyLine::numOfPixels as an integer's bits}, parallel to d_lines */
int cs_edl_run_keylines(cs_ctx *c, const uint8_t *imgs, bool imgs_on_device, int n_frames, int w, int h, int stride, int channels,
                        float line_length_thres, int cap, const float **d_lines, const int32_t **d_counts, const float **d_extra,
                        const int16_t **d_dx, const int16_t **d_dy /* the detector's own Sobel maps, the ones the descriptor reads */);
/* BinaryDescriptor::computeSobel for octave 0 (binary_descriptor.cpp:352-398: GaussianBlur 5 x 5 sigma 1, Sobel 3 x 3 to 16S) = the front
 * end of the EDLines detector: the two int16 maps of every frame, in HBM, owned by the EDLines workspace */
int cs_edl_sobel_maps(cs_ctx *c, const uint8_t *imgs, bool imgs_on_device, int n_frames, int w, int h, int stride, int channels,
                      const int16_t **d_dx, const int16_t **d_dy);
/* LSD flavour of detect_filter_lines from host frames (the body of cs_detect_lines_batch's LSD branch): filtered segments + counts in HBM */
int cs_lsd_run_host(cs_ctx *c, const uint8_t *imgs, int n_frames, int w, int h, int stride, int channels, float line_length_thres, int cap,
                    const float **d_lines, const int32_t **d_counts, const uint8_t **d_frames);
void **cs_ctx_lbd_slot(cs_ctx *c);          /* owned by cs_lbd.cu */
void cs_lbd_destroy(void *state);           /* called from cs_destroy */

#endif
