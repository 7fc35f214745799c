/*
 * shim/line_lbd_b200.cpp -- the detection half of class line_lbd_detect (line_lbd/include/line_lbd/line_lbd_allclass.h:22-70;
 * line_lbd/class/line_lbd_allclass.cpp:110-221) on libcubeslam_b200.so.  In the reference's line_lbd package, compile this file and drop
 * the definitions of the same member functions from line_lbd_allclass.cpp; the descriptor / matcher members (get_line_descriptors,
 * detect_descrip_lines*, match_line_descrip) keep the reference's implementation (LBD is outside the cuboid path).  Callers:
 * object_slam/src/main_obj.cpp:363-366,428 and line_lbd/src/detect_lines.cpp:60-69, unchanged.
 *
 * Guarded like detect_3d_cuboid_b200.cpp: an empty translation unit where OpenCV's C++ headers are absent.
 */
#if defined(__has_include)
#if __has_include(<opencv2/core/core.hpp>) && __has_include("line_lbd/line_lbd_allclass.h")
#define CS_SHIM_ENABLED 1
#endif
#endif

#ifdef CS_SHIM_ENABLED
#include <cmath>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <unordered_map>
#include <vector>

#include "cube_slam_b200.h"
#include "line_lbd/line_lbd_allclass.h"

namespace {
struct CtxTable { /* one context per detector object, see detect_3d_cuboid_b200.cpp */
    std::mutex mu;
    std::unordered_map<const line_lbd_detect *, cs_ctx *> map;
    cs_ctx *get(const line_lbd_detect *self)
    {
        std::lock_guard<std::mutex> g(mu);
        auto it = map.find(self);
        if (it != map.end()) return it->second;
        cs_ctx *c = cs_create(0, 2048, 2048, 1, 64, 8192);
        if (!c) throw std::runtime_error("cube_slam_b200: cs_create failed (no CUDA device?)");
        map.emplace(self, c);
        return c;
    }
};
CtxTable &table()
{
    static CtxTable t;
    return t;
}

/* n x 4 float [x1 y1 x2 y2], with or without the length filter (line_length_thres = 0 keeps every octave-0 segment) */
void run(line_lbd_detect *self, const cv::Mat &gray_img, float length_thres, std::vector<float> &seg, int32_t &n)
{
    cs_ctx *ctx = table().get(self);
    cs_line_params lp;
    cs_default_line_params(&lp);
    lp.use_LSD = self->use_LSD ? 1 : 0;
    lp.line_length_thres = length_thres;
    lp.numoctaves = self->numoctaves_;
    lp.octaveratio = self->octaveratio_;
    const cv::Mat img = gray_img.isContinuous() ? gray_img : gray_img.clone();
    seg.resize(4 * 8192);
    n = 8192;
    const int rc = cs_detect_lines(ctx, img.data, img.cols, img.rows, (int)img.step, img.channels(), &lp, seg.data(), &n);
    if (rc != CS_OK) throw std::runtime_error(cs_last_error(ctx)); /* LSDDetector throws on bad input too (LSDDetector.cpp:112-113,163-164) */
}

/* what the reference's KeyLine carries for a detected segment (LSDDetector.cpp:205-256): enough for filter_lines / keylines_to_mat */
void to_keylines(const std::vector<float> &seg, int n, std::vector<cv::line_descriptor::KeyLine> &out)
{
    out.clear();
    for (int i = 0; i < n; i++) {
        cv::line_descriptor::KeyLine kl;
        kl.startPointX = kl.sPointInOctaveX = seg[4 * i + 0];
        kl.startPointY = kl.sPointInOctaveY = seg[4 * i + 1];
        kl.endPointX = kl.ePointInOctaveX = seg[4 * i + 2];
        kl.endPointY = kl.ePointInOctaveY = seg[4 * i + 3];
        kl.lineLength = std::sqrt((seg[4 * i] - seg[4 * i + 2]) * (seg[4 * i] - seg[4 * i + 2]) + (seg[4 * i + 1] - seg[4 * i + 3]) * (seg[4 * i + 1] - seg[4 * i + 3]));
        kl.angle = std::atan2(kl.endPointY - kl.startPointY, kl.endPointX - kl.startPointX);
        kl.class_id = i;
        kl.octave = 0;
        kl.size = (kl.endPointX - kl.startPointX) * (kl.endPointY - kl.startPointY);
        kl.response = 0;
        kl.pt = cv::Point2f((kl.endPointX + kl.startPointX) / 2, (kl.endPointY + kl.startPointY) / 2);
        out.push_back(kl);
    }
}
}  // namespace

/* line_lbd_allclass.cpp:125-135 (one octave) */
void line_lbd_detect::detect_raw_lines(const cv::Mat &gray_img, std::vector<cv::line_descriptor::KeyLine> &keylines_out)
{
    std::vector<float> seg;
    int32_t n = 0;
    run(this, gray_img, 0.f, seg, n);
    to_keylines(seg, n, keylines_out);
}

/* line_lbd_allclass.cpp:216-221 */
void line_lbd_detect::detect_filter_lines(const cv::Mat &gray_img, cv::Mat &linesmat_out)
{
    std::vector<float> seg;
    int32_t n = 0;
    run(this, gray_img, line_length_thres, seg, n);
    linesmat_out.create(n, 4, CV_32FC1);
    if (n) std::memcpy(linesmat_out.data, seg.data(), sizeof(float) * 4 * (size_t)n);
}

void line_lbd_detect::detect_filter_lines(const cv::Mat &gray_img, std::vector<cv::line_descriptor::KeyLine> &keylines_out)
{
    std::vector<float> seg;
    int32_t n = 0;
    run(this, gray_img, line_length_thres, seg, n);
    to_keylines(seg, n, keylines_out);
}
#endif /* CS_SHIM_ENABLED */
