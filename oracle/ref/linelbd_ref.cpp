/*
 * oracle/ref/linelbd_ref.cpp -- CPU ORACLE, TEST INFRASTRUCTURE ONLY: the reference's own line_lbd_detect::detect_filter_lines, whole.
 *
 * One translation unit made of the reference's detection-side sources, included from where they lie under /root/reference:
 *   line_lbd/libs/lsd.cpp                  the vendored LSD
 *   line_lbd/libs/LSDDetector.cpp          LSDDetector::detect: the octave loop, KeyLine fill, border rejection
 *   line_lbd/libs/binary_descriptor.cpp    BinaryDescriptor::detect with the EDLines detector
 *   line_lbd/class/line_lbd_allclass.cpp   class line_lbd_detect: detect_raw_lines, filter_lines, detect_filter_lines, keylines_to_mat
 * with the reference's own headers, against oracle/ref/fakecv/opencv2/*.hpp -> oracle/ref/minicv.hpp in place of OpenCV (see that file for
 * what the stand-in implements and how it is pinned).  No reference source is copied.  The entry point does what
 * object_slam/src/main_obj.cpp:363-366,428 does: construct line_lbd_detect, set use_LSD and line_length_thres, call
 * detect_filter_lines(image, lines_mat) -- stage (i) of the north star executed by the reference's code from the first line to the last.
 *
 * Two functions are defined here instead of taken from the reference: BinaryDescriptorMatcher::createBinaryDescriptorMatcher and ::match
 * (the class constructor, line_lbd_allclass.cpp:117, makes an LBD matcher that the detection path never touches, match_line_descrip calls
 * it; their translation unit, binary_descriptor_matcher.cpp, is outside the cuboid path and is not compiled).
 */
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <exception>
#include <vector>

#include "/root/reference/line_lbd/libs/lsd.cpp"
#include "/root/reference/line_lbd/libs/LSDDetector.cpp"
#include "/root/reference/line_lbd/libs/binary_descriptor.cpp"
#include "/root/reference/line_lbd/class/line_lbd_allclass.cpp"

namespace cv {
namespace line_descriptor {
Ptr<BinaryDescriptorMatcher> BinaryDescriptorMatcher::createBinaryDescriptorMatcher() { return Ptr<BinaryDescriptorMatcher>(); }
void BinaryDescriptorMatcher::match(const Mat &, const Mat &, std::vector<DMatch> &, const Mat &) const { minicv_unreachable("BinaryDescriptorMatcher::match"); }
}  // namespace line_descriptor
}  // namespace cv

/* The reference reports on std::cout on every call ("BinaryDescriptor line detector reset save octave lines ..."): the C++ stream of this
 * process is switched off when the library loads, so that a host program's own stdout (bench.py prints one JSON line there) stays clean.
 * Python's sys.stdout and C stdio are not affected. */
namespace {
struct SilenceCout {
    SilenceCout() { std::cout.setstate(std::ios_base::failbit); }
} silence_cout;
}  // namespace

/* img: h x w x channels bytes (1 or 3 channels, BGR).  out: room for cap rows [x1 y1 x2 y2].  Returns the number of rows of the n x 4 CV_32F
 * matrix detect_filter_lines produced (-1: exception, message on stderr). */
extern "C" int ref_detect_filter_lines(const uint8_t *img, int w, int h, int channels, int use_LSD, float line_length_thres, float *out, int cap)
{
    try {
        /* one detector object per host thread, constructed once (main_obj.cpp:363 constructs it once per run) */
        thread_local line_lbd_detect det(1, 2.0f);
        det.use_LSD = use_LSD != 0;
        det.line_length_thres = line_length_thres;
        cv::Mat image(h, w, channels == 3 ? CV_8UC3 : CV_8UC1);
        std::memcpy(image.data, img, (size_t)w * h * channels);
        cv::Mat lines;
        det.detect_filter_lines(image, lines);
        if (lines.rows > 0 && (lines.cols != 4 || lines.type() != CV_32FC1)) return -2;
        const int n = lines.rows < cap ? lines.rows : cap;
        if (n) std::memcpy(out, lines.data, sizeof(float) * 4 * (size_t)n);
        return lines.rows;
    } catch (const std::exception &e) {
        fprintf(stderr, "ref_detect_filter_lines: %s\n", e.what());
        return -1;
    }
}
