/*
 * oracle/ref/lsd_ref.cpp -- CPU ORACLE, TEST INFRASTRUCTURE ONLY: the reference's own LSD, compiled from /root/reference.
 *
 * This translation unit pre-empts the include guard of line_lbd/libs/precomp.hpp (which would pull in OpenCV's C++ headers, absent from this
 * image), brings in oracle/ref/minicv.hpp instead, restates the two declarations lsd.cpp expects from
 * line_lbd/include/line_lbd/line_descriptor/descriptor.hpp (the LSD_REFINE_* constants :896-902 and the abstract class
 * LineSegmentDetector :916-966 -- interface only), and then includes the reference's lsd.cpp from where it lies.  No reference source is
 * copied.  Built by `make ref` in oracle/ into oracle/_ref/liblsd_ref.so (git-ignored; it travels to the GPU box with the snapshot).
 *
 * ref_lsd_detect() is the call LSDDetector::detectImpl makes for octave 0 (line_lbd/libs/LSDDetector.cpp:173,188):
 * createLineSegmentDetector(LSD_REFINE_ADV)->detect(gray, lines).
 */
#define __OPENCV_PRECOMP_H__
#include <iostream>
#include <vector>

#include "minicv.hpp"

namespace cv {
namespace line_descriptor {
enum { LSD_REFINE_NONE = 0, LSD_REFINE_STD = 1, LSD_REFINE_ADV = 2 };
class LineSegmentDetector : public Algorithm {
public:
    virtual void detect(InputArray _image, OutputArray _lines, OutputArray width = noArray(), OutputArray prec = noArray(), OutputArray nfa = noArray()) = 0;
    virtual void drawSegments(InputOutputArray _image, InputArray lines) = 0;
    virtual int compareSegments(const Size &size, InputArray lines1, InputArray lines2, InputOutputArray _image = noArray()) = 0;
    virtual ~LineSegmentDetector() {}
};
Ptr<LineSegmentDetector> createLineSegmentDetector(int _refine = LSD_REFINE_STD, double _scale = 0.8, double _sigma_scale = 0.6, double _quant = 2.0,
                                                   double _ang_th = 22.5, double _log_eps = 0, double _density_th = 0.7, int _n_bins = 1024);
}  // namespace line_descriptor
}  // namespace cv

#ifndef CS_REFERENCE_LSD_CPP
#define CS_REFERENCE_LSD_CPP "/root/reference/line_lbd/libs/lsd.cpp"
#endif
#include CS_REFERENCE_LSD_CPP

/* gray: h x w bytes.  lines_out: room for cap segments (x1 y1 x2 y2 floats).  Returns the number of segments the reference found. */
extern "C" int ref_lsd_detect(const unsigned char *gray, int w, int h, float *lines_out, int cap)
{
    try {
        cv::Mat img(h, w, CV_8UC1);
        memcpy(img.data, gray, (size_t)w * h);
        cv::Ptr<cv::line_descriptor::LineSegmentDetector> ls = cv::line_descriptor::createLineSegmentDetector(cv::line_descriptor::LSD_REFINE_ADV);
        std::vector<cv::Vec4f> lines;
        ls->detect(img, lines);
        const int n = (int)lines.size();
        for (int i = 0; i < n && i < cap; i++)
            for (int k = 0; k < 4; k++) lines_out[4 * i + k] = lines[i][k];
        return n;
    } catch (const std::exception &e) {
        fprintf(stderr, "ref_lsd_detect: %s\n", e.what());
        return -1;
    }
}
