/*
 * oracle/ref/edl_ref.cpp -- CPU ORACLE, TEST INFRASTRUCTURE ONLY: the reference's own EDLines detector, compiled from /root/reference.
 *
 * One translation unit: the reference's line_lbd/libs/binary_descriptor.cpp (BinaryDescriptor with its nested EDLineDetector: edge
 * drawing, least-squares line fitting, the Helmholtz validation, the octave bookkeeping of detectImpl) included from where it lies,
 * together with its own headers (precomp.hpp, line_lbd/line_descriptor/descriptor.hpp), against oracle/ref/fakecv/opencv2/*.hpp ->
 * oracle/ref/minicv.hpp in place of OpenCV.  No reference source is copied.  The entry point drives it the way
 * line_lbd_detect::detect_raw_lines does for use_LSD = false (line_lbd/class/line_lbd_allclass.cpp:110-124,165-169):
 * BinaryDescriptor::createBinaryDescriptor(params with numOfOctave_ = 1, Octave_ratio = 2.0)->detect(gray, keylines, mask of ones).
 *
 * What is NOT the reference here: the OpenCV primitives the path calls (GaussianBlur 5 x 5 on 8-bit, Sobel 3 x 3 to 16-bit, abs, add,
 * threshold, compare, Mat / 4) -- minicv.hpp implements them the way cv2 computes them, and tests/test_oracle_cv_parity.py /
 * tests/golden/cv_pins.npz pin those semantics against the in-container cv2.
 */
#include <cstdint>
#include <cstdio>
#include <iostream>
#include <vector>

#ifndef CS_REFERENCE_BINARY_DESCRIPTOR_CPP
#define CS_REFERENCE_BINARY_DESCRIPTOR_CPP "/root/reference/line_lbd/libs/binary_descriptor.cpp"
#endif
#include CS_REFERENCE_BINARY_DESCRIPTOR_CPP

/* gray: h x w bytes.  lines_out: room for cap segments (x1 y1 x2 y2 floats, the key line's start / end point in the input image, octave 0).
 * Returns the number of key lines the reference found (before detect_filter_lines' length filter). */
extern "C" int ref_edl_detect(const uint8_t *gray, int w, int h, float *lines_out, int cap)
{
    using namespace cv;
    using namespace cv::line_descriptor;
    try {
        BinaryDescriptor::Params params;
        params.numOfOctave_ = 1;
        params.Octave_ratio = 2.0;
        Ptr<BinaryDescriptor> lbd = BinaryDescriptor::createBinaryDescriptor(params);
        Mat img(h, w, CV_8UC1);
        memcpy(img.data, gray, (size_t)w * h);
        Mat mask = Mat::ones(img.size(), CV_8UC1);
        std::vector<KeyLine> keylines;
        lbd->detect(img, keylines, mask);
        int n = 0;
        for (const KeyLine &kl : keylines) {
            if (kl.octave != 0) continue;
            if (n < cap) {
                lines_out[4 * n + 0] = kl.startPointX;
                lines_out[4 * n + 1] = kl.startPointY;
                lines_out[4 * n + 2] = kl.endPointX;
                lines_out[4 * n + 3] = kl.endPointY;
            }
            n++;
        }
        return n;
    } catch (const std::exception &e) {
        fprintf(stderr, "ref_edl_detect: %s\n", e.what());
        return -1;
    }
}
