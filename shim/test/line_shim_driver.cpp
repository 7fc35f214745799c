/*
 * shim/test/line_shim_driver.cpp -- TEST HARNESS for shim/line_lbd_b200.cpp (not product code).
 *
 * Builds the shim the way a maintainer would -- against the REFERENCE'S OWN class header
 * (line_lbd/include/line_lbd/line_lbd_allclass.h:22-70, from /root/reference) and cv::Mat -- and calls it the way
 * object_slam/src/main_obj.cpp:363-366,428 does: construct a line_lbd_detect, set use_LSD / line_length_thres, call
 * detect_filter_lines(gray, lines_mat).  This image has no OpenCV C++ headers, so cv::Mat is oracle/ref/minicv.hpp (a container
 * stand-in, see that file); the class, the shim and libcubeslam_b200.so are the real things.  The one member the harness defines itself is
 * the constructor: the reference's (line_lbd_allclass.cpp:110-123) also creates the LBD descriptor / matcher objects, which live in
 * reference translation units outside the cuboid path.
 *
 * Built by oracle/Makefile (target `ref`, only where the reference checkout exists) into oracle/_ref/libshim_line.so;
 * tests/test_gpu_shim_runs.py loads it on the GPU box.
 */
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <exception>
#include <vector>

#include "line_lbd/line_lbd_allclass.h"

line_lbd_detect::line_lbd_detect(int numoctaves, float octaveratio) : numoctaves_(numoctaves), octaveratio_(octaveratio)
{
    use_LSD = false;        /* the class defaults, line_lbd_allclass.cpp:121-122 */
    line_length_thres = 50;
}

/* img: h x w x channels bytes.  out: room for cap rows of [x1 y1 x2 y2].  Returns the number of rows of the n x 4 CV_32F matrix the
 * shim's detect_filter_lines produced, or -1 (message on stderr). */
extern "C" int shim_line_detect_filter(const uint8_t *img, int w, int h, int channels, int use_LSD, float line_length_thres, float *out, int cap)
{
    try {
        line_lbd_detect det(1, 2.0f); /* main_obj.cpp:363: line_lbd_detect line_lbd_obj (one octave) */
        det.use_LSD = use_LSD != 0;
        det.line_length_thres = line_length_thres;
        cv::Mat gray(h, w, channels == 3 ? CV_8UC3 : CV_8UC1);
        std::memcpy(gray.data, img, (size_t)w * h * channels);
        cv::Mat lines;
        det.detect_filter_lines(gray, lines);
        if (lines.rows > 0 && (lines.cols != 4 || lines.type() != CV_32FC1)) return -2;
        const int n = lines.rows < cap ? lines.rows : cap;
        if (n) std::memcpy(out, lines.data, sizeof(float) * 4 * (size_t)n);
        return lines.rows;
    } catch (const std::exception &e) {
        fprintf(stderr, "shim_line_detect_filter: %s\n", e.what());
        return -1;
    }
}
