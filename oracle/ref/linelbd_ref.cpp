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
 *   line_lbd/libs/binary_descriptor_matcher.cpp   BinaryDescriptorMatcher (multi-index hashing), for match_line_descrip
 * The descriptor half runs from the same translation unit: ref_detect_descrip_lines is
 * line_lbd_detect::detect_descrip_lines(gray, keylines_out, line_descrips) (line_lbd_allclass.cpp:253-272), ref_lbd_compute is
 * lbd->compute(image, keylines, descriptors[, returnFloatDescr]) as get_line_descriptors calls it (:191-198), ref_match_line_descrip is
 * match_line_descrip (:341-356).
 */
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <exception>
#include <vector>
/* The reference calls cos / sin / round / atan2 unqualified on float arguments (binary_descriptor.cpp:1251-1252,1282, LSDDetector.cpp:244).
 * Which libm function that is depends on what the translation unit has seen: with <math.h> (libstdc++'s wrapper puts the float overloads
 * into the global namespace) it is cosf / sinf / roundf / atan2f, with <cmath> alone the double function.  lsd.cpp pulls <math.h> in for
 * this translation unit anyway; it is included here explicitly so that the choice does not hang on include order. */
#include <math.h>

#include "/root/reference/line_lbd/libs/lsd.cpp"
#include "/root/reference/line_lbd/libs/LSDDetector.cpp"
#include "/root/reference/line_lbd/libs/binary_descriptor.cpp"
#include "/root/reference/line_lbd/class/line_lbd_allclass.cpp"
#undef MAX_B
#include "/root/reference/line_lbd/libs/binary_descriptor_matcher.cpp"

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

/* detect_filter_lines of a detector built with MORE THAN ONE octave (line_lbd_detect(numoctaves, 2.0f), as line_lbd/src/detect_lines.cpp:57-60
 * parameterises it): filter_lines keeps octave 0 only (:200-207), so the matrix must be the one-octave one -- the claim the product's
 * acceptance of numoctaves > 1 rests on (tests/test_oracle_ref_octaves.py) */
extern "C" int ref_detect_filter_lines_octaves(const uint8_t *img, int w, int h, int channels, int use_LSD, float line_length_thres, int numoctaves,
                                               float *out, int cap, int *n_raw_all_octaves)
{
    try {
        line_lbd_detect det(numoctaves, 2.0f);
        det.use_LSD = use_LSD != 0;
        det.line_length_thres = line_length_thres;
        cv::Mat image(h, w, channels == 3 ? CV_8UC3 : CV_8UC1);
        std::memcpy(image.data, img, (size_t)w * h * channels);
        std::vector<KeyLine> raw;
        det.detect_raw_lines(image, raw);
        if (n_raw_all_octaves) *n_raw_all_octaves = (int)raw.size();
        cv::Mat lines;
        det.detect_filter_lines(image, lines);
        const int n = lines.rows < cap ? lines.rows : cap;
        if (n) std::memcpy(out, lines.data, sizeof(float) * 4 * (size_t)n);
        return lines.rows;
    } catch (const std::exception &e) {
        fprintf(stderr, "ref_detect_filter_lines_octaves: %s\n", e.what());
        return -1;
    }
}

/* the stand-in's pyrDown, for pinning against cv2 */
extern "C" void ref_minicv_pyrdown(const uint8_t *src, int w, int h, uint8_t *dst, int dw, int dh)
{
    cv::Mat a(h, w, CV_8UC1), b;
    std::memcpy(a.data, src, (size_t)w * h);
    cv::pyrDown(a, b, cv::Size(dw, dh));
    std::memcpy(dst, b.data, (size_t)dw * dh);
}
extern "C" void ref_minicv_resize_half(const uint8_t *src, int w, int h, uint8_t *dst)
{
    cv::Mat a(h, w, CV_8UC1), b;
    std::memcpy(a.data, src, (size_t)w * h);
    cv::resize(a, b, cv::Size(), 0.5, 0.5);
    std::memcpy(dst, b.data, (size_t)b.cols * b.rows);
}

/* the KeyLine fields anything downstream reads, octave 0 (same layout as oracle/lbd_oracle.cpp's lbd_keyline) */
struct ref_keyline {
    float sx, sy, ex, ey, angle, line_length, response, size;
    int32_t num_pixels, class_id;
};

static line_lbd_detect &ref_detector(int use_LSD, float line_length_thres)
{
    thread_local line_lbd_detect det(1, 2.0f);
    det.use_LSD = use_LSD != 0;
    det.line_length_thres = line_length_thres;
    return det;
}

/* detect_descrip_lines(gray_img, keylines_out, line_descrips): kept key lines and their 32-byte descriptors.  Returns the count (-1: exception). */
extern "C" int ref_detect_descrip_lines(const uint8_t *img, int w, int h, int channels, int use_LSD, float line_length_thres, ref_keyline *kl_out,
                                        uint8_t *desc_out, int cap)
{
    try {
        line_lbd_detect &det = ref_detector(use_LSD, line_length_thres);
        cv::Mat image(h, w, channels == 3 ? CV_8UC3 : CV_8UC1);
        std::memcpy(image.data, img, (size_t)w * h * channels);
        std::vector<KeyLine> keylines;
        cv::Mat descrips;
        det.detect_descrip_lines(image, keylines, descrips);
        const int n = (int)keylines.size();
        if (n > 0 && (descrips.rows != n || descrips.cols != 32 || descrips.type() != CV_8UC1)) return -2;
        for (int i = 0; i < n && i < cap; i++) {
            const KeyLine &k = keylines[i];
            if (k.octave != 0) return -4;
            ref_keyline &o = kl_out[i];
            o.sx = k.sPointInOctaveX;
            o.sy = k.sPointInOctaveY;
            o.ex = k.ePointInOctaveX;
            o.ey = k.ePointInOctaveY;
            if (o.sx != k.startPointX || o.sy != k.startPointY || o.ex != k.endPointX || o.ey != k.endPointY) return -5; /* octave 0 */
            o.angle = k.angle;
            o.line_length = k.lineLength;
            o.response = k.response;
            o.size = k.size;
            o.num_pixels = k.numOfPixels;
            o.class_id = k.class_id;
            std::memcpy(desc_out + (size_t)i * 32, descrips.ptr(i), 32);
        }
        return n;
    } catch (const std::exception &e) {
        fprintf(stderr, "ref_detect_descrip_lines: %s\n", e.what());
        return -1;
    }
}

/* detect_descrip_lines_octaves(gray_img, keylines_out, line_descrips) (line_lbd_allclass.cpp:285-339), octave 0 of the one-octave detector */
extern "C" int ref_detect_descrip_lines_octaves(const uint8_t *img, int w, int h, int channels, int use_LSD, float line_length_thres, ref_keyline *kl_out,
                                                uint8_t *desc_out, int cap)
{
    try {
        line_lbd_detect &det = ref_detector(use_LSD, line_length_thres);
        cv::Mat image(h, w, channels == 3 ? CV_8UC3 : CV_8UC1);
        std::memcpy(image.data, img, (size_t)w * h * channels);
        std::vector<std::vector<KeyLine>> keylines;
        std::vector<cv::Mat> descrips;
        det.detect_descrip_lines_octaves(image, keylines, descrips);
        if (keylines.size() != 1 || descrips.size() != 1) return -2;
        const int n = (int)keylines[0].size();
        if (n > 0 && (descrips[0].rows != n || descrips[0].cols != 32)) return -3;
        for (int i = 0; i < n && i < cap; i++) {
            const KeyLine &k = keylines[0][i];
            ref_keyline &o = kl_out[i];
            o.sx = k.startPointX;
            o.sy = k.startPointY;
            o.ex = k.endPointX;
            o.ey = k.endPointY;
            if (o.sx != k.sPointInOctaveX || o.sy != k.sPointInOctaveY || o.ex != k.ePointInOctaveX || o.ey != k.ePointInOctaveY) return -5;
            o.angle = k.angle;
            o.line_length = k.lineLength;
            o.response = k.response;
            o.size = k.size;
            o.num_pixels = k.numOfPixels;
            o.class_id = k.class_id;
            std::memcpy(desc_out + (size_t)i * 32, descrips[0].ptr(i), 32);
        }
        return n;
    } catch (const std::exception &e) {
        fprintf(stderr, "ref_detect_descrip_lines_octaves: %s\n", e.what());
        return -1;
    }
}

/* BinaryDescriptor::compute on caller-given key lines (octave 0, class_id as given): desc n x 32 bytes and / or fdesc n x 72 floats */
extern "C" int ref_lbd_compute(const uint8_t *img, int w, int h, int channels, const ref_keyline *kl, int n, uint8_t *desc, float *fdesc)
{
    try {
        line_lbd_detect &det = ref_detector(1, 0);
        cv::Mat image(h, w, channels == 3 ? CV_8UC3 : CV_8UC1);
        std::memcpy(image.data, img, (size_t)w * h * channels);
        std::vector<KeyLine> keylines(n);
        for (int i = 0; i < n; i++) {
            KeyLine &k = keylines[i];
            k.startPointX = k.sPointInOctaveX = kl[i].sx;
            k.startPointY = k.sPointInOctaveY = kl[i].sy;
            k.endPointX = k.ePointInOctaveX = kl[i].ex;
            k.endPointY = k.ePointInOctaveY = kl[i].ey;
            k.angle = kl[i].angle;
            k.lineLength = kl[i].line_length;
            k.response = kl[i].response;
            k.size = kl[i].size;
            k.numOfPixels = kl[i].num_pixels;
            k.class_id = kl[i].class_id;
            k.octave = 0;
        }
        for (int pass = 0; pass < 2; pass++) {
            if (pass == 0 ? !desc : !fdesc) continue;
            cv::Mat d;
            det.lbd->compute(image, keylines, d, pass == 1);
            if (d.rows != n) return -2;
            if (pass == 0)
                std::memcpy(desc, d.data, (size_t)n * 32);
            else
                std::memcpy(fdesc, d.data, (size_t)n * 72 * sizeof(float));
        }
        return n;
    } catch (const std::exception &e) {
        fprintf(stderr, "ref_lbd_compute: %s\n", e.what());
        return -1;
    }
}

/* match_line_descrip(query, train, good_matches, thres): returns the number of good matches */
extern "C" int ref_match_line_descrip(const uint8_t *q, int nq, const uint8_t *t, int nt, float thres, int32_t *query_idx, int32_t *train_idx, float *dist)
{
    try {
        line_lbd_detect &det = ref_detector(1, 0);
        cv::Mat mq(nq, 32, CV_8UC1), mt(nt, 32, CV_8UC1);
        if (nq) std::memcpy(mq.data, q, (size_t)nq * 32);
        if (nt) std::memcpy(mt.data, t, (size_t)nt * 32);
        std::vector<cv::DMatch> good;
        det.match_line_descrip(mq, mt, good, thres);
        for (size_t i = 0; i < good.size(); i++) {
            query_idx[i] = good[i].queryIdx;
            train_idx[i] = good[i].trainIdx;
            dist[i] = good[i].distance;
        }
        return (int)good.size();
    } catch (const std::exception &e) {
        fprintf(stderr, "ref_match_line_descrip: %s\n", e.what());
        return -1;
    }
}
