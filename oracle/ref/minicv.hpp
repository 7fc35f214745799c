/*
 * oracle/ref/minicv.hpp -- CPU ORACLE, TEST INFRASTRUCTURE ONLY.
 *
 * The smallest stand-in for the OpenCV C++ API that lets the reference's OWN line_lbd/libs/lsd.cpp compile unmodified from where it lies
 * (/root/reference; this image has no OpenCV C++ headers).  oracle/ref/lsd_ref.cpp defines the include guard of the reference's
 * precomp.hpp, includes this file and then the reference source: the seed loop, region growing, rectangle fitting, refinement and NFA
 * test that end up in oracle/_ref/liblsd_ref.so are the reference's code, not a restatement.
 *
 * What is NOT the reference here, and how it is pinned instead: the three OpenCV primitives lsd.cpp calls -- cv::GaussianBlur (7 x 7,
 * CV_64F), cv::resize (x 0.8, INTER_LINEAR, CV_64F) and cv::fastAtan2 -- are forwarded to the oracle's restatements
 * (lsd_orc_gaussian7 / lsd_orc_resize / lsd_orc_fast_atan2 in oracle/lsd_oracle.cpp), which tests/test_oracle_lines.py checks bit for
 * bit against the in-container cv2 4.13.  Everything else below is containers: Mat as a contiguous row-major buffer, Point, Vec, Size,
 * Input/OutputArray as thin views, and abort() stubs for the drawing helpers at the end of lsd.cpp that the path never calls.
 */
#ifndef ORC_MINICV_HPP
#define ORC_MINICV_HPP

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

extern "C" void lsd_orc_gaussian7(const double *src, int w, int h, double *dst);
extern "C" void lsd_orc_resize(const double *src, int w, int h, double scale, double *dst, int *dw, int *dh);
extern "C" float lsd_orc_fast_atan2(float y, float x);
extern "C" void edl_orc_gaussian5_u8(const uint8_t *src, int w, int h, uint8_t *dst);
extern "C" void orc_canny(const uint8_t *src, int w, int h, int stride, double low, double high, uint8_t *dst);
extern "C" void orc_chamfer_dt(const uint8_t *edges, int w, int h, float *dist);
extern "C" void orc_bgr2gray(const uint8_t *bgr, int w, int h, int stride, uint8_t *gray, int gstride, int fixed15);

#define CV_PI 3.1415926535897932384626433832795
#define CV_EXPORTS
#define CV_WRAP
#define CV_OUT
#define CV_IN_OUT
#define CV_8U 0
#define CV_8S 1
#define CV_16S 3
#define CV_32F 5
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_8SC1 CV_MAKETYPE(CV_8S, 1)
#define CV_16SC1 CV_MAKETYPE(CV_16S, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC4 CV_MAKETYPE(CV_32F, 4)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_BGR2GRAY 6
#define CV_DIST_L2 2
#define CV_AA 16
#define CV_Assert(expr)                                                                     \
    do {                                                                                    \
        if (!(expr)) throw std::runtime_error(std::string("CV_Assert failed: ") + #expr);  \
    } while (0)

namespace cv {
typedef unsigned char uchar;

template <typename T>
struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
};
typedef Point_<int> Point2i;
typedef Point2i Point;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;

template <typename T, int N>
struct Vec {
    T val[N];
    Vec() { for (int i = 0; i < N; i++) val[i] = T(0); }
    Vec(T a, T b, T c, T d) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
    T &operator[](int i) { return val[i]; }
    const T &operator[](int i) const { return val[i]; }
};
typedef Vec<float, 4> Vec4f;
typedef Vec<int, 4> Vec4i;

struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
    int area() const { return width * height; }
    bool operator==(const Size &o) const { return width == o.width && height == o.height; }
    bool operator!=(const Size &o) const { return !(*this == o); }
};

template <typename T>
struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T x_, T y_, T w_, T h_) : x(x_), y(y_), width(w_), height(h_) {}
    T area() const { return width * height; }
    Point_<T> tl() const { return Point_<T>(x, y); }
    Point_<T> br() const { return Point_<T>(x + width, y + height); }
};
template <typename T> Rect_<T> operator&(const Rect_<T> &a, const Rect_<T> &b)
{
    const T x1 = std::max(a.x, b.x), y1 = std::max(a.y, b.y), x2 = std::min(a.x + a.width, b.x + b.width), y2 = std::min(a.y + a.height, b.y + b.height);
    return (x2 <= x1 || y2 <= y1) ? Rect_<T>() : Rect_<T>(x1, y1, x2 - x1, y2 - y1);
}
template <typename T> Rect_<T> operator|(const Rect_<T> &a, const Rect_<T> &b)
{
    const T x1 = std::min(a.x, b.x), y1 = std::min(a.y, b.y), x2 = std::max(a.x + a.width, b.x + b.width), y2 = std::max(a.y + a.height, b.y + b.height);
    return Rect_<T>(x1, y1, x2 - x1, y2 - y1);
}
typedef Rect_<int> Rect;

struct Scalar {
    double val[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
    static Scalar all(double v) { return Scalar(v, v, v, v); }
};

template <typename T> struct DepthOf;
template <> struct DepthOf<uchar> { enum { value = CV_8U }; };
template <> struct DepthOf<signed char> { enum { value = CV_8S }; };
template <> struct DepthOf<short> { enum { value = CV_16S }; };
template <> struct DepthOf<int> { enum { value = 4 /* CV_32S */ }; };
template <> struct DepthOf<float> { enum { value = CV_32F }; };
template <> struct DepthOf<double> { enum { value = CV_64F }; };

inline size_t depth_bytes(int type)
{
    switch (type & 7) {
        case CV_8U: return 1;
        case CV_8S: return 1;
        case CV_16S: return 2;
        case 4: return 4; /* CV_32S */
        case CV_32F: return 4;
        case CV_64F: return 8;
        default: throw std::runtime_error("minicv: unsupported depth");
    }
}

class _OutputArray;
class Mat {
public:
    int rows, cols;
    uchar *data;
    size_t step = 0; /* bytes per row (always continuous here) */
    Mat() : rows(0), cols(0), data(nullptr), type_(CV_8UC1) {}
    Mat(int r, int c, int type) : rows(0), cols(0), data(nullptr), type_(type) { create(r, c, type); }
    Mat(Size s, int type, const Scalar &v) : rows(0), cols(0), data(nullptr), type_(type)
    {
        create(s.height, s.width, type);
        if (v.val[0] != 0) throw std::runtime_error("minicv: Mat(size, type, value) is provided for 0 only");
    }
    /* one row as a matrix of its own, and appending rows: the LBD descriptor helpers of line_lbd_allclass.cpp (outside the cuboid path) */
    Mat row(int r) const
    {
        Mat m(1, cols, type_);
        memcpy(m.data, ptr(r), (size_t)cols * elemSize());
        return m;
    }
    void push_back(const Mat &m)
    {
        if (empty()) {
            *this = m.clone();
            return;
        }
        if (m.cols != cols || m.type() != type_) throw std::runtime_error("minicv: push_back of a different row type");
        Mat o(rows + m.rows, cols, type_);
        memcpy(o.data, data, total() * elemSize());
        memcpy(o.data + total() * elemSize(), m.data, m.total() * m.elemSize());
        *this = o;
    }
    explicit Mat(const std::vector<Vec4f> &v) : rows(0), cols(0), data(nullptr), type_(CV_32FC4)
    {
        create((int)v.size(), 1, CV_32FC4);
        if (!v.empty()) memcpy(data, v.data(), v.size() * sizeof(Vec4f));
    }
    explicit Mat(const std::vector<double> &v) : rows(0), cols(0), data(nullptr), type_(CV_64FC1)
    {
        create((int)v.size(), 1, CV_64FC1);
        if (!v.empty()) memcpy(data, v.data(), v.size() * sizeof(double));
    }
    void create(int r, int c, int type)
    {
        if (r == rows && c == cols && type == type_ && data) return;
        rows = r;
        cols = c;
        type_ = type;
        store_ = std::make_shared<std::vector<uchar>>((size_t)r * c * elemSize() + 64, 0);
        data = store_->data();
        step = (size_t)c * elemSize();
    }
    void create(Size s, int type) { create(s.height, s.width, type); }
    bool empty() const { return data == nullptr || rows * cols == 0; }
    void release()
    {
        rows = cols = 0;
        data = nullptr;
        step = 0;
        store_.reset();
    }
    int type() const { return type_; }
    int depth() const { return type_ & 7; }
    int channels() const { return (type_ >> 3) + 1; }
    size_t elemSize() const { return depth_bytes(type_) * (size_t)channels(); }
    size_t total() const { return (size_t)rows * cols; }
    Size size() const { return Size(cols, rows); }
    bool isContinuous() const { return true; }
    uchar *ptr(int r = 0) { return data + (size_t)r * cols * elemSize(); }
    const uchar *ptr(int r = 0) const { return data + (size_t)r * cols * elemSize(); }
    template <typename T> T *ptr(int r = 0) { return reinterpret_cast<T *>(ptr(r)); }
    template <typename T> const T *ptr(int r = 0) const { return reinterpret_cast<const T *>(ptr(r)); }
    template <typename T> T &at(int i) { return reinterpret_cast<T *>(data)[i]; }
    template <typename T> const T &at(int i) const { return reinterpret_cast<const T *>(data)[i]; }
    /* at(row, col): indices are clamped into the matrix.  The reference reads dist_map.at<float>(int(y), int(x)) for box corners that may sit
     * on the inclusive border of the ROI (object_3d_util.cpp:441), one past the last row / column: undefined there, the last row / column
     * here (the behaviour oracle/cuboid_oracle.cpp defines too). */
    template <typename T> T &at(int r, int c) { return reinterpret_cast<T *>(data)[(size_t)clamp_(r, rows) * cols + clamp_(c, cols)]; }
    template <typename T> const T &at(int r, int c) const { return reinterpret_cast<const T *>(data)[(size_t)clamp_(r, rows) * cols + clamp_(c, cols)]; }
    static int clamp_(int v, int n) { return v < 0 ? 0 : (v >= n ? n - 1 : v); }
    Mat operator()(const Rect &r) const /* a COPY of the region (the callers only read it) */
    {
        if (r.x < 0 || r.y < 0 || r.x + r.width > cols || r.y + r.height > rows) throw std::runtime_error("minicv: ROI outside the image");
        Mat m(r.height, r.width, type_);
        const size_t es = elemSize();
        for (int y = 0; y < r.height; y++) memcpy(m.data + (size_t)y * r.width * es, data + ((size_t)(r.y + y) * cols + r.x) * es, (size_t)r.width * es);
        return m;
    }
    static Mat ones(Size s, int type)
    {
        Mat m(s.height, s.width, type);
        if ((type & 7) != CV_8U) throw std::runtime_error("minicv: Mat::ones is provided for 8-bit only");
        memset(m.data, 1, m.total() * m.elemSize());
        return m;
    }
    static Mat zeros(int r, int c, int type) { return Mat(r, c, type); } /* create() zero-fills */
    void setTo(int v)
    {
        if (v != 0) throw std::runtime_error("minicv: setTo is provided for 0 only");
        if (data) memset(data, 0, total() * elemSize());
    }
    Mat t() const; /* transpose (float / double single channel) */
    int checkVector(int) const { return -1; }
    Mat clone() const
    {
        Mat m(rows, cols, type_);
        if (!empty()) memcpy(m.data, data, total() * elemSize());
        return m;
    }
    /* depth conversion between the two types the path uses (8U -> 64F, or a copy) */
    void convertTo(Mat &dst, int rtype) const
    {
        if (channels() != 1) throw std::runtime_error("minicv: convertTo of a multi-channel Mat");
        Mat out(rows, cols, CV_MAKETYPE(rtype & 7, 1));
        const size_t n = total();
        if (depth() == (rtype & 7))
            memcpy(out.data, data, n * elemSize());
        else if (depth() == CV_8U && (rtype & 7) == CV_64F)
            for (size_t i = 0; i < n; i++) out.ptr<double>()[i] = (double)data[i];
        else if (depth() == 4 && (rtype & 7) == CV_32F) /* CV_32S -> CV_32F (EDLineDetector::InitEDLine_ sizes its float matrices this way) */
            for (size_t i = 0; i < n; i++) out.ptr<float>()[i] = (float)ptr<int>()[i];
        else
            throw std::runtime_error("minicv: unsupported conversion");
        dst = out;
    }
    void copyTo(const _OutputArray &dst) const;

protected:
    int type_;
    std::shared_ptr<std::vector<uchar>> store_;
};

template <typename T>
class Mat_ : public Mat {
public:
    struct Line { /* one row or column of the matrix, for setTo */
        T *p;
        size_t stride;
        int n;
        void setTo(double v)
        {
            for (int i = 0; i < n; i++) p[(size_t)i * stride] = (T)v;
        }
    };
    Mat_() : Mat() { type_ = CV_MAKETYPE(DepthOf<T>::value, 1); }
    Mat_(int r, int c) : Mat(r, c, CV_MAKETYPE(DepthOf<T>::value, 1)) {}
    explicit Mat_(Size s) : Mat(s.height, s.width, CV_MAKETYPE(DepthOf<T>::value, 1)) {}
    Mat_(const Mat &m) : Mat() { assign(m); } /* converts the depth like cv::Mat_<T>(const Mat&) does */
    Mat_ &operator=(const Mat &m)
    {
        assign(m);
        return *this;
    }
    static Mat_ zeros(Size s) { return Mat_(s); } /* create() zero-fills */
    T &operator()(int r, int c) { return reinterpret_cast<T *>(data)[(size_t)r * cols + c]; }
    T *operator[](int r) { return reinterpret_cast<T *>(data) + (size_t)r * cols; }
    const T *operator[](int r) const { return reinterpret_cast<const T *>(data) + (size_t)r * cols; }
    Line row(int r) { return Line{reinterpret_cast<T *>(data) + (size_t)r * cols, 1, cols}; }
    Line col(int c) { return Line{reinterpret_cast<T *>(data) + c, (size_t)cols, rows}; }
    using Mat::ptr;

private:
    void assign(const Mat &m)
    {
        const int want = CV_MAKETYPE(DepthOf<T>::value, 1);
        if (m.empty()) {
            static_cast<Mat &>(*this) = Mat();
            type_ = want;
        } else if (m.type() == want)
            static_cast<Mat &>(*this) = m;
        else {
            Mat tmp;
            m.convertTo(tmp, want);
            static_cast<Mat &>(*this) = tmp;
        }
    }
};

class _InputArray {
public:
    _InputArray() : m_(nullptr), v_(nullptr) {}
    _InputArray(const Mat &m) : m_(const_cast<Mat *>(&m)), v_(nullptr) {}
    _InputArray(const std::vector<Vec4f> &v) : m_(nullptr), v_(const_cast<std::vector<Vec4f> *>(&v)) {}
    Mat getMat() const { return m_ ? *m_ : (v_ ? Mat(*v_) : Mat()); }
    bool empty() const { return m_ ? m_->empty() : (v_ ? v_->empty() : true); }
    int channels() const { return m_ ? m_->channels() : 4; }
    Size size() const { return m_ ? m_->size() : Size(); }
    bool needed() const { return m_ != nullptr || v_ != nullptr; }

protected:
    Mat *m_;
    std::vector<Vec4f> *v_;
};
class _OutputArray : public _InputArray {
public:
    _OutputArray() {}
    _OutputArray(Mat &m) : _InputArray(m) {}
    template <typename T> _OutputArray(Mat_<T> &m) : _InputArray(m), typed_(CV_MAKETYPE(DepthOf<T>::value, 1)) {}
    _OutputArray(std::vector<Vec4f> &v) : _InputArray(v) {}
    Mat &getMatRef() const
    {
        if (!m_) throw std::runtime_error("minicv: getMatRef of a non-Mat array");
        return *m_;
    }
    void assign(const Mat &src) const
    {
        if (v_) {
            if (src.type() != CV_32FC4) throw std::runtime_error("minicv: only Vec4f vectors are supported as output");
            v_->resize(src.total());
            if (src.total()) memcpy(v_->data(), src.data, src.total() * sizeof(Vec4f));
        } else if (m_) {
            if (typed_ >= 0 && src.type() != typed_) throw std::runtime_error("minicv: output type mismatch");
            static_cast<Mat &>(*m_) = src;
        }
    }

private:
    int typed_ = -1;
};
typedef const _InputArray &InputArray;
typedef const _OutputArray &OutputArray;
typedef const _OutputArray &InputOutputArray;
inline const _OutputArray &noArray()
{
    static _OutputArray none;
    return none;
}
inline void Mat::copyTo(const _OutputArray &dst) const { dst.assign(clone()); }

inline Mat Mat::t() const
{
    Mat o(cols, rows, type_);
    const size_t es = elemSize();
    for (int r = 0; r < rows; r++)
        for (int c = 0; c < cols; c++) memcpy(o.data + ((size_t)c * rows + r) * es, data + ((size_t)r * cols + c) * es, es);
    return o;
}

/* declarations only: parameter persistence is never exercised */
class FileNode {
public:
    operator int() const { return 0; }
    template <typename T> void operator>>(T &) const {}
    FileNode operator[](const char *) const { return FileNode(); }
    bool empty() const { return true; }
};
class FileStorage {
public:
    template <typename T> FileStorage &operator<<(const T &) { return *this; }
};
typedef std::string String;
struct Range {
    int start, end;
    Range() : start(0), end(0) {}
    Range(int s, int e) : start(s), end(e) {}
};
struct KeyPoint {
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
};
struct DMatch {
    int queryIdx, trainIdx, imgIdx;
    float distance;
    DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(FLT_MAX) {}
    DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(-1), distance(d) {}
    DMatch(int q, int t, int i, float d) : queryIdx(q), trainIdx(t), imgIdx(i), distance(d) {}
    bool operator<(const DMatch &m) const { return distance < m.distance; }
};
enum { NORM_HAMMING = 6, COLOR_BGR2GRAY = 6, THRESH_TOZERO = 3, CMP_LT = 3, INTER_LINEAR = 1 };

class Algorithm {
public:
    virtual ~Algorithm() {}
    virtual void read(const FileNode &) {}
    virtual void write(FileStorage &) const {}
};
template <typename T> using Ptr = std::shared_ptr<T>;
template <typename T, typename... A> Ptr<T> makePtr(A &&...a) { return std::make_shared<T>(std::forward<A>(a)...); }

/* opencv2/core/cvstd.hpp puts these names into namespace cv; the reference's sources, written inside namespace cv::line_descriptor, call
 * them unqualified, so `1 / sqrt(float)` in BinaryDescriptor::computeLBD is a float root and a float division (with only the C library's
 * ::sqrt(double) visible it would be computed in double and rounded once) */
using std::abs;
using std::exp;
using std::log;
using std::max;
using std::min;
using std::pow;
using std::sqrt;
using std::swap;

/* cv::LineIterator as the KeyLine fill uses it: only `count`, the number of pixels of the 8-connected line between the two points (rounded
 * to pixels, clipped to the image by moving an outside end point onto the border).  KeyLine::numOfPixels feeds the LBD descriptor only;
 * nothing on the cuboid path reads it and no test compares it. */
class LineIterator {
public:
    int count;
    LineIterator(const Mat &img, Point2f p1, Point2f p2)
    {
        auto clampi = [](long v, int n) { return (int)(v < 0 ? 0 : (v >= n ? n - 1 : v)); };
        const int x1 = clampi(std::lrint(p1.x), img.cols), y1 = clampi(std::lrint(p1.y), img.rows);
        const int x2 = clampi(std::lrint(p2.x), img.cols), y2 = clampi(std::lrint(p2.y), img.rows);
        count = std::max(std::abs(x2 - x1), std::abs(y2 - y1)) + 1;
    }
};

/* the three primitives of the LSD path: forwarded to the oracle's cv2-pinned restatements (see the header) */
inline float fastAtan2(float y, float x) { return lsd_orc_fast_atan2(y, x); }
inline void GaussianBlur(InputArray src_, OutputArray dst, Size ksize, double sigma)
{
    const Mat src = src_.getMat();
    if (src.type() == CV_8UC1 && ksize.width == 5 && ksize.height == 5 && std::fabs(sigma - 1.0) < 1e-6) {
        /* BinaryDescriptor::OctaveKeyLines, octave 0 (binary_descriptor.cpp:813-814): OpenCV's fixed-point 8-bit smoothing */
        Mat out(src.rows, src.cols, CV_8UC1);
        edl_orc_gaussian5_u8(src.data, src.cols, src.rows, out.data);
        dst.assign(out);
        return;
    }
    if (src.type() != CV_64FC1 || ksize.width != 7 || ksize.height != 7 || std::fabs(sigma - 0.6 / 0.8) > 1e-12)
        throw std::runtime_error("minicv: GaussianBlur is only provided for the calls lsd.cpp (CV_64F, 7 x 7, sigma 0.6 / 0.8) and binary_descriptor.cpp (CV_8U, 5 x 5, sigma 1) make");
    Mat out(src.rows, src.cols, CV_64FC1);
    lsd_orc_gaussian7(src.ptr<double>(), src.cols, src.rows, out.ptr<double>());
    dst.assign(out);
}
inline void resize(InputArray src_, OutputArray dst, Size dsize, double fx, double fy)
{
    const Mat src = src_.getMat();
    if (src.type() == CV_8UC1 && dsize.area() == 0) {
        /* OctaveKeyLines shrinks the image for the NEXT octave after every octave, also after the last one (binary_descriptor.cpp:838):
         * with numOfOctave_ = 1 the result is never read; with more octaves it feeds the detection of the higher ones, whose lines
         * filter_lines drops again (tests/test_oracle_ref_octaves.py).  Sizes as cv::resize computes them.  Halving an image of even size
         * is the exact 2 x 2 mean with rounding, (a + b + c + d + 2) >> 2, as cv2 computes INTER_LINEAR there (checked against cv2 4.13);
         * any other case is plain bilinear interpolation at pixel centres, not pinned. */
        const int dw = (int)std::lrint(src.cols * fx), dh = (int)std::lrint(src.rows * fy);
        Mat out(dh, dw, CV_8UC1);
        if (fx == 0.5 && fy == 0.5 && src.cols % 2 == 0 && src.rows % 2 == 0) {
            for (int y = 0; y < dh; y++)
                for (int x = 0; x < dw; x++) {
                    const uchar *p = src.data + (size_t)(2 * y) * src.cols + 2 * x;
                    out.data[(size_t)y * dw + x] = (uchar)((p[0] + p[1] + p[src.cols] + p[src.cols + 1] + 2) >> 2);
                }
        } else {
            for (int y = 0; y < dh; y++)
                for (int x = 0; x < dw; x++) {
                    const double sx = std::max(0.0, (x + 0.5) / fx - 0.5), sy = std::max(0.0, (y + 0.5) / fy - 0.5);
                    const int x0 = std::min((int)sx, src.cols - 1), y0 = std::min((int)sy, src.rows - 1);
                    const int x1 = std::min(x0 + 1, src.cols - 1), y1 = std::min(y0 + 1, src.rows - 1);
                    const double ax = sx - x0, ay = sy - y0;
                    const double v = (1 - ay) * ((1 - ax) * src.data[(size_t)y0 * src.cols + x0] + ax * src.data[(size_t)y0 * src.cols + x1]) +
                                     ay * ((1 - ax) * src.data[(size_t)y1 * src.cols + x0] + ax * src.data[(size_t)y1 * src.cols + x1]);
                    out.data[(size_t)y * dw + x] = (uchar)std::lrint(v);
                }
        }
        dst.assign(out);
        return;
    }
    if (src.type() != CV_64FC1 || dsize.area() != 0 || fx != fy) throw std::runtime_error("minicv: resize is only provided for the call lsd.cpp makes");
    /* cv::resize: dsize = Size(saturate_cast<int>(cols * fx), saturate_cast<int>(rows * fy)), saturate_cast<int>(double) = lrint */
    const int dw = (int)std::lrint(src.cols * fx), dh = (int)std::lrint(src.rows * fy);
    Mat out(dh, dw, CV_64FC1);
    int ow = 0, oh = 0;
    lsd_orc_resize(src.ptr<double>(), src.cols, src.rows, fx, out.ptr<double>(), &ow, &oh);
    if (ow != dw || oh != dh) throw std::runtime_error("minicv: resize size mismatch");
    dst.assign(out);
}

/* ---- what EDLineDetector::EdgeDrawing and the line fit call (binary_descriptor.cpp:1617-1629, 2652-2655, 2743-2749), on the types they call
 * it with; integer semantics as OpenCV defines them (saturate_cast, cvRound = round half to even, BORDER_REFLECT_101) */
inline int reflect101_(int p, int n)
{
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = (p < 0) ? -p : 2 * n - 2 - p;
    return p;
}
inline short sat16_(int v) { return (short)(v < -32768 ? -32768 : (v > 32767 ? 32767 : v)); }
inline void Sobel(InputArray src_, OutputArray dst, int ddepth, int dx, int dy, int ksize)
{
    const Mat src = src_.getMat();
    if (src.type() != CV_8UC1 || (ddepth & 7) != CV_16S || ksize != 3 || dx + dy != 1) throw std::runtime_error("minicv: Sobel is only provided for 8-bit -> 16-bit, 3 x 3, first order");
    const int w = src.cols, h = src.rows;
    Mat out(h, w, CV_16SC1);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            auto px = [&](int yy, int xx) { return (int)src.data[(size_t)reflect101_(yy, h) * w + reflect101_(xx, w)]; };
            int v;
            if (dx == 1)
                v = (px(y - 1, x + 1) - px(y - 1, x - 1)) + 2 * (px(y, x + 1) - px(y, x - 1)) + (px(y + 1, x + 1) - px(y + 1, x - 1));
            else
                v = (px(y + 1, x - 1) - px(y - 1, x - 1)) + 2 * (px(y + 1, x) - px(y - 1, x)) + (px(y + 1, x + 1) - px(y - 1, x + 1));
            out.ptr<short>()[(size_t)y * w + x] = sat16_(v);
        }
    dst.assign(out);
}
inline Mat abs(const Mat &a)
{
    if (a.type() != CV_16SC1) throw std::runtime_error("minicv: abs is only provided for 16-bit");
    Mat o(a.rows, a.cols, CV_16SC1);
    for (size_t i = 0; i < a.total(); i++) o.ptr<short>()[i] = sat16_(std::abs((int)a.ptr<short>()[i]));
    return o;
}
inline void add(InputArray a_, InputArray b_, OutputArray dst)
{
    const Mat a = a_.getMat(), b = b_.getMat();
    if (a.type() != CV_16SC1 || b.type() != CV_16SC1 || a.size() != b.size()) throw std::runtime_error("minicv: add is only provided for 16-bit");
    Mat o(a.rows, a.cols, CV_16SC1);
    for (size_t i = 0; i < a.total(); i++) o.ptr<short>()[i] = sat16_((int)a.ptr<short>()[i] + (int)b.ptr<short>()[i]);
    dst.assign(o);
}
inline double threshold(InputArray src_, OutputArray dst, double thresh, double, int type)
{
    const Mat a = src_.getMat();
    if (a.type() != CV_16SC1 || type != THRESH_TOZERO) throw std::runtime_error("minicv: threshold is only provided for 16-bit THRESH_TOZERO");
    const int ith = (int)std::floor(thresh); /* the integer types compare against cvFloor(thresh) */
    Mat o(a.rows, a.cols, CV_16SC1);
    for (size_t i = 0; i < a.total(); i++) o.ptr<short>()[i] = a.ptr<short>()[i] > ith ? a.ptr<short>()[i] : (short)0;
    dst.assign(o);
    return thresh;
}
inline void compare(InputArray a_, InputArray b_, OutputArray dst, int op)
{
    const Mat a = a_.getMat(), b = b_.getMat();
    if (a.type() != CV_16SC1 || b.type() != CV_16SC1 || op != CMP_LT) throw std::runtime_error("minicv: compare is only provided for 16-bit CMP_LT");
    Mat o(a.rows, a.cols, CV_8UC1);
    for (size_t i = 0; i < a.total(); i++) o.data[i] = a.ptr<short>()[i] < b.ptr<short>()[i] ? 255 : 0;
    dst.assign(o);
}
/* Mat / s on 16-bit: a scaled conversion, saturate_cast<short>(v * (1 / s)) with round-half-to-even */
inline Mat operator/(const Mat &a, int s)
{
    if (a.type() != CV_16SC1) throw std::runtime_error("minicv: Mat / int is only provided for 16-bit");
    Mat o(a.rows, a.cols, CV_16SC1);
    const double alpha = 1.0 / s;
    for (size_t i = 0; i < a.total(); i++) o.ptr<short>()[i] = sat16_((int)std::nearbyint(a.ptr<short>()[i] * alpha));
    return o;
}
/* float matrix product: double accumulators, rounded to float once (cv::gemm's float path) */
inline Mat operator*(const Mat &a, const Mat &b)
{
    if (a.type() != CV_32FC1 || b.type() != CV_32FC1 || a.cols != b.rows) throw std::runtime_error("minicv: Mat * Mat is only provided for float matrices");
    Mat o(a.rows, b.cols, CV_32FC1);
    for (int r = 0; r < a.rows; r++)
        for (int c = 0; c < b.cols; c++) {
            double acc = 0;
            for (int k = 0; k < a.cols; k++) acc += (double)a.ptr<float>()[(size_t)r * a.cols + k] * (double)b.ptr<float>()[(size_t)k * b.cols + c];
            o.ptr<float>()[(size_t)r * b.cols + c] = (float)acc;
        }
    return o;
}
inline Mat operator+(const Mat &a, const Mat &b)
{
    if (a.type() != CV_32FC1 || b.type() != CV_32FC1 || a.size() != b.size()) throw std::runtime_error("minicv: Mat + Mat is only provided for float matrices");
    Mat o(a.rows, a.cols, CV_32FC1);
    for (size_t i = 0; i < a.total(); i++) o.ptr<float>()[i] = a.ptr<float>()[i] + b.ptr<float>()[i];
    return o;
}

/* ---- detect_3d_cuboid::detect_cuboid (box_proposal_detail.cpp:196-199): Canny on the ROI, 255 - edges, L2 3 x 3 distance transform: the
 * oracle's restatements, pinned bit for bit against cv2 (tests/test_oracle_cv_parity.py, tests/golden/cv_pins.npz) */
inline void Canny(InputArray src_, OutputArray dst, double low, double high)
{
    const Mat src = src_.getMat();
    if (src.type() != CV_8UC1) throw std::runtime_error("minicv: Canny is provided for 8-bit single channel");
    Mat out(src.rows, src.cols, CV_8UC1);
    orc_canny(src.data, src.cols, src.rows, src.cols, low, high, out.data);
    dst.assign(out);
}
inline Mat operator-(int s, const Mat &a)
{
    if (a.type() != CV_8UC1) throw std::runtime_error("minicv: scalar - Mat is provided for 8-bit");
    Mat o(a.rows, a.cols, CV_8UC1);
    for (size_t i = 0; i < a.total(); i++) {
        const int v = s - (int)a.data[i];
        o.data[i] = (uchar)(v < 0 ? 0 : (v > 255 ? 255 : v));
    }
    return o;
}
inline void distanceTransform(InputArray src_, OutputArray dst, int distanceType, int maskSize)
{
    const Mat src = src_.getMat();
    if (src.type() != CV_8UC1 || distanceType != CV_DIST_L2 || maskSize != 3) throw std::runtime_error("minicv: distanceTransform is provided for 8-bit, DIST_L2, 3 x 3");
    /* distance to the nearest ZERO pixel; the oracle's routine takes the edge map (non-zero = distance 0) */
    Mat edges(src.rows, src.cols, CV_8UC1);
    for (size_t i = 0; i < src.total(); i++) edges.data[i] = src.data[i] == 0 ? 255 : 0;
    Mat out(src.rows, src.cols, CV_32FC1);
    orc_chamfer_dt(edges.data, src.cols, src.rows, out.ptr<float>());
    dst.assign(out);
}
enum { NORM_MINMAX = 32 };
template <typename... A> inline void imshow(A &&...) {}
inline int waitKey(int = 0) { return -1; }
template <typename... A> inline void normalize(A &&...) {}

/* drawing helpers referenced by LineSegmentDetectorImpl::drawSegments / compareSegments, which the path never calls */
[[noreturn]] inline void minicv_unreachable(const char *what)
{
    fprintf(stderr, "minicv: %s is not provided (the line-detection path does not call it)\n", what);
    abort();
}
inline void cvtColor(InputArray src_, OutputArray dst, int code)
{
    /* LSDDetector::detectImpl / BinaryDescriptor::detectImpl convert a colour input themselves (LSDDetector.cpp:165-168,
     * binary_descriptor.cpp:487-490): OpenCV's 15-bit fixed-point BGR -> gray, the oracle's cv2-pinned restatement */
    const Mat src = src_.getMat();
    if (code != COLOR_BGR2GRAY || src.type() != CV_8UC3) minicv_unreachable("cvtColor other than 8-bit BGR2GRAY");
    Mat out(src.rows, src.cols, CV_8UC1);
    orc_bgr2gray(src.data, src.cols, src.rows, src.cols * 3, out.data, src.cols, 1);
    dst.assign(out);
}
inline void merge(const std::vector<Mat> &, OutputArray) { minicv_unreachable("merge"); }
template <typename P, typename... A> inline void line(InputOutputArray, P, P, const Scalar &, A...) { minicv_unreachable("line"); }
inline void bitwise_xor(InputArray, InputArray, OutputArray) { minicv_unreachable("bitwise_xor"); }
inline int countNonZero(InputArray) { minicv_unreachable("countNonZero"); }
/* cv::pyrDown on an 8-bit image (LSDDetector::computeGaussianPyramid, LSDDetector.cpp:66-72, and BinaryDescriptor::computeGaussianPyramid
 * for octaves above 0): the separable 5-tap kernel (1 4 6 4 1) / 16 in both directions at the even source positions, BORDER_REFLECT_101,
 * one rounding (sum + 128) >> 8.  Equal to cv2.pyrDown bit for bit (tests/test_oracle_ref_octaves.py). */
inline void pyrDown(InputArray src_, OutputArray dst, Size dsize)
{
    const Mat src = src_.getMat();
    if (src.type() != CV_8UC1) minicv_unreachable("pyrDown other than 8-bit single channel");
    const int w = src.cols, h = src.rows, dw = dsize.width, dh = dsize.height;
    if (dw <= 0 || dh <= 0 || std::abs(dw * 2 - w) > 2 || std::abs(dh * 2 - h) > 2) throw std::runtime_error("minicv: pyrDown destination size");
    auto r101 = [](int p, int n) {
        if (n == 1) return 0;
        while (p < 0 || p >= n) p = (p < 0) ? -p : 2 * n - 2 - p;
        return p;
    };
    static const int k[5] = {1, 4, 6, 4, 1};
    std::vector<int> t((size_t)h * dw);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < dw; x++) {
            int a = 0;
            for (int i = 0; i < 5; i++) a += k[i] * src.data[(size_t)y * w + r101(2 * x + i - 2, w)];
            t[(size_t)y * dw + x] = a;
        }
    Mat out(dh, dw, CV_8UC1);
    for (int y = 0; y < dh; y++)
        for (int x = 0; x < dw; x++) {
            int a = 0;
            for (int i = 0; i < 5; i++) a += k[i] * t[(size_t)r101(2 * y + i - 2, h) * dw + x];
            out.data[(size_t)y * dw + x] = (uchar)((a + 128) >> 8);
        }
    dst.assign(out);
}

}  // namespace cv
#endif /* ORC_MINICV_HPP */
