/*
 * oracle/ref/minieigen.hpp -- CPU ORACLE, TEST INFRASTRUCTURE ONLY.
 *
 * A small stand-in for the part of the Eigen API that the reference's detect_3d_cuboid sources use (this image has no Eigen), so that
 * box_proposal_detail.cpp, object_3d_util.cpp and matrix_utils.cpp compile UNMODIFIED from where they lie under /root/reference
 * (oracle/ref/cuboid_ref.cpp).  Dense matrices with value semantics, everything evaluated eagerly (no expression templates), views
 * for row / col / block / head / tail / corners that write through, `.array()` for coefficient-wise arithmetic, the comma initialiser,
 * Quaternion.  Arithmetic follows what Eigen documents for small matrices: products are coefficient-wise sums in index order, 3 x 3 and
 * 4 x 4 inverses by cofactors, Quaternion(Matrix3) by the trace / largest-diagonal branches.  Where Eigen's vectorised kernels would
 * associate a sum differently the last bits of CONTINUOUS outputs may differ from a build against the real library; every DISCRETE decision
 * of the reference (which proposals are valid, which ids are kept, which cuboid wins) is taken by the reference's own code.
 */
#ifndef ORC_MINIEIGEN_HPP
#define ORC_MINIEIGEN_HPP

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <iostream>
#include <numeric>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

namespace Eigen {

const int Dynamic = -1;
enum NoChange_t { NoChange };
typedef std::ptrdiff_t Index;

template <typename T> class Dense;
template <typename T> class Block;
template <typename T> class ArrayX;
template <typename T, int R, int C> class Matrix;

[[noreturn]] inline void me_fail(const char *what) { throw std::runtime_error(std::string("minieigen: ") + what); }

/* ---- read-only operations shared by matrices and views: Derived provides eval() -> Dense<T> ---- */
template <typename D, typename T>
struct ReadOps {
    const D &self() const { return *static_cast<const D *>(this); }
    Dense<T> ev() const { return self().eval(); }
    T norm() const { return ev().norm_(); }
    T squaredNorm() const { return ev().sqnorm_(); }
    T sum() const { return ev().sum_(); }
    T mean() const { return ev().mean_(); }
    T minCoeff() const { int i; return ev().min_(&i); }
    T maxCoeff() const { int i; return ev().max_(&i); }
    template <typename I> T minCoeff(I *idx) const { int i; T v = ev().min_(&i); *idx = (I)i; return v; }
    template <typename I> T maxCoeff(I *idx) const { int i; T v = ev().max_(&i); *idx = (I)i; return v; }
    Dense<T> transpose() const { return ev().transpose_(); }
    Dense<T> inverse() const { return ev().inverse_(); }
    Dense<T> normalized() const { Dense<T> e = ev(); T n = e.norm_(); return e / n; }
    ArrayX<T> array() const;
    template <typename U> Dense<U> cast() const { return ev().template cast_<U>(); }
    Dense<T> cross(const Dense<T> &o) const { return ev().cross_(o); }
    T dot(const Dense<T> &o) const { return ev().dot_(o); }
    Dense<T> asDiagonal() const { return ev().asDiagonal_(); }
    template <int RF, int CF> Dense<T> replicate() const { return ev().replicate_(RF, CF); }
    Dense<T> replicate(int rf, int cf) const { return ev().replicate_(rf, cf); }
    Dense<T> cwiseProduct(const Dense<T> &o) const { return ev().cwise_(o, 0); }
    Dense<T> cwiseQuotient(const Dense<T> &o) const { return ev().cwise_(o, 1); }
    Dense<T> cwiseAbs() const { return ev().abs_(); }
    struct RowwiseProxy {
        Dense<T> m;
        Dense<T> norm() const;
        Dense<T> sum() const;
    };
    RowwiseProxy rowwise() const { return RowwiseProxy{ev()}; }
};

/* ---- dense storage, column-major like Eigen ---- */
template <typename T>
class Dense : public ReadOps<Dense<T>, T> {
public:
    typedef T Scalar;
    Dense() : r_(0), c_(0) {}
    Dense(int r, int c) : r_(r), c_(c), d_((size_t)r * c, T(0)) {}
    Dense(const Block<T> &b);
    Dense(const ArrayX<T> &a);
    const Dense &eval() const { return *this; }
    int rows() const { return r_; }
    int cols() const { return c_; }
    int size() const { return r_ * c_; }
    T *data() { return d_.data(); }
    const T *data() const { return d_.data(); }
    T &operator()(int i, int j) { return d_[(size_t)j * r_ + i]; }
    const T &operator()(int i, int j) const { return d_[(size_t)j * r_ + i]; }
    T &operator()(int i) { return d_[i]; }
    const T &operator()(int i) const { return d_[i]; }
    T &operator[](int i) { return d_[i]; }
    const T &operator[](int i) const { return d_[i]; }
    T &x() { return d_[0]; }
    T &y() { return d_[1]; }
    T &z() { return d_[2]; }
    T &w() { return d_[3]; }
    const T &x() const { return d_[0]; }
    const T &y() const { return d_[1]; }
    const T &z() const { return d_[2]; }
    const T &w() const { return d_[3]; }
    void resize(int r, int c)
    {
        if (r == r_ && c == c_) return;
        r_ = r;
        c_ = c;
        d_.assign((size_t)r * c, T(0));
    }
    void resize(int n) /* vectors */
    {
        if (c_ == 1 || (r_ == 0 && c_ == 0 && !rowvec_))
            resize(n, 1);
        else
            resize(1, n);
    }
    void resize(int r, NoChange_t) { resize(r, c_); }
    void resize(NoChange_t, int c) { resize(r_, c); }
    void conservativeResize(int r, int c)
    {
        Dense o(r, c);
        for (int j = 0; j < std::min(c, c_); j++)
            for (int i = 0; i < std::min(r, r_); i++) o(i, j) = (*this)(i, j);
        const bool rv = rowvec_;
        *this = o;
        rowvec_ = rv;
    }
    void conservativeResize(int r, NoChange_t) { conservativeResize(r, c_); }
    void conservativeResize(NoChange_t, int c) { conservativeResize(r_, c); }
    void conservativeResize(int n) { c_ == 1 || !rowvec_ ? conservativeResize(n, 1) : conservativeResize(1, n); }
    Dense &setZero() { std::fill(d_.begin(), d_.end(), T(0)); return *this; }
    Dense &setZero(int r, int c) { resize(r, c); return setZero(); }
    Dense &setOnes() { std::fill(d_.begin(), d_.end(), T(1)); return *this; }
    Dense &setConstant(T v) { std::fill(d_.begin(), d_.end(), v); return *this; }
    Dense &fill(T v) { return setConstant(v); }
    Dense &setIdentity()
    {
        setZero();
        for (int i = 0; i < std::min(r_, c_); i++) (*this)(i, i) = T(1);
        return *this;
    }
    /* views */
    Block<T> block(int i, int j, int r, int c);
    Dense block(int i, int j, int r, int c) const { return sub_(i, j, r, c); }
    template <int BR, int BC> Block<T> block(int i, int j);
    template <int BR, int BC> Dense block(int i, int j) const { return sub_(i, j, BR, BC); }
    Block<T> row(int i);
    Dense row(int i) const { return sub_(i, 0, 1, c_); }
    Block<T> col(int j);
    Dense col(int j) const { return sub_(0, j, r_, 1); }
    Block<T> head(int n);
    Dense head(int n) const { return vec_sub_(0, n); }
    template <int N> Block<T> head();
    template <int N> Dense head() const { return vec_sub_(0, N); }
    Block<T> tail(int n);
    Dense tail(int n) const { return vec_sub_(size() - n, n); }
    template <int N> Block<T> tail();
    template <int N> Dense tail() const { return vec_sub_(size() - N, N); }
    Block<T> segment(int s, int n);
    Dense segment(int s, int n) const { return vec_sub_(s, n); }
    template <int N> Block<T> segment(int s);
    template <int N> Dense segment(int s) const { return vec_sub_(s, N); }
    /* a 1 x 1 result used as a number (v.transpose() * w) */
    operator T() const
    {
        if (r_ != 1 || c_ != 1) me_fail("a matrix that is not 1 x 1 used as a scalar");
        return d_[0];
    }
    Block<T> topRows(int n);
    Dense topRows(int n) const { return sub_(0, 0, n, c_); }
    template <int N> Block<T> topRows();
    template <int N> Dense topRows() const { return sub_(0, 0, N, c_); }
    Block<T> bottomRows(int n);
    Dense bottomRows(int n) const { return sub_(r_ - n, 0, n, c_); }
    Block<T> leftCols(int n);
    Dense leftCols(int n) const { return sub_(0, 0, r_, n); }
    Block<T> rightCols(int n);
    Dense rightCols(int n) const { return sub_(0, c_ - n, r_, n); }
    Block<T> topLeftCorner(int r, int c);
    Dense topLeftCorner(int r, int c) const { return sub_(0, 0, r, c); }
    template <int BR, int BC> Block<T> topLeftCorner();
    template <int BR, int BC> Dense topLeftCorner() const { return sub_(0, 0, BR, BC); }
    Block<T> topRightCorner(int r, int c);
    Dense topRightCorner(int r, int c) const { return sub_(0, c_ - c, r, c); }
    template <int BR, int BC> Block<T> topRightCorner();
    template <int BR, int BC> Dense topRightCorner() const { return sub_(0, c_ - BC, BR, BC); }
    Block<T> bottomLeftCorner(int r, int c);
    Dense bottomLeftCorner(int r, int c) const { return sub_(r_ - r, 0, r, c); }
    Block<T> bottomRightCorner(int r, int c);
    Dense bottomRightCorner(int r, int c) const { return sub_(r_ - r, c_ - c, r, c); }
    ArrayX<T> array();       /* writable: m.array() -= 1 */
    ArrayX<T> array() const;
    /* arithmetic */
    Dense &operator+=(const Dense &o) { chk_(o); for (size_t k = 0; k < d_.size(); k++) d_[k] += o.d_[k]; return *this; }
    Dense &operator-=(const Dense &o) { chk_(o); for (size_t k = 0; k < d_.size(); k++) d_[k] -= o.d_[k]; return *this; }
    Dense &operator*=(T s) { for (auto &v : d_) v *= s; return *this; }
    Dense &operator/=(T s) { for (auto &v : d_) v /= s; return *this; }
    Dense operator-() const { Dense o(*this); for (auto &v : o.d_) v = -v; return o; }
    /* helpers used by ReadOps */
    T sqnorm_() const { T s = T(0); for (const auto &v : d_) s += v * v; return s; }
    T norm_() const { return std::sqrt(sqnorm_()); }
    T sum_() const { T s = T(0); for (const auto &v : d_) s += v; return s; }
    T mean_() const { return sum_() / T(d_.size()); }
    T min_(int *idx) const { int b = 0; for (int k = 1; k < size(); k++) if (d_[k] < d_[b]) b = k; *idx = b; return d_[b]; }
    T max_(int *idx) const { int b = 0; for (int k = 1; k < size(); k++) if (d_[k] > d_[b]) b = k; *idx = b; return d_[b]; }
    Dense transpose_() const
    {
        Dense o(c_, r_);
        for (int j = 0; j < c_; j++)
            for (int i = 0; i < r_; i++) o(j, i) = (*this)(i, j);
        return o;
    }
    Dense inverse_() const;
    template <typename U> Dense<U> cast_() const
    {
        Dense<U> o(r_, c_);
        for (int k = 0; k < size(); k++) o(k) = (U)d_[k];
        return o;
    }
    Dense cross_(const Dense &o) const
    {
        if (size() != 3 || o.size() != 3) me_fail("cross of non-3-vectors");
        Dense v(3, 1);
        v(0) = d_[1] * o(2) - d_[2] * o(1);
        v(1) = d_[2] * o(0) - d_[0] * o(2);
        v(2) = d_[0] * o(1) - d_[1] * o(0);
        return v;
    }
    T dot_(const Dense &o) const { T s = T(0); for (int k = 0; k < size(); k++) s += d_[k] * o(k); return s; }
    Dense asDiagonal_() const
    {
        Dense o(size(), size());
        for (int k = 0; k < size(); k++) o(k, k) = d_[k];
        return o;
    }
    Dense replicate_(int rf, int cf) const
    {
        Dense o(r_ * rf, c_ * cf);
        for (int j = 0; j < o.cols(); j++)
            for (int i = 0; i < o.rows(); i++) o(i, j) = (*this)(i % r_, j % c_);
        return o;
    }
    Dense cwise_(const Dense &o, int op) const
    {
        chk_(o);
        Dense r(r_, c_);
        for (int k = 0; k < size(); k++) r(k) = op == 0 ? d_[k] * o(k) : d_[k] / o(k);
        return r;
    }
    Dense abs_() const { Dense o(*this); for (auto &v : o.d_) v = v < T(0) ? -v : v; return o; }
    void mark_rowvec_() { rowvec_ = true; }
    bool same_shape_or_vec_(const Dense &o) const { return (r_ == o.r_ && c_ == o.c_) || ((r_ == 1 || c_ == 1) && (o.r_ == 1 || o.c_ == 1) && size() == o.size()); }

protected:
    Dense sub_(int i, int j, int r, int c) const
    {
        if (i < 0 || j < 0 || i + r > r_ || j + c > c_) me_fail("block out of range");
        Dense o(r, c);
        for (int b = 0; b < c; b++)
            for (int a = 0; a < r; a++) o(a, b) = (*this)(i + a, j + b);
        return o;
    }
    Dense vec_sub_(int s, int n) const { return c_ == 1 ? sub_(s, 0, n, 1) : (r_ == 1 ? sub_(0, s, 1, n) : (me_fail("head / tail / segment of a matrix"), Dense())); }
    void chk_(const Dense &o) const { if (!same_shape_or_vec_(o)) me_fail("shape mismatch"); }
    int r_, c_;
    std::vector<T> d_;
    bool rowvec_ = false;
};

/* ---- a writable view of a rectangular part of a Dense ---- */
template <typename T>
class Block : public ReadOps<Block<T>, T> {
public:
    Block(Dense<T> *p, int i, int j, int r, int c) : p_(p), i_(i), j_(j), r_(r), c_(c)
    {
        if (p && (i < 0 || j < 0 || i + r > p->rows() || j + c > p->cols())) me_fail("block out of range");
    }
    Dense<T> eval() const
    {
        Dense<T> o(r_, c_);
        for (int b = 0; b < c_; b++)
            for (int a = 0; a < r_; a++) o(a, b) = (*p_)(i_ + a, j_ + b);
        return o;
    }
    int rows() const { return r_; }
    int cols() const { return c_; }
    int size() const { return r_ * c_; }
    T &operator()(int a, int b) { return (*p_)(i_ + a, j_ + b); }
    const T &operator()(int a, int b) const { return (*p_)(i_ + a, j_ + b); }
    T &operator()(int k) { return c_ == 1 ? (*p_)(i_ + k, j_) : (*p_)(i_, j_ + k); }
    const T &operator()(int k) const { return c_ == 1 ? (*p_)(i_ + k, j_) : (*p_)(i_, j_ + k); }
    T &operator[](int k) { return (*this)(k); }
    void assign(const Dense<T> &m)
    {
        if (m.rows() == r_ && m.cols() == c_) {
            for (int b = 0; b < c_; b++)
                for (int a = 0; a < r_; a++) (*this)(a, b) = m(a, b);
        } else if ((r_ == 1 || c_ == 1) && (m.rows() == 1 || m.cols() == 1) && m.size() == size()) {
            for (int k = 0; k < size(); k++) (*this)(k) = m(k);
        } else
            me_fail("assignment to a block of another shape");
    }
    Block &operator=(const Dense<T> &m) { assign(m); return *this; }
    Block &operator=(const Block &o) { assign(o.eval()); return *this; }
    Block &operator=(const ArrayX<T> &a);
    Block &operator+=(const Dense<T> &m) { Dense<T> e = eval(); e += (e.rows() == m.rows() ? m : m.transpose_()); assign(e); return *this; }
    Block &operator-=(const Dense<T> &m) { Dense<T> e = eval(); e -= (e.rows() == m.rows() ? m : m.transpose_()); assign(e); return *this; }
    Block &operator*=(T s) { Dense<T> e = eval(); e *= s; assign(e); return *this; }
    Block &operator/=(T s) { Dense<T> e = eval(); e /= s; assign(e); return *this; }
    Block &setZero() { Dense<T> e(r_, c_); assign(e); return *this; }
    /* views of views (vectors) */
    Block head(int n) { return c_ == 1 ? Block(p_, i_, j_, n, 1) : Block(p_, i_, j_, 1, n); }
    template <int N> Block head() { return head(N); }
    Block tail(int n) { return c_ == 1 ? Block(p_, i_ + r_ - n, j_, n, 1) : Block(p_, i_, j_ + c_ - n, 1, n); }
    template <int N> Block tail() { return tail(N); }
    Block segment(int s, int n) { return c_ == 1 ? Block(p_, i_ + s, j_, n, 1) : Block(p_, i_, j_ + s, 1, n); }
    template <int N> Block segment(int s) { return segment(s, N); }
    Block row(int a) { return Block(p_, i_ + a, j_, 1, c_); }
    Block col(int b) { return Block(p_, i_, j_ + b, r_, 1); }
    Block block(int a, int b, int r, int c) { return Block(p_, i_ + a, j_ + b, r, c); }
    ArrayX<T> array(); /* writable */
    ArrayX<T> array() const;

private:
    Dense<T> *p_;
    int i_, j_, r_, c_;
};

template <typename T> Dense<T>::Dense(const Block<T> &b) : r_(0), c_(0) { *this = b.eval(); }
template <typename T> Block<T> Dense<T>::block(int i, int j, int r, int c) { return Block<T>(this, i, j, r, c); }
template <typename T> template <int BR, int BC> Block<T> Dense<T>::block(int i, int j) { return Block<T>(this, i, j, BR, BC); }
template <typename T> Block<T> Dense<T>::row(int i) { return Block<T>(this, i, 0, 1, c_); }
template <typename T> Block<T> Dense<T>::col(int j) { return Block<T>(this, 0, j, r_, 1); }
template <typename T> Block<T> Dense<T>::head(int n) { return c_ == 1 ? Block<T>(this, 0, 0, n, 1) : Block<T>(this, 0, 0, 1, n); }
template <typename T> template <int N> Block<T> Dense<T>::head() { return head(N); }
template <typename T> Block<T> Dense<T>::tail(int n) { return c_ == 1 ? Block<T>(this, r_ - n, 0, n, 1) : Block<T>(this, 0, c_ - n, 1, n); }
template <typename T> template <int N> Block<T> Dense<T>::tail() { return tail(N); }
template <typename T> template <int N> Block<T> Dense<T>::segment(int s) { return segment(s, N); }
template <typename T> Block<T> Dense<T>::segment(int s, int n) { return c_ == 1 ? Block<T>(this, s, 0, n, 1) : Block<T>(this, 0, s, 1, n); }
template <typename T> Block<T> Dense<T>::topRows(int n) { return Block<T>(this, 0, 0, n, c_); }
template <typename T> template <int N> Block<T> Dense<T>::topRows() { return topRows(N); }
template <typename T> Block<T> Dense<T>::bottomRows(int n) { return Block<T>(this, r_ - n, 0, n, c_); }
template <typename T> Block<T> Dense<T>::leftCols(int n) { return Block<T>(this, 0, 0, r_, n); }
template <typename T> Block<T> Dense<T>::rightCols(int n) { return Block<T>(this, 0, c_ - n, r_, n); }
template <typename T> Block<T> Dense<T>::topLeftCorner(int r, int c) { return Block<T>(this, 0, 0, r, c); }
template <typename T> template <int BR, int BC> Block<T> Dense<T>::topLeftCorner() { return topLeftCorner(BR, BC); }
template <typename T> Block<T> Dense<T>::topRightCorner(int r, int c) { return Block<T>(this, 0, c_ - c, r, c); }
template <typename T> template <int BR, int BC> Block<T> Dense<T>::topRightCorner() { return topRightCorner(BR, BC); }
template <typename T> Block<T> Dense<T>::bottomLeftCorner(int r, int c) { return Block<T>(this, r_ - r, 0, r, c); }
template <typename T> Block<T> Dense<T>::bottomRightCorner(int r, int c) { return Block<T>(this, r_ - r, c_ - c, r, c); }

/* ---- coefficient-wise values: what .array() gives.  Holds a copy; remembers where it came from for the in-place operators ---- */
template <typename T>
class ArrayX {
public:
    ArrayX(const Dense<T> &m) : v(m), dst_dense(nullptr), has_block(false), dst_block(nullptr, 0, 0, 0, 0) {}
    ArrayX(const Dense<T> &m, Dense<T> *d) : v(m), dst_dense(d), has_block(false), dst_block(d, 0, 0, 0, 0) {}
    ArrayX(const Dense<T> &m, const Block<T> &b) : v(m), dst_dense(nullptr), has_block(true), dst_block(b) {}
    Dense<T> v;
    Dense<T> matrix() const { return v; }
    ArrayX abs() const { return ArrayX(v.abs_()); }
    ArrayX square() const { return ArrayX(v.cwise_(v, 0)); }
    ArrayX sqrt() const { Dense<T> o(v); for (int k = 0; k < o.size(); k++) o(k) = std::sqrt(o(k)); return ArrayX(o); }
    T sum() const { return v.sum_(); }
    T mean() const { return v.mean_(); }
    T minCoeff() const { int i; return v.min_(&i); }
    T maxCoeff() const { int i; return v.max_(&i); }
    int rows() const { return v.rows(); }
    int cols() const { return v.cols(); }
    int size() const { return v.size(); }
    const T &operator()(int k) const { return v(k); }
    void store_()
    {
        if (has_block)
            dst_block.assign(v);
        else if (dst_dense)
            *dst_dense = v;
        else
            me_fail("in-place operation on a temporary array");
    }
    ArrayX &operator-=(T s) { for (int k = 0; k < v.size(); k++) v(k) -= s; store_(); return *this; }
    ArrayX &operator+=(T s) { for (int k = 0; k < v.size(); k++) v(k) += s; store_(); return *this; }
    ArrayX &operator*=(T s) { for (int k = 0; k < v.size(); k++) v(k) *= s; store_(); return *this; }
    ArrayX &operator/=(T s) { for (int k = 0; k < v.size(); k++) v(k) /= s; store_(); return *this; }

private:
    Dense<T> *dst_dense;
    bool has_block;
    Block<T> dst_block;
};
struct BoolArray {
    std::vector<char> b;
    bool any() const { for (char c : b) if (c) return true; return false; }
    bool all() const { for (char c : b) if (!c) return false; return true; }
    int count() const { int n = 0; for (char c : b) n += c ? 1 : 0; return n; }
};
template <typename T> Dense<T>::Dense(const ArrayX<T> &a) : r_(0), c_(0) { *this = a.v; }
template <typename D, typename T> ArrayX<T> ReadOps<D, T>::array() const { return ArrayX<T>(ev()); }
template <typename T> ArrayX<T> Dense<T>::array() { return ArrayX<T>(*this, this); }
template <typename T> ArrayX<T> Dense<T>::array() const { return ArrayX<T>(*this); }
template <typename T> ArrayX<T> Block<T>::array() { return ArrayX<T>(eval(), *this); }
template <typename T> ArrayX<T> Block<T>::array() const { return ArrayX<T>(eval()); }
template <typename T> Block<T> &Block<T>::operator=(const ArrayX<T> &a) { assign(a.v); return *this; }

#define ME_ARR_BIN(op)                                                                                              \
    template <typename T> ArrayX<T> operator op(const ArrayX<T> &a, const ArrayX<T> &b)                             \
    {                                                                                                               \
        if (!a.v.same_shape_or_vec_(b.v)) me_fail("array shape mismatch");                                          \
        Dense<T> o(a.v);                                                                                            \
        for (int k = 0; k < o.size(); k++) o(k) = a.v(k) op b.v(k);                                                 \
        return ArrayX<T>(o);                                                                                        \
    }                                                                                                               \
    template <typename T, typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>       \
    ArrayX<T> operator op(const ArrayX<T> &a, S s)                                                                  \
    {                                                                                                               \
        Dense<T> o(a.v);                                                                                            \
        for (int k = 0; k < o.size(); k++) o(k) = a.v(k) op (T)s;                                                   \
        return ArrayX<T>(o);                                                                                        \
    }                                                                                                               \
    template <typename T, typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>       \
    ArrayX<T> operator op(S s, const ArrayX<T> &a)                                                                  \
    {                                                                                                               \
        Dense<T> o(a.v);                                                                                            \
        for (int k = 0; k < o.size(); k++) o(k) = (T)s op a.v(k);                                                   \
        return ArrayX<T>(o);                                                                                        \
    }
ME_ARR_BIN(+)
ME_ARR_BIN(-)
ME_ARR_BIN(*)
ME_ARR_BIN(/)
#undef ME_ARR_BIN
#define ME_ARR_CMP(op)                                                                                              \
    template <typename T, typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>       \
    BoolArray operator op(const ArrayX<T> &a, S s)                                                                  \
    {                                                                                                               \
        BoolArray r;                                                                                                \
        for (int k = 0; k < a.size(); k++) r.b.push_back(a.v(k) op (T)s);                                           \
        return r;                                                                                                   \
    }                                                                                                               \
    template <typename T> BoolArray operator op(const ArrayX<T> &a, const ArrayX<T> &b)                             \
    {                                                                                                               \
        BoolArray r;                                                                                                \
        for (int k = 0; k < a.size(); k++) r.b.push_back(a.v(k) op b.v(k));                                         \
        return r;                                                                                                   \
    }
ME_ARR_CMP(<)
ME_ARR_CMP(>)
ME_ARR_CMP(<=)
ME_ARR_CMP(>=)
ME_ARR_CMP(==)
#undef ME_ARR_CMP
template <typename T> ArrayX<T> operator-(const ArrayX<T> &a) { return ArrayX<T>(-a.v); }

template <typename D, typename T> Dense<T> ReadOps<D, T>::RowwiseProxy::norm() const
{
    Dense<T> o(m.rows(), 1);
    for (int i = 0; i < m.rows(); i++) {
        T s = T(0);
        for (int j = 0; j < m.cols(); j++) s += m(i, j) * m(i, j);
        o(i) = std::sqrt(s);
    }
    return o;
}
template <typename D, typename T> Dense<T> ReadOps<D, T>::RowwiseProxy::sum() const
{
    Dense<T> o(m.rows(), 1);
    for (int i = 0; i < m.rows(); i++) {
        T s = T(0);
        for (int j = 0; j < m.cols(); j++) s += m(i, j);
        o(i) = s;
    }
    return o;
}

/* ---- matrix arithmetic ---- */
template <typename T> Dense<T> operator+(const Dense<T> &a, const Dense<T> &b) { Dense<T> o(a); o += (a.rows() == b.rows() && a.cols() == b.cols()) ? b : b.transpose_(); return o; }
template <typename T> Dense<T> operator-(const Dense<T> &a, const Dense<T> &b) { Dense<T> o(a); o -= (a.rows() == b.rows() && a.cols() == b.cols()) ? b : b.transpose_(); return o; }
template <typename T> Dense<T> operator*(const Dense<T> &a, const Dense<T> &b)
{
    if (a.cols() != b.rows()) me_fail("product of incompatible shapes");
    Dense<T> o(a.rows(), b.cols());
    for (int j = 0; j < b.cols(); j++)
        for (int i = 0; i < a.rows(); i++) {
            T s = a(i, 0) * b(0, j);
            for (int k = 1; k < a.cols(); k++) s += a(i, k) * b(k, j);
            o(i, j) = s;
        }
    return o;
}
template <typename T, typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type> Dense<T> operator*(const Dense<T> &a, S s) { Dense<T> o(a); o *= (T)s; return o; }
template <typename T, typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type> Dense<T> operator*(S s, const Dense<T> &a) { Dense<T> o(a); for (int k = 0; k < o.size(); k++) o(k) = (T)s * a(k); return o; }
template <typename T, typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type> Dense<T> operator/(const Dense<T> &a, S s) { Dense<T> o(a); o /= (T)s; return o; }
/* views and matrices mix freely */
#define ME_MIX(op)                                                                                                   \
    template <typename T> Dense<T> operator op(const Block<T> &a, const Block<T> &b) { return a.eval() op b.eval(); } \
    template <typename T> Dense<T> operator op(const Block<T> &a, const Dense<T> &b) { return a.eval() op b; }        \
    template <typename T> Dense<T> operator op(const Dense<T> &a, const Block<T> &b) { return a op b.eval(); }
ME_MIX(+)
ME_MIX(-)
ME_MIX(*)
#undef ME_MIX
template <typename T, typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type> Dense<T> operator*(const Block<T> &a, S s) { return a.eval() * s; }
template <typename T, typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type> Dense<T> operator*(S s, const Block<T> &a) { return s * a.eval(); }
template <typename T, typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type> Dense<T> operator/(const Block<T> &a, S s) { return a.eval() / s; }
template <typename T> Dense<T> operator-(const Block<T> &a) { return -a.eval(); }

template <typename T> std::ostream &operator<<(std::ostream &os, const Dense<T> &m)
{
    for (int i = 0; i < m.rows(); i++) {
        for (int j = 0; j < m.cols(); j++) os << (j ? " " : "") << m(i, j);
        if (i + 1 < m.rows()) os << "\n";
    }
    return os;
}
template <typename T> std::ostream &operator<<(std::ostream &os, const Block<T> &b) { return os << b.eval(); }

/* inverses by cofactors (what Eigen does up to 4 x 4) */
template <typename T> Dense<T> Dense<T>::inverse_() const
{
    if (r_ != c_) me_fail("inverse of a non-square matrix");
    const Dense &m = *this;
    Dense o(r_, c_);
    if (r_ == 2) {
        const T det = m(0, 0) * m(1, 1) - m(1, 0) * m(0, 1), inv = T(1) / det;
        o(0, 0) = m(1, 1) * inv;
        o(1, 0) = -m(1, 0) * inv;
        o(0, 1) = -m(0, 1) * inv;
        o(1, 1) = m(0, 0) * inv;
        return o;
    }
    if (r_ == 3) {
        auto cof = [&](int i, int j) {
            const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
            return m(i1, j1) * m(i2, j2) - m(i1, j2) * m(i2, j1);
        };
        const T c00 = cof(0, 0), c10 = cof(1, 0), c20 = cof(2, 0);
        const T det = c00 * m(0, 0) + c10 * m(1, 0) + c20 * m(2, 0), inv = T(1) / det;
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) o(j, i) = cof(i, j) * inv;
        return o;
    }
    if (r_ == 4) {
        /* cofactor expansion through 3 x 3 minors */
        auto minor3 = [&](int r0, int r1, int r2, int c0, int c1, int c2) {
            return m(r0, c0) * (m(r1, c1) * m(r2, c2) - m(r1, c2) * m(r2, c1)) - m(r0, c1) * (m(r1, c0) * m(r2, c2) - m(r1, c2) * m(r2, c0)) +
                   m(r0, c2) * (m(r1, c0) * m(r2, c1) - m(r1, c1) * m(r2, c0));
        };
        T cofm[4][4];
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++) {
                int rr[3], cc[3], a = 0, b = 0;
                for (int k = 0; k < 4; k++) {
                    if (k != i) rr[a++] = k;
                    if (k != j) cc[b++] = k;
                }
                const T mn = minor3(rr[0], rr[1], rr[2], cc[0], cc[1], cc[2]);
                cofm[i][j] = ((i + j) & 1) ? -mn : mn;
            }
        T det = T(0);
        for (int j = 0; j < 4; j++) det += m(0, j) * cofm[0][j];
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++) o(j, i) = cofm[i][j] / det;
        return o;
    }
    me_fail("inverse is provided up to 4 x 4");
}

/* ---- the comma initialiser ---- */
template <typename T>
class CommaInit {
public:
    CommaInit(Dense<T> &m, T s) : m_(m), row_(0), col_(1), cur_(1) { m_(0, 0) = s; }
    CommaInit(Dense<T> &m, const Dense<T> &o) : m_(m), row_(0), col_(0), cur_(o.rows()) { place_(o); }
    CommaInit &operator,(T s)
    {
        if (col_ == m_.cols()) {
            row_ += cur_;
            col_ = 0;
            cur_ = 1;
        }
        m_(row_, col_++) = s;
        return *this;
    }
    CommaInit &operator,(const Dense<T> &o)
    {
        if (col_ == m_.cols()) {
            row_ += cur_;
            col_ = 0;
            cur_ = o.rows();
        }
        place_(o);
        return *this;
    }
    CommaInit &operator,(const Block<T> &b) { return (*this), b.eval(); }

private:
    void place_(const Dense<T> &o)
    {
        Dense<T> src = o;
        /* a vector may be given in either orientation */
        if (row_ + src.rows() > m_.rows() || col_ + src.cols() > m_.cols()) src = o.transpose_();
        if (row_ + src.rows() > m_.rows() || col_ + src.cols() > m_.cols()) me_fail("comma initialiser: too many coefficients");
        for (int j = 0; j < src.cols(); j++)
            for (int i = 0; i < src.rows(); i++) m_(row_ + i, col_ + j) = src(i, j);
        cur_ = src.rows();
        col_ += src.cols();
    }
    Dense<T> &m_;
    int row_, col_, cur_;
};
template <typename T, typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type> CommaInit<T> operator<<(Dense<T> &m, S s) { return CommaInit<T>(m, (T)s); }
template <typename T> CommaInit<T> operator<<(Dense<T> &m, const Dense<T> &o) { return CommaInit<T>(m, o); }
template <typename T> CommaInit<T> operator<<(Dense<T> &m, const Block<T> &o) { return CommaInit<T>(m, o.eval()); }

/* ---- the Matrix<T, R, C> family: sizes in the type only decide the default shape ---- */
template <typename T, int R, int C>
class Matrix : public Dense<T> {
public:
    Matrix() : Dense<T>(R == Dynamic ? 0 : R, C == Dynamic ? 0 : C) { init_(); }
    explicit Matrix(int n) : Dense<T>(C == 1 ? n : (R == 1 ? 1 : n), C == 1 ? 1 : (R == 1 ? n : 1)) { init_(); }
    /* two arguments: the coefficients of a fixed 2-vector, else the sizes (as Eigen) */
    template <typename A, typename B, typename = typename std::enable_if<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value>::type>
    Matrix(A a, B b) : Dense<T>()
    {
        if ((R == 2 && C == 1) || (R == 1 && C == 2)) {
            Dense<T>::resize(R, C);
            init_();
            (*this)(0) = (T)a;
            (*this)(1) = (T)b;
        } else {
            Dense<T>::resize((int)a, (int)b);
            init_();
        }
    }
    Matrix(T a, T b, T c) : Dense<T>(R == 1 ? 1 : 3, R == 1 ? 3 : 1) { init_(); (*this)(0) = a; (*this)(1) = b; (*this)(2) = c; }
    Matrix(T a, T b, T c, T d) : Dense<T>(R == 1 ? 1 : 4, R == 1 ? 4 : 1) { init_(); (*this)(0) = a; (*this)(1) = b; (*this)(2) = c; (*this)(3) = d; }
    Matrix(const Dense<T> &m) : Dense<T>() { init_(); adopt_(m); }
    Matrix(const Block<T> &b) : Dense<T>() { init_(); adopt_(b.eval()); }
    Matrix(const ArrayX<T> &a) : Dense<T>() { init_(); adopt_(a.v); }
    Matrix &operator=(const Dense<T> &m) { adopt_(m); return *this; }
    Matrix &operator=(const Block<T> &b) { adopt_(b.eval()); return *this; }
    Matrix &operator=(const ArrayX<T> &a) { adopt_(a.v); return *this; }
    static Matrix Zero() { return Matrix(); }
    static Matrix Zero(int r, int c) { Matrix m; m.resize(r, c); return m; }
    static Matrix Zero(int n) { return Matrix(n); }
    static Matrix Ones() { Matrix m; m.setOnes(); return m; }
    static Matrix Ones(int r, int c) { Matrix m; m.resize(r, c); m.setOnes(); return m; }
    static Matrix Ones(int n) { Matrix m(n); m.setOnes(); return m; }
    static Matrix Constant(int r, int c, T v) { Matrix m; m.resize(r, c); m.setConstant(v); return m; }
    static Matrix Constant(int n, T v) { Matrix m(n); m.setConstant(v); return m; }
    static Matrix Constant(T v) { Matrix m; m.setConstant(v); return m; }
    static Matrix Identity() { Matrix m; m.setIdentity(); return m; }
    static Matrix Identity(int r, int c) { Matrix m; m.resize(r, c); m.setIdentity(); return m; }

private:
    void init_() { if (R == 1 && C != 1) this->mark_rowvec_(); }
    void adopt_(const Dense<T> &m)
    {
        /* a vector type takes a vector of the other orientation (Eigen transposes implicitly there); fixed sizes must agree */
        const bool want_col = (C == 1), want_row = (R == 1 && C != 1);
        Dense<T> src = m;
        if (want_col && m.cols() != 1 && m.rows() == 1) src = m.transpose_();
        if (want_row && m.rows() != 1 && m.cols() == 1) src = m.transpose_();
        if ((R != Dynamic && src.rows() != R) || (C != Dynamic && src.cols() != C)) me_fail("assignment between fixed sizes that differ");
        static_cast<Dense<T> &>(*this) = src;
        init_();
    }
};

#define ME_TYPEDEFS(T, S)                          \
    typedef Matrix<T, Dynamic, Dynamic> MatrixX##S; \
    typedef Matrix<T, Dynamic, 1> VectorX##S;       \
    typedef Matrix<T, 1, Dynamic> RowVectorX##S;    \
    typedef Matrix<T, 2, 2> Matrix2##S;             \
    typedef Matrix<T, 3, 3> Matrix3##S;             \
    typedef Matrix<T, 4, 4> Matrix4##S;             \
    typedef Matrix<T, 2, 1> Vector2##S;             \
    typedef Matrix<T, 3, 1> Vector3##S;             \
    typedef Matrix<T, 4, 1> Vector4##S;             \
    typedef Matrix<T, 2, Dynamic> Matrix2X##S;      \
    typedef Matrix<T, 3, Dynamic> Matrix3X##S;      \
    typedef Matrix<T, 4, Dynamic> Matrix4X##S;      \
    typedef Matrix<T, Dynamic, 2> MatrixX2##S;      \
    typedef Matrix<T, Dynamic, 3> MatrixX3##S;      \
    typedef Matrix<T, Dynamic, 4> MatrixX4##S;
ME_TYPEDEFS(double, d)
ME_TYPEDEFS(float, f)
ME_TYPEDEFS(int, i)
#undef ME_TYPEDEFS

/* ---- quaternions ---- */
template <typename T>
class Quaternion {
public:
    Quaternion() : w_(1), x_(0), y_(0), z_(0) {}
    Quaternion(T w, T x, T y, T z) : w_(w), x_(x), y_(y), z_(z) {}
    explicit Quaternion(const Dense<T> &m)
    {
        /* Eigen's quaternionbase_assign_impl for a 3 x 3 rotation: trace branch, else the largest diagonal element */
        T t = m(0, 0) + m(1, 1) + m(2, 2);
        if (t > T(0)) {
            t = std::sqrt(t + T(1.0));
            w_ = T(0.5) * t;
            t = T(0.5) / t;
            x_ = (m(2, 1) - m(1, 2)) * t;
            y_ = (m(0, 2) - m(2, 0)) * t;
            z_ = (m(1, 0) - m(0, 1)) * t;
        } else {
            int i = 0;
            if (m(1, 1) > m(0, 0)) i = 1;
            if (m(2, 2) > m(i, i)) i = 2;
            const int j = (i + 1) % 3, k = (j + 1) % 3;
            t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + T(1.0));
            T q[3];
            q[i] = T(0.5) * t;
            t = T(0.5) / t;
            w_ = (m(k, j) - m(j, k)) * t;
            q[j] = (m(j, i) + m(i, j)) * t;
            q[k] = (m(k, i) + m(i, k)) * t;
            x_ = q[0];
            y_ = q[1];
            z_ = q[2];
        }
    }
    T &w() { return w_; }
    T &x() { return x_; }
    T &y() { return y_; }
    T &z() { return z_; }
    const T &w() const { return w_; }
    const T &x() const { return x_; }
    const T &y() const { return y_; }
    const T &z() const { return z_; }
    Matrix<T, 3, 3> toRotationMatrix() const
    {
        Matrix<T, 3, 3> r;
        const T tx = T(2) * x_, ty = T(2) * y_, tz = T(2) * z_;
        const T twx = tx * w_, twy = ty * w_, twz = tz * w_, txx = tx * x_, txy = ty * x_, txz = tz * x_, tyy = ty * y_, tyz = tz * y_, tzz = tz * z_;
        r(0, 0) = T(1) - (tyy + tzz);
        r(0, 1) = txy - twz;
        r(0, 2) = txz + twy;
        r(1, 0) = txy + twz;
        r(1, 1) = T(1) - (txx + tzz);
        r(1, 2) = tyz - twx;
        r(2, 0) = txz - twy;
        r(2, 1) = tyz + twx;
        r(2, 2) = T(1) - (txx + tyy);
        return r;
    }

private:
    T w_, x_, y_, z_;
};
typedef Quaternion<double> Quaterniond;
typedef Quaternion<float> Quaternionf;

}  // namespace Eigen
#endif /* ORC_MINIEIGEN_HPP */
