// cvshim.hpp - the slice of the OpenCV core API that the reference's usac/ sources touch, so that those sources can be
// compiled WHERE THEY LIE under /root/reference (oracle/Makefile.ref -> oracle/_ref/libusac_ref.so) in an image that has no
// OpenCV C++ headers. TEST INFRASTRUCTURE ONLY (oracle/ header contract): nothing in ransac_b200/ may include this.
//
// What is faithful and what is not:
//  * cv::Mat here is a 2-D, single-channel, row-major matrix of float / double / int with OpenCV's header semantics:
//    copy = shared data, row()/rowRange()/colRange()/operator()(Range,Range) are views, clone()/copyTo() copy,
//    `view = expression` writes INTO the view when size and type match (cv::Mat::operator=(const MatExpr&) -> create()),
//    outputs of SVD/eigen/invert are written in place into pre-allocated matrices of the right size (OutputArray::create).
//  * The numerical routines are this file's own: one-sided Jacobi SVD in the matrix's own precision, cyclic Jacobi
//    eigen-solver, LU determinant / inverse, closed-form 2x2 / 3x3 inverse in double (what OpenCV does for float input),
//    solveCubic by the trigonometric / Cardano method in double. They are checked against Python cv2 (tests/golden/
//    cv_primitives.npz) but are NOT OpenCV's code: singular vectors of distinct singular values agree up to sign, a basis
//    of a null space of dimension > 1 is an arbitrary orthonormal basis (OpenCV's own choice differs between its Jacobi
//    and LAPACK builds, too). Everything the reference computes THROUGH such a basis (root order of the five-point solver)
//    inherits that freedom - see DESIGN.md section 3.
#pragma once
#include <algorithm>
#include <cassert>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <iostream>
#include <list>
#include <map>
#include <memory>
#include <numeric>
#include <queue>
#include <set>
#include <string>
#include <vector>

#define CV_8U 0
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_32FC1 CV_32F
#define CV_64FC1 CV_64F
#define CV_32SC1 CV_32S
#define CV_PI 3.1415926535897932384626433832795

namespace cv {

template <class T> struct DepthOf;
template <> struct DepthOf<float> { enum { value = CV_32F }; };
template <> struct DepthOf<double> { enum { value = CV_64F }; };
template <> struct DepthOf<int> { enum { value = CV_32S }; };
template <> struct DepthOf<unsigned char> { enum { value = CV_8U }; };
template <class T> struct Point_;
template <> struct DepthOf<Point_<float>> { enum { value = CV_32F }; };     // evsac_sampler.hpp:65 reads a row of an N x 2 matrix as a point

inline size_t elem_size(int type) { return type == CV_64F ? 8 : type == CV_8U ? 1 : 4; }

struct Range {
    int start, end;
    Range() : start(0), end(0) {}
    Range(int s, int e) : start(s), end(e) {}
    static Range all() { return Range(INT32_MIN, INT32_MAX); }
    bool is_all() const { return start == INT32_MIN && end == INT32_MAX; }
};

template <class T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    template <class U> Point_(const Point_<U>& o) : x((T)o.x), y((T)o.y) {}
};
typedef Point_<int> Point;
typedef Point_<int> Point2i;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;
struct Scalar { double v[4]; Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { v[0] = a; v[1] = b; v[2] = c; v[3] = d; } };
struct Size { int width, height; Size(int w = 0, int h = 0) : width(w), height(h) {} };

enum { DECOMP_LU = 0, DECOMP_SVD = 1 };

class MatExpr;

class Mat {
public:
    int rows = 0, cols = 0;
    unsigned char* data = nullptr;
    size_t step = 0;                 // bytes between rows
    int type_ = CV_32F;
    std::shared_ptr<std::vector<unsigned char>> buf;   // null: borrowed memory

    Mat() = default;
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(int r, int c, int type, void* borrowed) : rows(r), cols(c), data((unsigned char*)borrowed), step((size_t)c * elem_size(type)), type_(type) {}
    Mat(const MatExpr& e);
    template <class T> explicit Mat(const std::vector<T>& v) { create((int)v.size(), 1, DepthOf<T>::value); if (!v.empty()) std::memcpy(data, v.data(), v.size() * sizeof(T)); }

    // OutputArray::create: keep the storage when size and type already match
    void create(int r, int c, int type) {
        if (data && rows == r && cols == c && type_ == type) return;
        rows = r; cols = c; type_ = type; step = (size_t)c * elem_size(type);
        buf = std::make_shared<std::vector<unsigned char>>((size_t)r * step + 16, (unsigned char)0);
        data = buf->data();
    }
    void release() { rows = cols = 0; data = nullptr; step = 0; buf.reset(); }
    int type() const { return type_; }
    int depth() const { return type_; }
    int channels() const { return 1; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    size_t total() const { return (size_t)rows * cols; }
    Size size() const { return Size(cols, rows); }
    bool isContinuous() const { return step == (size_t)cols * elem_size(type_); }

    template <class T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data + (size_t)r * step); }
    template <class T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + (size_t)r * step); }
    unsigned char* ptr(int r = 0) { return data + (size_t)r * step; }
    const unsigned char* ptr(int r = 0) const { return data + (size_t)r * step; }
    template <class T> T& at(int r, int c) { assert(DepthOf<T>::value == type_); return ptr<T>(r)[c]; }
    template <class T> const T& at(int r, int c) const { assert(DepthOf<T>::value == type_); return ptr<T>(r)[c]; }
    template <class T> T& at(int i) { if (sizeof(T) > elem_size(type_)) return *reinterpret_cast<T*>(ptr(i)); return (rows == 1) ? at<T>(0, i) : (cols == 1 ? at<T>(i, 0) : at<T>(i / cols, i % cols)); }
    template <class T> const T& at(int i) const { return (rows == 1) ? at<T>(0, i) : (cols == 1 ? at<T>(i, 0) : at<T>(i / cols, i % cols)); }

    double get(int r, int c) const {
        switch (type_) {
            case CV_64F: return ptr<double>(r)[c];
            case CV_32S: return ptr<int>(r)[c];
            case CV_8U: return ptr<unsigned char>(r)[c];
            default: return ptr<float>(r)[c];
        }
    }
    void set(int r, int c, double v) {
        switch (type_) {
            case CV_64F: ptr<double>(r)[c] = v; break;
            case CV_32S: ptr<int>(r)[c] = (int)std::lrint(v); break;
            case CV_8U: ptr<unsigned char>(r)[c] = (unsigned char)v; break;
            default: ptr<float>(r)[c] = (float)v; break;
        }
    }

    Mat view(int r0, int r1, int c0, int c1) const {
        Mat m;
        m.rows = r1 - r0; m.cols = c1 - c0; m.type_ = type_; m.step = step; m.buf = buf;
        m.data = data + (size_t)r0 * step + (size_t)c0 * elem_size(type_);
        return m;
    }
    Mat row(int r) const { return view(r, r + 1, 0, cols); }
    Mat col(int c) const { return view(0, rows, c, c + 1); }
    Mat rowRange(int a, int b) const { return view(a, b, 0, cols); }
    Mat colRange(int a, int b) const { return view(0, rows, a, b); }
    Mat rowRange(const Range& r) const { return r.is_all() ? *this : rowRange(r.start, r.end); }
    Mat colRange(const Range& r) const { return r.is_all() ? *this : colRange(r.start, r.end); }
    Mat operator()(const Range& rr, const Range& cr) const {
        return view(rr.is_all() ? 0 : rr.start, rr.is_all() ? rows : rr.end, cr.is_all() ? 0 : cr.start, cr.is_all() ? cols : cr.end);
    }
    Mat clone() const {
        Mat m;
        if (empty()) return m;
        m.create(rows, cols, type_);
        for (int r = 0; r < rows; r++) std::memcpy(m.ptr(r), ptr(r), (size_t)cols * elem_size(type_));
        return m;
    }
    void copyTo(Mat& dst) const {
        if (empty()) { dst.release(); return; }
        if (dst.data == data && dst.rows == rows && dst.cols == cols) return;
        Mat tmp = clone();                      // source and destination may alias (nearest_neighbors.cpp:154)
        dst.create(rows, cols, type_);
        for (int r = 0; r < rows; r++) std::memcpy(dst.ptr(r), tmp.ptr(r), (size_t)cols * elem_size(type_));
    }
    void convertTo(Mat& dst, int type) const {
        Mat tmp;
        tmp.create(rows, cols, type);
        for (int r = 0; r < rows; r++) for (int c = 0; c < cols; c++) tmp.set(r, c, get(r, c));
        dst = tmp;
    }
    // reshape(cn, rows): the reference only uses reshape(3, 3) on a 1 x 9 row (-> read back as a 3 x 3 single-channel matrix by
    // the Mat_<float> constructor) and reshape(1) (no-op for single-channel data)
    Mat reshape(int cn, int new_rows = 0) const {
        if (new_rows == 0 || new_rows == rows) return *this;
        assert(isContinuous() && total() % new_rows == 0);
        (void)cn;
        Mat m = *this;
        m.rows = new_rows; m.cols = (int)(total() / new_rows); m.step = (size_t)m.cols * elem_size(type_);
        return m;
    }
    void push_back(const Mat& r) {
        if (empty()) { *this = r.clone(); return; }
        assert(r.cols == cols && r.type_ == type_);
        Mat m;
        m.create(rows + r.rows, cols, type_);
        for (int i = 0; i < rows; i++) std::memcpy(m.ptr(i), ptr(i), (size_t)cols * elem_size(type_));
        for (int i = 0; i < r.rows; i++) std::memcpy(m.ptr(rows + i), r.ptr(i), (size_t)cols * elem_size(type_));
        *this = m;
    }
    Mat& setTo(double v) { for (int r = 0; r < rows; r++) for (int c = 0; c < cols; c++) set(r, c, v); return *this; }

    Mat& operator=(const Mat&) = default;
    Mat(const Mat&) = default;
    Mat& operator=(const MatExpr& e);            // evaluates INTO this matrix when size/type match (view assignment), else rebinds

    MatExpr t() const;
    MatExpr inv(int method = DECOMP_LU) const;
    MatExpr mul(const Mat& o, double scale = 1) const;
    Mat cross(const Mat& o) const;
    double dot(const Mat& o) const {
        double s = 0;
        for (int r = 0; r < rows; r++) for (int c = 0; c < cols; c++) s += get(r, c) * o.get(r, c);
        return s;
    }
    static MatExpr zeros(int r, int c, int type);
    static MatExpr ones(int r, int c, int type);
    static MatExpr eye(int r, int c, int type);
};

class MatExpr : public Mat {
public:
    MatExpr() = default;
    explicit MatExpr(const Mat& m) : Mat(m) {}
};

inline Mat::Mat(const MatExpr& e) : Mat(static_cast<const Mat&>(e)) {}
inline Mat& Mat::operator=(const MatExpr& e) {
    if (data && !e.empty() && rows == e.rows && cols == e.cols && type_ == e.type_) {
        if (data != e.data) {
            Mat tmp = e.clone();
            for (int r = 0; r < rows; r++) std::memcpy(ptr(r), tmp.ptr(r), (size_t)cols * elem_size(type_));
        }
        return *this;
    }
    return *this = static_cast<const Mat&>(e);
}

// cv::InputArray: a proxy over a Mat or over an arbitrary object (getObj(): napsac_sampler.hpp:68 passes a
// std::vector<std::vector<int>> through it)
class _InputArray {
    Mat m;
    const void* obj = nullptr;
    bool empty_ = true;
public:
    _InputArray(const Mat& mat) : m(mat), obj(&mat), empty_(mat.empty()) {}
    _InputArray(const MatExpr& e) : m(e), obj(nullptr), empty_(e.empty()) {}
    template <class T> _InputArray(const std::vector<T>& v) : obj(&v), empty_(v.empty()) {}
    Mat getMat() const { return m; }
    void* getObj() const { return const_cast<void*>(obj); }
    bool empty() const { return empty_; }
    Size size() const { return m.size(); }
};
typedef const _InputArray& InputArray;
typedef Mat& OutputArray;
typedef Mat& InputOutputArray;

inline MatExpr Mat::zeros(int r, int c, int type) { Mat m; m.create(r, c, type); return MatExpr(m); }
inline MatExpr Mat::ones(int r, int c, int type) { Mat m; m.create(r, c, type); m.setTo(1); return MatExpr(m); }
inline MatExpr Mat::eye(int r, int c, int type) { Mat m; m.create(r, c, type); for (int i = 0; i < std::min(r, c); i++) m.set(i, i, 1); return MatExpr(m); }

// ---- element-wise arithmetic in the matrix's own precision (one rounding per operator, like cv's non-FMA SSE paths)
template <class T, class F> inline Mat binary_t(const Mat& a, const Mat& b, F f) {
    Mat m; m.create(a.rows, a.cols, a.type_);
    for (int r = 0; r < a.rows; r++) for (int c = 0; c < a.cols; c++) m.ptr<T>(r)[c] = f(a.ptr<T>(r)[c], b.ptr<T>(r)[c]);
    return m;
}
template <class F> inline MatExpr binary(const Mat& a, const Mat& b, F f) {
    assert(a.rows == b.rows && a.cols == b.cols && a.type_ == b.type_);
    if (a.type_ == CV_64F) return MatExpr(binary_t<double>(a, b, f));
    if (a.type_ == CV_32S) return MatExpr(binary_t<int>(a, b, f));
    return MatExpr(binary_t<float>(a, b, f));
}
inline MatExpr operator+(const Mat& a, const Mat& b) { return binary(a, b, [](auto x, auto y) { return x + y; }); }
inline MatExpr operator-(const Mat& a, const Mat& b) { return binary(a, b, [](auto x, auto y) { return x - y; }); }
inline MatExpr scale(const Mat& a, double s, bool divide) {
    Mat m; m.create(a.rows, a.cols, a.type_);
    for (int r = 0; r < a.rows; r++)
        for (int c = 0; c < a.cols; c++) {
            if (a.type_ == CV_64F) m.ptr<double>(r)[c] = divide ? a.ptr<double>(r)[c] / s : a.ptr<double>(r)[c] * s;
            else if (a.type_ == CV_32S) m.ptr<int>(r)[c] = (int)std::lrint(divide ? a.ptr<int>(r)[c] / s : a.ptr<int>(r)[c] * s);
            // cv evaluates A*alpha for float data as (float)(a * (double)alpha); division as a * (1/alpha)
            else m.ptr<float>(r)[c] = divide ? (float)(a.ptr<float>(r)[c] * (1.0 / s)) : (float)(a.ptr<float>(r)[c] * s);
        }
    return MatExpr(m);
}
inline MatExpr operator*(const Mat& a, double s) { return scale(a, s, false); }
inline MatExpr operator*(double s, const Mat& a) { return scale(a, s, false); }
inline MatExpr operator/(const Mat& a, double s) { return scale(a, s, true); }
inline MatExpr operator-(const Mat& a) { return scale(a, -1.0, false); }
// matrix product (cv::gemm): accumulation in double for float data, like cv's generic GEMM kernels
inline MatExpr operator*(const Mat& a, const Mat& b) {
    assert(a.cols == b.rows && a.type_ == b.type_);
    Mat m; m.create(a.rows, b.cols, a.type_);
    for (int r = 0; r < a.rows; r++)
        for (int c = 0; c < b.cols; c++) {
            double s = 0;
            for (int k = 0; k < a.cols; k++) s += a.get(r, k) * b.get(k, c);
            m.set(r, c, s);
        }
    return MatExpr(m);
}
inline MatExpr Mat::t() const {
    Mat m; m.create(cols, rows, type_);
    for (int r = 0; r < rows; r++) for (int c = 0; c < cols; c++) m.set(c, r, get(r, c));
    return MatExpr(m);
}
inline MatExpr Mat::mul(const Mat& o, double s) const {
    Mat m; m.create(rows, cols, type_);
    for (int r = 0; r < rows; r++) for (int c = 0; c < cols; c++) m.set(r, c, get(r, c) * o.get(r, c) * s);
    return MatExpr(m);
}
inline Mat Mat::cross(const Mat& o) const {
    assert(total() == 3 && o.total() == 3 && type_ == o.type_);
    Mat m; m.create(rows, cols, type_);
    auto g = [](const Mat& x, int i) { return x.rows == 1 ? x.get(0, i) : x.get(i, 0); };
    auto put = [&](int i, double v) { if (rows == 1) m.set(0, i, v); else m.set(i, 0, v); };
    if (type_ == CV_32F) {          // float arithmetic, one rounding per operator
        const float a0 = (float)g(*this, 0), a1 = (float)g(*this, 1), a2 = (float)g(*this, 2), b0 = (float)g(o, 0), b1 = (float)g(o, 1), b2 = (float)g(o, 2);
        put(0, a1 * b2 - a2 * b1); put(1, a2 * b0 - a0 * b2); put(2, a0 * b1 - a1 * b0);
    } else {
        const double a0 = g(*this, 0), a1 = g(*this, 1), a2 = g(*this, 2), b0 = g(o, 0), b1 = g(o, 1), b2 = g(o, 2);
        put(0, a1 * b2 - a2 * b1); put(1, a2 * b0 - a0 * b2); put(2, a0 * b1 - a1 * b0);
    }
    return m;
}

inline std::ostream& operator<<(std::ostream& os, const Mat& m) {
    os << "[";
    for (int r = 0; r < m.rows; r++) {
        for (int c = 0; c < m.cols; c++) os << m.get(r, c) << (c + 1 < m.cols ? ", " : "");
        os << (r + 1 < m.rows ? ";\n " : "");
    }
    return os << "]";
}

// ---- Mat_<T> ----------------------------------------------------------------------------------------------------
template <class T> class MatCommaInitializer_;

template <class T> class Mat_ : public Mat {
public:
    Mat_() { type_ = DepthOf<T>::value; }
    Mat_(int r, int c) { create(r, c, DepthOf<T>::value); }
    Mat_(int r, int c, T* borrowed) : Mat(r, c, DepthOf<T>::value, borrowed) {}
    Mat_(int r, int c, const T& value) { create(r, c, DepthOf<T>::value); for (int i = 0; i < r; i++) for (int j = 0; j < c; j++) ptr<T>(i)[j] = value; }
    Mat_(const Mat& m) { assign(m); }
    Mat_(const MatExpr& e) { assign(e); }
    Mat_(const Mat_& m) = default;
    Mat_& operator=(const Mat_& m) = default;
    Mat_& operator=(const Mat& m) { assign(m); return *this; }
    Mat_& operator=(const MatExpr& e) {
        if (e.type_ == DepthOf<T>::value) Mat::operator=(e); else assign(e);
        return *this;
    }
    void assign(const Mat& m) {
        if (m.empty()) { release(); type_ = DepthOf<T>::value; return; }
        if (m.type_ == DepthOf<T>::value) { Mat::operator=(m); return; }
        Mat tmp;
        m.convertTo(tmp, DepthOf<T>::value);
        Mat::operator=(tmp);
    }
    T& operator()(int r, int c) { return ptr<T>(r)[c]; }
    const T& operator()(int r, int c) const { return ptr<T>(r)[c]; }
    T& operator()(int i) { return at<T>(i); }
    Mat operator()(const Range& rr, const Range& cr) const { return Mat::operator()(rr, cr); }
    static MatExpr zeros(int r, int c) { return Mat::zeros(r, c, DepthOf<T>::value); }
    static MatExpr ones(int r, int c) { return Mat::ones(r, c, DepthOf<T>::value); }
    static MatExpr eye(int r, int c) { return Mat::eye(r, c, DepthOf<T>::value); }
    Mat_ clone() const { return Mat_(Mat::clone()); }
};

template <class T> class MatCommaInitializer_ {
public:
    Mat_<T> m;
    size_t pos = 0;
    explicit MatCommaInitializer_(const Mat_<T>& m_) : m(m_) {}
    template <class V> MatCommaInitializer_& operator,(V v) {
        assert(pos < m.total());
        m.template ptr<T>((int)(pos / m.cols))[pos % m.cols] = (T)v;
        pos++;
        return *this;
    }
    operator Mat_<T>() const { return m; }
};
template <class T, class V> inline MatCommaInitializer_<T> operator<<(const Mat_<T>& m, V v) {
    MatCommaInitializer_<T> ci(m);
    return (ci, v);
}

// ---- numerical routines -----------------------------------------------------------------------------------------
namespace shim {

// One-sided Jacobi SVD (Hestenes) of an m x n matrix in precision T. The work matrix is padded with zero rows to
// max(m, n) x n so that V is a complete n x n orthogonal matrix (null-space columns included). Singular values descending.
// Outputs: w[min(m,n)], vt[n x n] (rows = right singular vectors), u[m x m] when want_u (completed to an orthonormal basis).
template <class T>
void jacobi_svd(const Mat& A, std::vector<T>& w, std::vector<T>& vt, std::vector<T>& u, int ucols) {
    const int m = A.rows, n = A.cols, M = std::max(m, n);
    std::vector<T> a((size_t)M * n, T(0)), v((size_t)n * n, T(0));
    for (int r = 0; r < m; r++) for (int c = 0; c < n; c++) a[(size_t)r * n + c] = (T)A.get(r, c);
    for (int i = 0; i < n; i++) v[(size_t)i * n + i] = 1;
    // convergence like a float / double Jacobi SVD: a pair is orthogonal when |<p,q>| <= eps |p||q|; columns that have both
    // collapsed to the null space (norms below eps^2 of the largest) are left alone - any basis of the null space will do
    const double eps = sizeof(T) == 4 ? 2.0 * FLT_EPSILON : 10.0 * DBL_EPSILON;
    double scale2 = 0;
    for (int c = 0; c < n; c++) { double s2 = 0; for (int r = 0; r < M; r++) s2 += (double)a[(size_t)r * n + c] * a[(size_t)r * n + c]; scale2 = std::max(scale2, s2); }
    const double floor2 = scale2 * eps * eps;
    for (int sweep = 0; sweep < std::max(M, 30); sweep++) {
        bool changed = false;
        for (int p = 0; p < n - 1; p++)
            for (int q = p + 1; q < n; q++) {
                double alpha = 0, beta = 0, gamma = 0;
                for (int r = 0; r < M; r++) {
                    const double x = a[(size_t)r * n + p], y = a[(size_t)r * n + q];
                    alpha += x * x; beta += y * y; gamma += x * y;
                }
                if (std::fabs(gamma) <= eps * std::sqrt(alpha * beta) || gamma == 0 || (alpha <= floor2 && beta <= floor2)) continue;
                changed = true;
                const double zeta = (beta - alpha) / (2 * gamma);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1 + zeta * zeta));
                const T c = (T)(1 / std::sqrt(1 + t * t)), s = (T)(c * t);
                for (int r = 0; r < M; r++) {
                    const T x = a[(size_t)r * n + p], y = a[(size_t)r * n + q];
                    a[(size_t)r * n + p] = c * x - s * y;
                    a[(size_t)r * n + q] = s * x + c * y;
                }
                for (int r = 0; r < n; r++) {
                    const T x = v[(size_t)r * n + p], y = v[(size_t)r * n + q];
                    v[(size_t)r * n + p] = c * x - s * y;
                    v[(size_t)r * n + q] = s * x + c * y;
                }
            }
        if (!changed) break;
    }
    std::vector<double> norm(n);
    std::vector<int> order(n);
    for (int j = 0; j < n; j++) {
        double s = 0;
        for (int r = 0; r < M; r++) s += (double)a[(size_t)r * n + j] * a[(size_t)r * n + j];
        norm[j] = std::sqrt(s);
        order[j] = j;
    }
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return norm[x] > norm[y]; });
    const int k = std::min(m, n);
    w.assign(k, T(0));
    vt.assign((size_t)n * n, T(0));
    for (int i = 0; i < n; i++) {
        const int j = order[i];
        if (i < k) w[i] = (T)norm[j];
        for (int r = 0; r < n; r++) vt[(size_t)i * n + r] = v[(size_t)r * n + j];
    }
    if (ucols <= 0) return;
    // U (m x ucols; ucols = min(m, n) for the thin form, m for FULL_UV): normalised columns of A V for non-zero singular values,
    // completed by Gram-Schmidt against the unit vectors
    u.assign((size_t)m * ucols, T(0));
    std::vector<std::vector<double>> cols;
    const double tiny = eps * (norm[order[0]] > 0 ? norm[order[0]] : 1.0) * std::max(m, n);
    for (int i = 0; i < k && (int)cols.size() < ucols; i++) {
        const int j = order[i];
        std::vector<double> cvec(m);
        if (norm[j] > tiny) { for (int r = 0; r < m; r++) cvec[r] = a[(size_t)r * n + j] / norm[j]; cols.push_back(cvec); }
        else break;
    }
    for (int e = 0; e < m && (int)cols.size() < ucols; e++) {
        std::vector<double> cvec(m, 0.0);
        cvec[e] = 1;
        for (int pass = 0; pass < 2; pass++)
            for (auto& b : cols) { double d = 0; for (int r = 0; r < m; r++) d += b[r] * cvec[r]; for (int r = 0; r < m; r++) cvec[r] -= d * b[r]; }
        double nn = 0;
        for (int r = 0; r < m; r++) nn += cvec[r] * cvec[r];
        if (nn < 1e-8) continue;
        nn = std::sqrt(nn);
        for (int r = 0; r < m; r++) cvec[r] /= nn;
        cols.push_back(cvec);
    }
    for (int j = 0; j < (int)cols.size(); j++) for (int r = 0; r < m; r++) u[(size_t)r * ucols + j] = (T)cols[j][r];
}

template <class T> void put(Mat& dst, int r, int c, const T* src, size_t ld) {
    dst.create(r, c, DepthOf<T>::value);
    for (int i = 0; i < r; i++) for (int j = 0; j < c; j++) dst.ptr<T>(i)[j] = src[(size_t)i * ld + j];
}

// The thin SVD of a WIDE matrix (m < n; the 8 x 9 system of DLt::DLT4p, dlt.cpp:43) the way cv::SVD does it: OpenCV transposes a
// wide matrix and orthogonalises the m ROWS of A (stored in T, dot products and angles in double); row i ends as sigma_i v_i', so
// the right singular vectors are the normalised rows themselves and not a product of rotations. That matters for DLT4p: its rows
// all have the scale of the largest entries (1e6 for pixel coordinates), its columns range from 1 to 1e6, and the vector it
// takes belongs to sigma_8 ~ 0.1. Rotating columns in float32 (jacobi_svd above) loses that vector (30 % off cv2.SVDecomp on the
// same matrices); rotating rows reproduces OpenCV's to ~1e-5 (tests/golden/cv_primitives.npz, test_ref_build.py).
template <class T>
void jacobi_svd_wide(const Mat& A, std::vector<T>& w, std::vector<T>& vt, std::vector<T>& u) {
    const int m = A.rows, n = A.cols;
    std::vector<T> a((size_t)m * n), g((size_t)m * m, T(0));
    for (int r = 0; r < m; r++) for (int c = 0; c < n; c++) a[(size_t)r * n + c] = (T)A.get(r, c);
    for (int i = 0; i < m; i++) g[(size_t)i * m + i] = 1;
    const double eps = sizeof(T) == 4 ? 2.0 * FLT_EPSILON : 10.0 * DBL_EPSILON;
    for (int sweep = 0; sweep < std::max(n, 30); sweep++) {
        bool changed = false;
        for (int p = 0; p < m - 1; p++)
            for (int q = p + 1; q < m; q++) {
                double alpha = 0, beta = 0, gamma = 0;
                for (int c = 0; c < n; c++) {
                    const double x = a[(size_t)p * n + c], y = a[(size_t)q * n + c];
                    alpha += x * x; beta += y * y; gamma += x * y;
                }
                if (gamma == 0 || std::fabs(gamma) <= eps * std::sqrt(alpha * beta)) continue;
                changed = true;
                const double zeta = (beta - alpha) / (2 * gamma);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1 + zeta * zeta));
                const double c = 1 / std::sqrt(1 + t * t), sn = c * t;
                for (int k = 0; k < n; k++) {
                    const double x = a[(size_t)p * n + k], y = a[(size_t)q * n + k];
                    a[(size_t)p * n + k] = (T)(c * x - sn * y);
                    a[(size_t)q * n + k] = (T)(sn * x + c * y);
                }
                for (int k = 0; k < m; k++) {                     // U = (product of the row rotations)': its columns p, q turn the same way
                    const double x = g[(size_t)k * m + p], y = g[(size_t)k * m + q];
                    g[(size_t)k * m + p] = (T)(c * x - sn * y);
                    g[(size_t)k * m + q] = (T)(sn * x + c * y);
                }
            }
        if (!changed) break;
    }
    std::vector<double> norm(m);
    std::vector<int> order(m);
    for (int i = 0; i < m; i++) {
        double s2 = 0;
        for (int c = 0; c < n; c++) s2 += (double)a[(size_t)i * n + c] * a[(size_t)i * n + c];
        norm[i] = std::sqrt(s2);
        order[i] = i;
    }
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return norm[x] > norm[y]; });
    w.assign(m, T(0)); vt.assign((size_t)m * n, T(0)); u.assign((size_t)m * m, T(0));
    for (int i = 0; i < m; i++) {
        const int j = order[i];
        w[i] = (T)norm[j];
        for (int c = 0; c < n; c++) vt[(size_t)i * n + c] = norm[j] > 0 ? (T)(a[(size_t)j * n + c] / norm[j]) : T(c == i);
        for (int r = 0; r < m; r++) u[(size_t)r * m + i] = g[(size_t)r * m + j];
    }
}

template <class T> void svd_t(const Mat& A, Mat& w, Mat& u, Mat& vt, bool full) {
    std::vector<T> ws, vts, us;
    const int m = A.rows, n = A.cols, k = std::min(m, n);
    if (!full && m < n) {
        jacobi_svd_wide<T>(A, ws, vts, us);
        put<T>(w, m, 1, ws.data(), 1);
        put<T>(vt, m, n, vts.data(), n);
        put<T>(u, m, m, us.data(), m);
        return;
    }
    jacobi_svd<T>(A, ws, vts, us, full ? m : k);
    put<T>(w, k, 1, ws.data(), 1);
    put<T>(vt, full ? n : k, n, vts.data(), n);
    put<T>(u, m, full ? m : k, us.data(), full ? m : k);
}
inline void svd(const Mat& A, Mat& w, Mat& u, Mat& vt, int flags) {
    const bool full = (flags & 4) != 0;
    if (A.type_ == CV_64F) svd_t<double>(A, w, u, vt, full); else svd_t<float>(A, w, u, vt, full);
}

// LU with partial pivoting in double: determinant, inverse
inline bool lu(std::vector<double>& a, int n, std::vector<int>& perm, int& sign) {
    perm.resize(n); sign = 1;
    for (int i = 0; i < n; i++) perm[i] = i;
    for (int k = 0; k < n; k++) {
        int p = k;
        for (int i = k + 1; i < n; i++) if (std::fabs(a[(size_t)i * n + k]) > std::fabs(a[(size_t)p * n + k])) p = i;
        if (std::fabs(a[(size_t)p * n + k]) < DBL_EPSILON) return false;
        if (p != k) { for (int j = 0; j < n; j++) std::swap(a[(size_t)p * n + j], a[(size_t)k * n + j]); std::swap(perm[p], perm[k]); sign = -sign; }
        for (int i = k + 1; i < n; i++) {
            const double f = a[(size_t)i * n + k] / a[(size_t)k * n + k];
            a[(size_t)i * n + k] = f;
            for (int j = k + 1; j < n; j++) a[(size_t)i * n + j] -= f * a[(size_t)k * n + j];
        }
    }
    return true;
}
}   // namespace shim

inline double determinant(const Mat& m) {
    assert(m.rows == m.cols);
    const int n = m.rows;
    if (n == 2) return m.get(0, 0) * m.get(1, 1) - m.get(0, 1) * m.get(1, 0);
    if (n == 3)
        return m.get(0, 0) * (m.get(1, 1) * m.get(2, 2) - m.get(1, 2) * m.get(2, 1)) - m.get(0, 1) * (m.get(1, 0) * m.get(2, 2) - m.get(1, 2) * m.get(2, 0)) +
               m.get(0, 2) * (m.get(1, 0) * m.get(2, 1) - m.get(1, 1) * m.get(2, 0));
    std::vector<double> a((size_t)n * n);
    for (int r = 0; r < n; r++) for (int c = 0; c < n; c++) a[(size_t)r * n + c] = m.get(r, c);
    std::vector<int> perm; int sign;
    if (!shim::lu(a, n, perm, sign)) return 0;
    double d = sign;
    for (int i = 0; i < n; i++) d *= a[(size_t)i * n + i];
    return d;
}

inline double invert(const Mat& src, Mat& dst, int method = DECOMP_LU) {
    assert(src.rows == src.cols);
    const int n = src.rows;
    Mat out; out.create(n, n, src.type_);
    if (method == DECOMP_SVD) {
        Mat w, u, vt;
        shim::svd(src, w, u, vt, 4);
        for (int r = 0; r < n; r++) for (int c = 0; c < n; c++) {
            double s = 0;
            for (int k = 0; k < n; k++) { const double sv = w.get(k, 0); if (sv > DBL_EPSILON) s += vt.get(k, r) * u.get(c, k) / sv; }
            out.set(r, c, s);
        }
        dst = out; return 1;
    }
    if (n == 2) {
        const double a = src.get(0, 0), b = src.get(0, 1), c = src.get(1, 0), d = src.get(1, 1);
        double det = a * d - b * c;
        if (det != 0) { det = 1 / det; out.set(0, 0, d * det); out.set(0, 1, -b * det); out.set(1, 0, -c * det); out.set(1, 1, a * det); }
        dst = out; return det;
    }
    if (n == 3) {   // closed form in double, rounded once to the destination type (cv::invert's 3x3 path); det == 0 -> zeros
        double s[9];
        for (int i = 0; i < 9; i++) s[i] = src.get(i / 3, i % 3);
        double d = s[0] * (s[4] * s[8] - s[5] * s[7]) - s[1] * (s[3] * s[8] - s[5] * s[6]) + s[2] * (s[3] * s[7] - s[4] * s[6]);
        if (d != 0) {
            d = 1 / d;
            double t[9];
            t[0] = (s[4] * s[8] - s[5] * s[7]) * d; t[1] = (s[2] * s[7] - s[1] * s[8]) * d; t[2] = (s[1] * s[5] - s[2] * s[4]) * d;
            t[3] = (s[5] * s[6] - s[3] * s[8]) * d; t[4] = (s[0] * s[8] - s[2] * s[6]) * d; t[5] = (s[2] * s[3] - s[0] * s[5]) * d;
            t[6] = (s[3] * s[7] - s[4] * s[6]) * d; t[7] = (s[1] * s[6] - s[0] * s[7]) * d; t[8] = (s[0] * s[4] - s[1] * s[3]) * d;
            for (int i = 0; i < 9; i++) out.set(i / 3, i % 3, t[i]);
        }
        dst = out; return d;
    }
    std::vector<double> a((size_t)n * n);
    for (int r = 0; r < n; r++) for (int c = 0; c < n; c++) a[(size_t)r * n + c] = src.get(r, c);
    std::vector<int> perm; int sign;
    if (!shim::lu(a, n, perm, sign)) { dst = out; return 0; }
    for (int col = 0; col < n; col++) {
        std::vector<double> x(n);
        for (int i = 0; i < n; i++) {
            double s = (perm[i] == col) ? 1.0 : 0.0;
            for (int j = 0; j < i; j++) s -= a[(size_t)i * n + j] * x[j];
            x[i] = s;
        }
        for (int i = n - 1; i >= 0; i--) {
            double s = x[i];
            for (int j = i + 1; j < n; j++) s -= a[(size_t)i * n + j] * x[j];
            x[i] = s / a[(size_t)i * n + i];
        }
        for (int i = 0; i < n; i++) out.set(i, col, x[i]);
    }
    dst = out; return 1;
}
inline MatExpr Mat::inv(int method) const { Mat d; invert(*this, d, method); return MatExpr(d); }

class SVD {
public:
    enum { MODIFY_A = 1, NO_UV = 2, FULL_UV = 4 };
    Mat u, w, vt;
    SVD() = default;
    SVD(const Mat& A, int flags = 0) { (*this)(A, flags); }
    SVD& operator()(const Mat& A, int flags = 0) { shim::svd(A, w, u, vt, flags); return *this; }
    static void compute(const Mat& A, Mat& w, Mat& u, Mat& vt, int flags = 0) {
        if (A.empty()) { w.release(); u.release(); vt.release(); return; }
        Mat ww, uu, vv;
        shim::svd(A, ww, uu, vv, flags);
        write(w, ww); write(u, uu); write(vt, vv);
    }
    static void compute(const Mat& A, Mat& w, int flags = 0) { Mat u, vt; compute(A, w, u, vt, flags); }
    static void write(Mat& dst, const Mat& src) {      // OutputArray::create + copy: in place when the size already fits
        dst.create(src.rows, src.cols, src.type_);
        for (int r = 0; r < src.rows; r++) std::memcpy(dst.ptr(r), src.ptr(r), (size_t)src.cols * elem_size(src.type_));
    }
};
inline void SVDecomp(const Mat& A, Mat& w, Mat& u, Mat& vt, int flags = 0) { SVD::compute(A, w, u, vt, flags); }

// symmetric eigen-decomposition: eigenvalues descending, eigenvectors as ROWS (cv::eigen)
inline bool eigen(const Mat& src, Mat& evals, Mat& evecs) {
    const int n = src.rows;
    std::vector<double> a((size_t)n * n), v((size_t)n * n, 0.0);
    for (int r = 0; r < n; r++) for (int c = 0; c < n; c++) a[(size_t)r * n + c] = src.get(r, c);
    for (int i = 0; i < n; i++) v[(size_t)i * n + i] = 1;
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0;
        for (int p = 0; p < n; p++) for (int q = p + 1; q < n; q++) off += a[(size_t)p * n + q] * a[(size_t)p * n + q];
        if (off < 1e-300) break;
        for (int p = 0; p < n - 1; p++)
            for (int q = p + 1; q < n; q++) {
                const double apq = a[(size_t)p * n + q];
                if (apq == 0) continue;
                const double theta = (a[(size_t)q * n + q] - a[(size_t)p * n + p]) / (2 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1));
                const double c = 1 / std::sqrt(t * t + 1), s = t * c;
                for (int k = 0; k < n; k++) {
                    const double akp = a[(size_t)k * n + p], akq = a[(size_t)k * n + q];
                    a[(size_t)k * n + p] = c * akp - s * akq; a[(size_t)k * n + q] = s * akp + c * akq;
                }
                for (int k = 0; k < n; k++) {
                    const double apk = a[(size_t)p * n + k], aqk = a[(size_t)q * n + k];
                    a[(size_t)p * n + k] = c * apk - s * aqk; a[(size_t)q * n + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < n; k++) {
                    const double vkp = v[(size_t)k * n + p], vkq = v[(size_t)k * n + q];
                    v[(size_t)k * n + p] = c * vkp - s * vkq; v[(size_t)k * n + q] = s * vkp + c * vkq;
                }
            }
    }
    std::vector<int> order(n);
    for (int i = 0; i < n; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return a[(size_t)x * n + x] > a[(size_t)y * n + y]; });
    evals.create(n, 1, src.type_);
    evecs.create(n, n, src.type_);
    for (int i = 0; i < n; i++) {
        evals.set(i, 0, a[(size_t)order[i] * n + order[i]]);
        for (int k = 0; k < n; k++) evecs.set(i, k, v[(size_t)k * n + order[i]]);
    }
    return true;
}

// roots of c0 x^3 + c1 x^2 + c2 x + c3 (4 coefficients) or the monic cubic (3 coefficients): number of DISTINCT real roots
inline int solveCubic(const Mat& coeffs, Mat& roots) {
    const int nc = (int)coeffs.total();
    auto cf = [&](int i) { return coeffs.rows == 1 ? coeffs.get(0, i) : coeffs.get(i, 0); };
    double a0 = 1, a1, a2, a3;
    if (nc == 4) { a0 = cf(0); a1 = cf(1); a2 = cf(2); a3 = cf(3); } else { a1 = cf(0); a2 = cf(1); a3 = cf(2); }
    double x0 = 0, x1 = 0, x2 = 0;
    int n = 0;
    if (a0 == 0) {
        if (a1 == 0) {
            if (a2 == 0) n = a3 == 0 ? -1 : 0;
            else { x0 = -a3 / a2; n = 1; }
        } else {
            double d = a2 * a2 - 4 * a1 * a3;
            if (d >= 0) {
                d = std::sqrt(d);
                const double q1 = (-a2 + d) * 0.5, q2 = (a2 + d) * -0.5;
                if (std::fabs(q1) > std::fabs(q2)) { x0 = q1 / a1; x1 = a3 / q1; } else { x0 = q2 / a1; x1 = a3 / q2; }
                n = d > 0 ? 2 : 1;
            }
        }
    } else {
        a0 = 1. / a0; a1 *= a0; a2 *= a0; a3 *= a0;
        const double Q = (a1 * a1 - 3 * a2) * (1. / 9), R = (2 * a1 * a1 * a1 - 9 * a1 * a2 + 27 * a3) * (1. / 54);
        const double Qcubed = Q * Q * Q;
        // OpenCV >= 4.5 expands Qcubed - R R so that the common terms a1^6 / 729 and a1^4 a2 / 81 cancel before rounding (mathfuncs.cpp);
        // with the plain difference one of the 512 cubics of tests/golden/cv_primitives.npz gets another root COUNT than cv2's
        double d = (a1 * a1 * (a2 * a2 - 4 * a1 * a3) + 2 * a2 * (9 * a1 * a3 - 2 * a2 * a2) - 27 * a3 * a3) * (1. / 108);
        if (d > 0) {
            const double theta = std::acos(R / std::sqrt(Qcubed)), sqrtQ = std::sqrt(Q);
            const double t0 = -2 * sqrtQ, t1 = theta * (1. / 3), t2 = a1 * (1. / 3);
            x0 = t0 * std::cos(t1) - t2; x1 = t0 * std::cos(t1 + (2. * CV_PI / 3)) - t2; x2 = t0 * std::cos(t1 + (4. * CV_PI / 3)) - t2;
            n = 3;
        } else if (d == 0) {
            if (R >= 0) { x0 = -2 * std::pow(R, 1. / 3) - a1 / 3; x1 = std::pow(R, 1. / 3) - a1 / 3; }
            else { x0 = 2 * std::pow(-R, 1. / 3) - a1 / 3; x1 = -std::pow(-R, 1. / 3) - a1 / 3; }
            x2 = 0;
            n = x0 == x1 ? 1 : 2;
            x1 = x0 == x1 ? 0 : x1;
        } else {
            d = std::sqrt(-d);
            double e = std::pow(d + std::fabs(R), 1. / 3);
            if (R > 0) e = -e;
            x0 = (e + Q / e) - a1 * (1. / 3);
            n = 1;
        }
    }
    if (roots.empty()) roots.create(1, 3, coeffs.type_);
    const double xs[3] = {x0, x1, x2};
    for (int i = 0; i < 3; i++) { if (roots.rows == 1) roots.set(0, i, xs[i]); else roots.set(i, 0, xs[i]); }
    return n;
}

inline double sampsonDistance(const Mat&, const Mat&, const Mat&) { return 0; }   // debug-only call sites (commented out in the reference)
inline double norm(const Mat& m) { return std::sqrt(m.dot(m)); }

template <class E> inline void cv2eigen(const Mat& src, E& dst) {
    dst.resize(src.rows, src.cols);
    for (int r = 0; r < src.rows; r++) for (int c = 0; c < src.cols; c++) dst(r, c) = (typename E::Scalar)src.get(r, c);
}

namespace flann {
struct LinearIndexParams {};
// brute-force kNN (what a linear flann index computes): squared L2 over all columns, ascending, ties by index
class Index {
    Mat pts;
public:
    Index(const Mat& points, const LinearIndexParams&) : pts(points.clone()) {}
    void knnSearch(const Mat& queries, Mat& indices, Mat& dists, int knn) {
        indices.create(queries.rows, knn, CV_32S); dists.create(queries.rows, knn, CV_32F);
        std::vector<std::pair<float, int>> d(pts.rows);
        for (int q = 0; q < queries.rows; q++) {
            for (int p = 0; p < pts.rows; p++) {
                float s = 0;
                for (int c = 0; c < pts.cols; c++) { const float t = (float)(queries.get(q, c) - pts.get(p, c)); s += t * t; }
                d[p] = {s, p};
            }
            std::partial_sort(d.begin(), d.begin() + knn, d.end());
            for (int k = 0; k < knn; k++) { indices.at<int>(q, k) = d[k].second; dists.at<float>(q, k) = d[k].first; }
        }
    }
};
}   // namespace flann

// highgui / imgproc names that the unfinished ProgressiveNapsac constructor (progressive_sampler.hpp:129-146, debug drawing)
// mentions; never executed by anything the oracle drives
inline void imshow(const char*, const Mat&) {}
inline void imshow(const std::string&, const Mat&) {}
inline int waitKey(int = 0) { return 0; }
inline void circle(Mat&, Point, int, Scalar, int = 1) {}
inline void line(Mat&, Point, Point, Scalar, int = 1) {}
inline void rectangle(Mat&, Point, Point, Scalar, int = 1) {}
inline Mat imread(const std::string&) { return Mat(); }
inline bool imwrite(const std::string&, const Mat&) { return false; }
inline void hconcat(const Mat& a, const Mat& b, Mat& dst) {
    Mat m; m.create(a.rows, a.cols + b.cols, a.type_);
    for (int r = 0; r < a.rows; r++) { for (int c = 0; c < a.cols; c++) m.set(r, c, a.get(r, c)); for (int c = 0; c < b.cols; c++) m.set(r, a.cols + c, b.get(r, c)); }
    dst = m;
}
inline void vconcat(const Mat& a, const Mat& b, Mat& dst) { Mat m = a.clone(); m.push_back(b); dst = m; }
inline void transpose(const Mat& a, Mat& dst) { dst = a.t(); }
}   // namespace cv
