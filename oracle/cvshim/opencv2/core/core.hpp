// cvshim -- a MINIMAL stand-in for the OpenCV 3.x C++ API surface that hySLAM's feature path touches.
// TEST INFRASTRUCTURE (oracle/), not product code.
//
// Why it exists: the reference (bmhopkinson/hyslam) links OpenCV 3.4 (CMakeLists.txt:32-35), whose C++ development
// files are not in this image.  To execute the reference's OWN translation units here (oracle/Makefile target
// `_ref/libhyslam_ref.so` compiles them unmodified from /root/reference), this header supplies the cv:: types those
// files use, and cvshim.cpp supplies the handful of OpenCV algorithms they call (FAST, resize, GaussianBlur,
// copyMakeBorder, fastAtan2, small-matrix gemm / norm) -- every one of which is pinned bit-for-bit against the real
// library (cv2 4.13) by tests/test_oracle_vs_cv2.py and tests/test_cvshim_vs_cv2.py.
//
// Semantics that matter for parity and are reproduced on purpose:
//  * cv::Mat is a reference-counted header over shared storage; operator= and copy are shallow; ROIs (operator()(Rect),
//    rowRange, colRange, row, col) alias the parent.
//  * Mat::create() is a no-op when size and type already match -- so `resize(src, roi, ...)`, `x.copyTo(roi)` and
//    `m = Mat::zeros(r, c, t)` write IN PLACE into an existing view (ORBFinder.cpp:71 relies on it).
//  * cvRound = round-half-to-even (cvtss2si / cvtsd2si under the default rounding mode).
//  * float matrix products follow OpenCV's small-matrix gemm: products and sums in fp32, left to right, the `+ C`
//    term of a fused `A*B + C` expression added in double (matmul.cpp); cv::norm / Mat::dot accumulate in double.
#pragma once
#include <algorithm>
#include <cassert>
#include <cfloat>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <limits>
#include <list>
#include <map>
#include <set>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

typedef unsigned char uchar;
typedef unsigned short ushort;

#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_CN_SHIFT 3
#define CV_MAT_DEPTH(t) ((t) & 7)
#define CV_MAT_CN(t) ((((t) >> CV_CN_SHIFT) & 511) + 1)
#define CV_MAKETYPE(depth, cn) (CV_MAT_DEPTH(depth) + (((cn) - 1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_8UC4 CV_MAKETYPE(CV_8U, 4)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_PI 3.1415926535897932384626433832795

static inline int cvRound(double v) { return (int)lrint(v); }
static inline int cvRound(float v) { return (int)lrintf(v); }
static inline int cvRound(int v) { return v; }
static inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
static inline int cvFloor(float v) { int i = (int)v; return i - (i > v); }
static inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }
static inline int cvCeil(float v) { int i = (int)v; return i + (i < v); }

namespace cv {

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    template <typename U> explicit Point_(const Point_<U> &o) : x((T)o.x), y((T)o.y) {}
};
template <typename T> static inline Point_<T> &operator*=(Point_<T> &a, float b) { a.x = (T)(a.x * b); a.y = (T)(a.y * b); return a; }
template <typename T> static inline Point_<T> &operator*=(Point_<T> &a, double b) { a.x = (T)(a.x * b); a.y = (T)(a.y * b); return a; }
template <typename T> static inline Point_<T> &operator*=(Point_<T> &a, int b) { a.x = (T)(a.x * b); a.y = (T)(a.y * b); return a; }
template <typename T> static inline Point_<T> &operator+=(Point_<T> &a, const Point_<T> &b) { a.x += b.x; a.y += b.y; return a; }
template <typename T> static inline Point_<T> &operator-=(Point_<T> &a, const Point_<T> &b) { a.x -= b.x; a.y -= b.y; return a; }
template <typename T> static inline Point_<T> operator+(const Point_<T> &a, const Point_<T> &b) { return Point_<T>(a.x + b.x, a.y + b.y); }
template <typename T> static inline Point_<T> operator-(const Point_<T> &a, const Point_<T> &b) { return Point_<T>(a.x - b.x, a.y - b.y); }
template <typename T> static inline bool operator==(const Point_<T> &a, const Point_<T> &b) { return a.x == b.x && a.y == b.y; }
typedef Point_<int> Point2i;
typedef Point2i Point;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;

template <typename T> struct Point3_ {
    T x, y, z;
    Point3_() : x(0), y(0), z(0) {}
    Point3_(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {}
};
typedef Point3_<float> Point3f;
typedef Point3_<double> Point3d;

template <typename T> struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
    T area() const { return width * height; }
};
typedef Size_<int> Size;

template <typename T> struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T x_, T y_, T w, T h) : x(x_), y(y_), width(w), height(h) {}
};
typedef Rect_<int> Rect;

struct Range {
    int start, end;
    Range() : start(0), end(0) {}
    Range(int s, int e) : start(s), end(e) {}
    static Range all() { return Range(INT32_MIN, INT32_MAX); }
};

template <typename T> struct Scalar_ {
    T val[4];
    Scalar_() { val[0] = val[1] = val[2] = val[3] = 0; }
    Scalar_(T v0, T v1 = 0, T v2 = 0, T v3 = 0) { val[0] = v0; val[1] = v1; val[2] = v2; val[3] = v3; }
    T operator[](int i) const { return val[i]; }
};
typedef Scalar_<double> Scalar;

// 28 bytes, member order of OpenCV's cv::KeyPoint (types.hpp)
class KeyPoint {
public:
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(Point2f pt_, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1)
        : pt(pt_), size(size_), angle(angle_), response(response_), octave(octave_), class_id(class_id_) {}
    KeyPoint(float x, float y, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1)
        : pt(x, y), size(size_), angle(angle_), response(response_), octave(octave_), class_id(class_id_) {}
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
};

enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4,
       BORDER_REFLECT101 = 4, BORDER_DEFAULT = 4, BORDER_ISOLATED = 16 };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1, INTER_CUBIC = 2, INTER_AREA = 3 };
enum { NORM_INF = 1, NORM_L1 = 2, NORM_L2 = 4, NORM_HAMMING = 6 };
enum { DECOMP_LU = 0, DECOMP_SVD = 1 };
enum { COLOR_BGR2GRAY = 6, COLOR_RGB2GRAY = 7, COLOR_BGRA2GRAY = 10, COLOR_RGBA2GRAY = 11 };
#define CV_BGR2GRAY cv::COLOR_BGR2GRAY
#define CV_RGB2GRAY cv::COLOR_RGB2GRAY
#define CV_BGRA2GRAY cv::COLOR_BGRA2GRAY
#define CV_RGBA2GRAY cv::COLOR_RGBA2GRAY

class Mat;
class MatExpr;

// what Mat::zeros / ones / eye return: assigned INTO an existing Mat of matching geometry (MatExpr semantics)
struct MatInit { int rows, cols, type, kind; /* 0 zeros, 1 ones, 2 eye */ };

class Mat {
public:
    int flags;              // low 12 bits: type
    int rows, cols;
    uchar *data;
    size_t step;            // bytes per row (OpenCV: MatStep, convertible to size_t)

    Mat() : flags(0), rows(0), cols(0), data(nullptr), step(0) {}
    Mat(int r, int c, int type) : Mat() { create(r, c, type); }
    Mat(Size sz, int type) : Mat() { create(sz.height, sz.width, type); }
    Mat(int r, int c, int type, const Scalar &s) : Mat() { create(r, c, type); setTo(s); }
    Mat(int r, int c, int type, void *ext, size_t stp = 0) : flags(type & 0xFFF), rows(r), cols(c), data((uchar *)ext), step(stp ? stp : (size_t)c * esz_of(type)) {}
    Mat(const Mat &m, const Rect &roi) : flags(m.flags), rows(roi.height), cols(roi.width), data(m.data + (size_t)roi.y * m.step + (size_t)roi.x * m.elemSize()), step(m.step), store_(m.store_)
    { assert(roi.x >= 0 && roi.y >= 0 && roi.x + roi.width <= m.cols && roi.y + roi.height <= m.rows); }
    Mat(const MatInit &e) : Mat() { *this = e; }
    Mat(const MatExpr &e);

    Mat &operator=(const MatInit &e);
    Mat &operator=(const MatExpr &e);
    Mat &operator=(const Scalar &s) { setTo(s); return *this; }

    static size_t esz1_of(int type) { static const size_t t[8] = {1, 1, 2, 2, 4, 4, 8, 2}; return t[CV_MAT_DEPTH(type)]; }
    static size_t esz_of(int type) { return esz1_of(type) * (size_t)CV_MAT_CN(type); }

    void create(int r, int c, int type)
    {
        type &= 0xFFF;
        if (data && rows == r && cols == c && this->type() == type) return;      // OpenCV: no reallocation
        flags = type; rows = r; cols = c; step = (size_t)c * esz_of(type);
        const size_t bytes = step * (size_t)r;
        store_ = std::shared_ptr<uchar>(bytes ? new uchar[bytes] : nullptr, std::default_delete<uchar[]>());
        data = store_.get();
    }
    void create(Size sz, int type) { create(sz.height, sz.width, type); }
    void release() { store_.reset(); data = nullptr; rows = cols = 0; step = 0; }

    int type() const { return flags & 0xFFF; }
    int depth() const { return CV_MAT_DEPTH(flags); }
    int channels() const { return CV_MAT_CN(flags); }
    size_t elemSize() const { return esz_of(flags); }
    size_t elemSize1() const { return esz1_of(flags); }
    size_t step1() const { return step / elemSize1(); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    size_t total() const { return (size_t)rows * cols; }
    Size size() const { return Size(cols, rows); }
    bool isContinuous() const { return rows <= 1 || step == (size_t)cols * elemSize(); }

    Mat operator()(const Rect &roi) const { return Mat(*this, roi); }
    Mat operator()(Range rr, Range cr) const
    {
        const int r0 = rr.start == INT32_MIN ? 0 : rr.start, r1 = rr.end == INT32_MAX ? rows : rr.end;
        const int c0 = cr.start == INT32_MIN ? 0 : cr.start, c1 = cr.end == INT32_MAX ? cols : cr.end;
        return Mat(*this, Rect(c0, r0, c1 - c0, r1 - r0));
    }
    Mat rowRange(int a, int b) const { return Mat(*this, Rect(0, a, cols, b - a)); }
    Mat colRange(int a, int b) const { return Mat(*this, Rect(a, 0, b - a, rows)); }
    Mat row(int r) const { return rowRange(r, r + 1); }
    Mat col(int c) const { return colRange(c, c + 1); }

    Mat clone() const
    {
        Mat m;
        if (empty()) return m;
        m.create(rows, cols, type());
        const size_t rb = (size_t)cols * elemSize();
        for (int r = 0; r < rows; r++) std::memcpy(m.data + (size_t)r * m.step, data + (size_t)r * step, rb);
        return m;
    }
    void copyTo(const class _OutputArray &dst) const;
    void convertTo(const class _OutputArray &dst, int rtype, double alpha = 1, double beta = 0) const;
    Mat &setTo(const Scalar &s);
    void resize(size_t nrows);                    // keeps the leading rows (Camera.cpp:83)
    void push_back(const Mat &m);                 // appends rows
    Mat reshape(int cn, int nrows = 0) const;

    template <typename T> T &at(int r, int c) { return *(T *)(data + (size_t)r * step + (size_t)c * sizeof(T)); }
    template <typename T> const T &at(int r, int c) const { return *(const T *)(data + (size_t)r * step + (size_t)c * sizeof(T)); }
    // single index: element i of a row or column vector (OpenCV mat.inl.hpp at(int i0))
    template <typename T> T &at(int i) { return rows == 1 ? *(T *)(data + (size_t)i * sizeof(T)) : (cols == 1 ? *(T *)(data + (size_t)i * step) : at<T>(i / cols, i % cols)); }
    template <typename T> const T &at(int i) const { return const_cast<Mat *>(this)->at<T>(i); }
    uchar *ptr(int r = 0) { return data + (size_t)r * step; }
    const uchar *ptr(int r = 0) const { return data + (size_t)r * step; }
    template <typename T> T *ptr(int r = 0) { return (T *)(data + (size_t)r * step); }
    template <typename T> const T *ptr(int r = 0) const { return (const T *)(data + (size_t)r * step); }

    static MatInit zeros(int r, int c, int type) { return MatInit{r, c, type, 0}; }
    static MatInit zeros(Size s, int type) { return MatInit{s.height, s.width, type, 0}; }
    static MatInit ones(int r, int c, int type) { return MatInit{r, c, type, 1}; }
    static MatInit eye(int r, int c, int type) { return MatInit{r, c, type, 2}; }

    // float algebra (cvshim.cpp); A*B is lazy so that `A*B + C` is ONE gemm like OpenCV's MatExpr
    MatExpr t() const;
    Mat inv(int method = DECOMP_LU) const;
    double dot(const Mat &m) const;
    Mat cross(const Mat &m) const;
    Mat mul(const Mat &m, double scale = 1) const;

private:
    std::shared_ptr<uchar> store_;
};

template <typename T> struct DataDepth;
template <> struct DataDepth<uchar> { enum { value = CV_8U }; };
template <> struct DataDepth<int> { enum { value = CV_32S }; };
template <> struct DataDepth<float> { enum { value = CV_32F }; };
template <> struct DataDepth<double> { enum { value = CV_64F }; };

template <typename T> class Mat_ : public Mat {
public:
    Mat_() : Mat() {}
    Mat_(int r, int c) : Mat(r, c, DataDepth<T>::value) {}
    Mat_(const Mat &m) : Mat(m) { assert(m.empty() || m.type() == DataDepth<T>::value); }
    T &operator()(int r, int c) { return this->template at<T>(r, c); }
    const T &operator()(int r, int c) const { return this->template at<T>(r, c); }
    T &operator()(int i) { return this->template at<T>(i); }
};
// `Mat_<float>(3,1) << x, y, z` (Camera.cpp:158)
template <typename T> class MatCommaInitializer_ {
public:
    explicit MatCommaInitializer_(Mat_<T> *m) : m_(m), i_(0) {}
    template <typename U> MatCommaInitializer_ &operator,(U v) { put((T)v); return *this; }
    void put(T v) { m_->template at<T>(i_ / m_->cols, i_ % m_->cols) = v; i_++; }
    operator Mat_<T>() const { return *m_; }
    operator Mat() const { return *m_; }
private:
    Mat_<T> *m_; int i_;
};
template <typename T, typename U> static inline MatCommaInitializer_<T> operator<<(const Mat_<T> &m, U v)
{
    MatCommaInitializer_<T> ci(const_cast<Mat_<T> *>(&m));
    ci.put((T)v);
    return ci;
}

// lazy matrix expression: alpha * op(A) * B + beta * C   |   or an already evaluated Mat
class MatExpr {
public:
    enum Kind { VALUE, GEMM, GEMM_TA, TRANSPOSE, SCALED };   // GEMM_TA: op(A) = A^T; TRANSPOSE / SCALED: alpha * A^T, alpha * A
    Kind kind; Mat a, b, c; double alpha, beta;
    MatExpr() : kind(VALUE), alpha(1), beta(0) {}
    MatExpr(const Mat &m) : kind(VALUE), a(m), alpha(1), beta(0) {}
    static MatExpr gemm(const Mat &A, const Mat &B, double alpha_, const Mat &C, double beta_) { MatExpr e; e.kind = GEMM; e.a = A; e.b = B; e.c = C; e.alpha = alpha_; e.beta = beta_; return e; }
    Mat eval() const;
    operator Mat() const { return eval(); }
    MatExpr t() const { return eval().t(); }
    Mat inv(int method = DECOMP_LU) const { return eval().inv(method); }
    Mat clone() const { return eval().clone(); }
    Mat rowRange(int a_, int b_) const { return eval().rowRange(a_, b_); }
    Mat colRange(int a_, int b_) const { return eval().colRange(a_, b_); }
    Mat row(int r) const { return eval().row(r); }
    Mat col(int r) const { return eval().col(r); }
    double dot(const Mat &m) const { return eval().dot(m); }
    template <typename T> T at(int r, int c) const { return eval().at<T>(r, c); }
    template <typename T> T at(int i) const { return eval().at<T>(i); }
};
inline Mat::Mat(const MatExpr &e) : Mat() { *this = e.eval(); }

MatExpr operator*(const Mat &a, const Mat &b);
MatExpr operator*(const MatExpr &a, const Mat &b);
MatExpr operator*(const Mat &a, const MatExpr &b);
MatExpr operator*(const MatExpr &a, const MatExpr &b);
MatExpr operator*(const Mat &a, double s);
MatExpr operator*(double s, const Mat &a);
MatExpr operator*(const MatExpr &a, double s);
MatExpr operator*(double s, const MatExpr &a);
MatExpr operator/(const Mat &a, double s);
MatExpr operator/(const MatExpr &a, double s);
MatExpr operator+(const Mat &a, const Mat &b);
MatExpr operator+(const MatExpr &a, const Mat &b);
MatExpr operator+(const Mat &a, const MatExpr &b);
MatExpr operator+(const MatExpr &a, const MatExpr &b);
MatExpr operator-(const Mat &a, const Mat &b);
MatExpr operator-(const MatExpr &a, const Mat &b);
MatExpr operator-(const Mat &a, const MatExpr &b);
MatExpr operator-(const MatExpr &a, const MatExpr &b);
MatExpr operator-(const Mat &a);
MatExpr operator-(const MatExpr &a);
std::ostream &operator<<(std::ostream &os, const Mat &m);

class _InputArray {            // cv::InputArray = const _InputArray&
public:
    _InputArray() {}
    _InputArray(const Mat &m) : m_(m) {}
    _InputArray(const MatExpr &e) : m_(e.eval()) {}
    Mat getMat() const { return m_; }
    bool empty() const { return m_.empty(); }
    Size size() const { return m_.size(); }
    int type() const { return m_.type(); }
protected:
    Mat m_;
};
typedef const _InputArray &InputArray;

class _OutputArray {           // cv::OutputArray = const _OutputArray&; binds lvalue Mats and temporary views
public:
    _OutputArray(Mat &m) : p_(&m) {}
    _OutputArray(const Mat &m) : tmp_(m), p_(&tmp_) {}     // temporary header: shares storage, so in-place writes land
    Mat &ref() const { return *p_; }
    void create(int r, int c, int type) const { p_->create(r, c, type); }
private:
    mutable Mat tmp_;
    Mat *p_;
};
typedef const _OutputArray &OutputArray;
typedef const _OutputArray &InputOutputArray;

// ---- the algorithms the reference calls (cvshim.cpp) ----
float fastAtan2(float y, float x);
double norm(InputArray a, int normType = NORM_L2);
double norm(InputArray a, InputArray b, int normType = NORM_L2);
void FAST(InputArray image, std::vector<KeyPoint> &keypoints, int threshold, bool nonmaxSuppression = true);
void resize(InputArray src, OutputArray dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR);
void GaussianBlur(InputArray src, OutputArray dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_DEFAULT);
void copyMakeBorder(InputArray src, OutputArray dst, int top, int bottom, int left, int right, int borderType, const Scalar &value = Scalar());
void cvtColor(InputArray src, OutputArray dst, int code, int dstCn = 0);

// ---- configuration I/O: declared so that headers parse; the path never reads a file ----
class FileNode {
public:
    FileNode() {}
    FileNode operator[](const char *) const { return FileNode(); }
    FileNode operator[](const std::string &) const { return FileNode(); }
    std::string name() const { return std::string(); }
    std::string string() const { return std::string(); }
    bool empty() const { return true; }
    operator int() const { return 0; }
    operator float() const { return 0.f; }
    operator double() const { return 0.0; }
    operator std::string() const { return std::string(); }
};
class FileStorage {
public:
    enum { READ = 0, WRITE = 1 };
    FileStorage() {}
    FileStorage(const std::string &, int) {}
    bool isOpened() const { return false; }
    FileNode operator[](const char *) const { return FileNode(); }
    FileNode operator[](const std::string &) const { return FileNode(); }
    void release() {}
};
template <typename T> static inline void operator>>(const FileNode &n, T &v) { v = (T)n; }

}  // namespace cv
