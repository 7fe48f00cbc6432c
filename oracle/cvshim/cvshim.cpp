// cvshim.cpp -- the OpenCV algorithms hySLAM's feature path calls, restated for oracle/_ref.
// TEST INFRASTRUCTURE (oracle/), not product code.  See opencv2/core/core.hpp for why this exists.
//
// Each routine follows the algorithm OpenCV 3.4 publishes for 8-bit single-channel input and is pinned bit-for-bit
// against the real library (cv2 4.13) by tests/test_cvshim_vs_cv2.py:
//   cv::FAST            features2d/fast.cpp FAST_t<16> + fast_score.cpp cornerScore<16>   (call site ORBFinder.cpp:67)
//   cv::resize          imgproc/resize.cpp  HResizeLinear + VResizeLinear<uchar,int,short>  (ORBExtractor.cpp:577)
//   cv::GaussianBlur    imgproc/smooth.cpp  fixed-point separable path                       (ORBExtractor.cpp:537)
//   cv::copyMakeBorder  core/copy.cpp                                                        (ORBExtractor.cpp:579-585)
//   cv::fastAtan2       core/mathfuncs_core atan_f32 scalar path                             (ORBFinder.cpp:42)
//   gemm / norm / dot   core/matmul.cpp small-matrix path, double accumulation elsewhere
//
// These are written to be FAST as well as exact (the CPU baseline of bench.py runs the reference's code on top of them):
// the corner test and score are evaluated densely on byte vectors (GCC vector extensions, AVX2 clone when the host has it),
// which is quicker than OpenCV's own early-out + scalar cornerScore on corner-rich frames.
#include <opencv2/core/core.hpp>
#include <algorithm>
#include <cfloat>
#include <cstdlib>

namespace cv {

// ------------------------------------------------------------------------------------------------------------------
// Mat plumbing
// ------------------------------------------------------------------------------------------------------------------
Mat &Mat::setTo(const Scalar &s)
{
    for (int r = 0; r < rows; r++) {
        uchar *p = data + (size_t)r * step;
        const int n = cols * channels();
        switch (depth()) {
        case CV_8U: for (int i = 0; i < n; i++) p[i] = (uchar)std::min(255.0, std::max(0.0, std::nearbyint(s.val[i % channels()]))); break;
        case CV_32S: for (int i = 0; i < n; i++) ((int *)p)[i] = (int)s.val[i % channels()]; break;
        case CV_32F: for (int i = 0; i < n; i++) ((float *)p)[i] = (float)s.val[i % channels()]; break;
        case CV_64F: for (int i = 0; i < n; i++) ((double *)p)[i] = s.val[i % channels()]; break;
        default: std::abort();
        }
    }
    return *this;
}

Mat &Mat::operator=(const MatInit &e)
{
    create(e.rows, e.cols, e.type);            // no-op on a matching view: the fill below lands in the parent's storage
    setTo(Scalar(e.kind == 1 ? 1.0 : 0.0));
    if (e.kind == 2) {
        const int n = std::min(rows, cols);
        for (int i = 0; i < n; i++) {
            switch (depth()) {
            case CV_8U: at<uchar>(i, i) = 1; break;
            case CV_32S: at<int>(i, i) = 1; break;
            case CV_32F: at<float>(i, i) = 1.f; break;
            case CV_64F: at<double>(i, i) = 1.0; break;
            default: std::abort();
            }
        }
    }
    return *this;
}

Mat &Mat::operator=(const MatExpr &e) { *this = e.eval(); return *this; }

void Mat::copyTo(const _OutputArray &out) const
{
    Mat &dst = out.ref();
    if (empty()) { dst.release(); return; }
    dst.create(rows, cols, type());
    if (dst.data == data && dst.step == step) return;
    const size_t rb = (size_t)cols * elemSize();
    for (int r = 0; r < rows; r++) std::memmove(dst.data + (size_t)r * dst.step, data + (size_t)r * step, rb);
}

static double get_elem(const Mat &m, int r, int c)
{
    switch (m.depth()) {
    case CV_8U: return m.at<uchar>(r, c);
    case CV_32S: return m.at<int>(r, c);
    case CV_32F: return m.at<float>(r, c);
    case CV_64F: return m.at<double>(r, c);
    default: std::abort();
    }
}

void Mat::convertTo(const _OutputArray &out, int rtype, double alpha, double beta) const
{
    Mat src = *this;                            // `m.convertTo(m, ...)`: keep the source alive
    Mat d;
    d.create(rows, cols, CV_MAKETYPE(CV_MAT_DEPTH(rtype), 1));
    assert(channels() == 1);
    for (int r = 0; r < rows; r++)
        for (int c = 0; c < cols; c++) {
            const double v = get_elem(src, r, c) * alpha + beta;
            switch (d.depth()) {
            case CV_8U: d.at<uchar>(r, c) = (uchar)std::min(255, std::max(0, cvRound(v))); break;
            case CV_32S: d.at<int>(r, c) = cvRound(v); break;
            case CV_32F: d.at<float>(r, c) = (float)v; break;
            case CV_64F: d.at<double>(r, c) = v; break;
            default: std::abort();
            }
        }
    d.copyTo(out);
}

void Mat::resize(size_t nrows)
{
    Mat m(int(nrows), cols, type());
    const int keep = std::min((int)nrows, rows);
    for (int r = 0; r < keep; r++) std::memcpy(m.ptr(r), ptr(r), (size_t)cols * elemSize());
    for (int r = keep; r < (int)nrows; r++) std::memset(m.ptr(r), 0, (size_t)cols * elemSize());
    *this = m;
}

void Mat::push_back(const Mat &m)
{
    if (empty()) { *this = m.clone(); return; }
    assert(m.cols == cols && m.type() == type());
    Mat out(rows + m.rows, cols, type());
    for (int r = 0; r < rows; r++) std::memcpy(out.ptr(r), ptr(r), (size_t)cols * elemSize());
    for (int r = 0; r < m.rows; r++) std::memcpy(out.ptr(rows + r), m.ptr(r), (size_t)cols * elemSize());
    *this = out;
}

Mat Mat::reshape(int cn, int nrows) const
{
    assert((cn == 0 || cn == channels()) && isContinuous());
    Mat m = *this;
    if (nrows > 0 && nrows != rows) { m.cols = (int)(total() / nrows); m.rows = nrows; m.step = (size_t)m.cols * elemSize(); }
    return m;
}

std::ostream &operator<<(std::ostream &os, const Mat &m)
{
    os << "[";
    for (int r = 0; r < m.rows; r++) {
        for (int c = 0; c < m.cols; c++) os << get_elem(m, r, c) << (c + 1 < m.cols ? ", " : "");
        os << (r + 1 < m.rows ? ";\n " : "");
    }
    return os << "]";
}

// ------------------------------------------------------------------------------------------------------------------
// float / double algebra.  gemm: D = alpha * op(A) * B + beta * C.
// core/matmul.cpp: with no transposition flag and an inner length of 2..4 that equals a dimension of D, the CV_32F case
// forms each product sum in fp32, left to right, then `(float)(t * alpha + c * beta)` with alpha / beta in double.
// Everything else goes through GEMMSingleMul<float, double>: accumulation in double.
// ------------------------------------------------------------------------------------------------------------------
static Mat gemm_eval(const Mat &A0, bool ta, const Mat &B, double alpha, const Mat &C, double beta)
{
    assert(A0.depth() == B.depth() && (A0.depth() == CV_32F || A0.depth() == CV_64F) && A0.channels() == 1);
    const int M = ta ? A0.cols : A0.rows, K = ta ? A0.rows : A0.cols, N = B.cols;
    assert(B.rows == K);
    const bool haveC = !C.empty() && beta != 0;
    if (haveC) assert(C.rows == M && C.cols == N && C.depth() == A0.depth());
    Mat D(M, N, A0.type());
    if (A0.depth() == CV_32F) {
        const bool small = !ta && K >= 2 && K <= 4 && (K == N || K == M);
        for (int i = 0; i < M; i++)
            for (int j = 0; j < N; j++) {
                const float c = haveC ? C.at<float>(i, j) : 0.f;
                if (small) {
                    float t = A0.at<float>(i, 0) * B.at<float>(0, j);
                    for (int k = 1; k < K; k++) t = t + A0.at<float>(i, k) * B.at<float>(k, j);
                    D.at<float>(i, j) = (float)(t * alpha + c * (haveC ? beta : 0.0));
                } else {
                    double s = 0;
                    for (int k = 0; k < K; k++) s += (double)(ta ? A0.at<float>(k, i) : A0.at<float>(i, k)) * (double)B.at<float>(k, j);
                    D.at<float>(i, j) = (float)(s * alpha + c * (haveC ? beta : 0.0));
                }
            }
    } else {
        for (int i = 0; i < M; i++)
            for (int j = 0; j < N; j++) {
                double s = 0;
                for (int k = 0; k < K; k++) s += (ta ? A0.at<double>(k, i) : A0.at<double>(i, k)) * B.at<double>(k, j);
                D.at<double>(i, j) = s * alpha + (haveC ? C.at<double>(i, j) * beta : 0.0);
            }
    }
    return D;
}

static Mat transpose_eval(const Mat &a)
{
    Mat d(a.cols, a.rows, a.type());
    const size_t es = a.elemSize();
    for (int r = 0; r < a.rows; r++)
        for (int c = 0; c < a.cols; c++) std::memcpy(d.data + (size_t)c * d.step + (size_t)r * es, a.data + (size_t)r * a.step + (size_t)c * es, es);
    return d;
}

// element-wise a*sa + b*sb.  CV_32F follows OpenCV's work type for float data: a scaled Mat (`M * s`, `M / s`, `s * M.t()`) is
// Mat::convertTo(..., alpha) = cvtScale_<float, float, float>: src * (float)alpha in FLOAT arithmetic; plain sums and
// differences are cv::add / cv::subtract (exact float ops); a general two-term combination is addWeighted in float.
static Mat lincomb(const Mat &a, double sa, const Mat &b, double sb)
{
    assert(b.empty() || (a.rows == b.rows && a.cols == b.cols && a.type() == b.type()));
    Mat d(a.rows, a.cols, a.type());
    const bool plain = (sa == 1 || sa == -1) && (sb == 1 || sb == -1 || b.empty());
    const float fa = (float)sa, fb = (float)sb;
    for (int r = 0; r < a.rows; r++)
        for (int c = 0; c < a.cols; c++) {
            if (a.depth() == CV_32F) {
                const float x = a.at<float>(r, c), y = b.empty() ? 0.f : b.at<float>(r, c);
                if (b.empty()) d.at<float>(r, c) = plain ? (sa < 0 ? -x : x) : x * fa;
                else if (plain) d.at<float>(r, c) = (sa < 0 ? -x : x) + (sb < 0 ? -y : y);
                else d.at<float>(r, c) = x * fa + y * fb;
            } else if (a.depth() == CV_64F) {
                const double x = a.at<double>(r, c), y = b.empty() ? 0.0 : b.at<double>(r, c);
                d.at<double>(r, c) = x * sa + y * sb;
            } else std::abort();
        }
    return d;
}

Mat MatExpr::eval() const
{
    if (kind == VALUE) return a;
    if (kind == GEMM) return gemm_eval(a, false, b, alpha, c, beta);
    if (kind == GEMM_TA) return gemm_eval(a, true, b, alpha, c, beta);
    if (kind == TRANSPOSE) return alpha == 1 ? transpose_eval(a) : lincomb(transpose_eval(a), alpha, Mat(), 0);
    if (kind == SCALED) return lincomb(a, alpha, Mat(), 0);
    std::abort();
}

MatExpr Mat::t() const { MatExpr e; e.kind = MatExpr::TRANSPOSE; e.a = *this; return e; }

static MatExpr mul_expr(const MatExpr &x, const MatExpr &y)
{
    // (s * A^T) * B and (s * A) * B stay one gemm; anything else is evaluated first
    MatExpr e; e.beta = 0;
    const Mat B = y.kind == MatExpr::SCALED ? y.a : y.eval();
    const double sy = y.kind == MatExpr::SCALED ? y.alpha : 1.0;
    if (x.kind == MatExpr::TRANSPOSE) { e.kind = MatExpr::GEMM_TA; e.a = x.a; e.alpha = x.alpha * sy; }
    else if (x.kind == MatExpr::SCALED) { e.kind = MatExpr::GEMM; e.a = x.a; e.alpha = x.alpha * sy; }
    else { e.kind = MatExpr::GEMM; e.a = x.eval(); e.alpha = sy; }
    e.b = B;
    return e;
}
static MatExpr scale_expr(const MatExpr &x, double s)
{
    MatExpr e = x;
    if (x.kind == MatExpr::VALUE) { e.kind = MatExpr::SCALED; e.alpha = s; return e; }
    if (x.kind == MatExpr::GEMM || x.kind == MatExpr::GEMM_TA) { e.alpha *= s; e.beta *= s; return e; }
    if (x.kind == MatExpr::TRANSPOSE || x.kind == MatExpr::SCALED) { e.alpha *= s; return e; }
    std::abort();
}
static MatExpr add_expr(const MatExpr &x, const MatExpr &y, double sy)
{
    // gemm + C (either side): fold into the gemm when it has no C yet
    if ((x.kind == MatExpr::GEMM || x.kind == MatExpr::GEMM_TA) && x.c.empty() && y.kind == MatExpr::VALUE) { MatExpr e = x; e.c = y.a; e.beta = sy; return e; }
    if ((y.kind == MatExpr::GEMM || y.kind == MatExpr::GEMM_TA) && y.c.empty() && x.kind == MatExpr::VALUE && sy == 1) { MatExpr e = y; e.c = x.a; e.beta = 1; return e; }
    const double sx = x.kind == MatExpr::SCALED ? x.alpha : 1.0;
    const double syy = (y.kind == MatExpr::SCALED ? y.alpha : 1.0) * sy;
    return MatExpr(lincomb(x.kind == MatExpr::SCALED ? x.a : x.eval(), sx, y.kind == MatExpr::SCALED ? y.a : y.eval(), syy));
}

MatExpr operator*(const Mat &a, const Mat &b) { return mul_expr(MatExpr(a), MatExpr(b)); }
MatExpr operator*(const MatExpr &a, const Mat &b) { return mul_expr(a, MatExpr(b)); }
MatExpr operator*(const Mat &a, const MatExpr &b) { return mul_expr(MatExpr(a), b); }
MatExpr operator*(const MatExpr &a, const MatExpr &b) { return mul_expr(a, b); }
MatExpr operator*(const Mat &a, double s) { return scale_expr(MatExpr(a), s); }
MatExpr operator*(double s, const Mat &a) { return scale_expr(MatExpr(a), s); }
MatExpr operator*(const MatExpr &a, double s) { return scale_expr(a, s); }
MatExpr operator*(double s, const MatExpr &a) { return scale_expr(a, s); }
MatExpr operator/(const Mat &a, double s) { return scale_expr(MatExpr(a), 1.0 / s); }
MatExpr operator/(const MatExpr &a, double s) { return scale_expr(a, 1.0 / s); }
MatExpr operator+(const Mat &a, const Mat &b) { return add_expr(MatExpr(a), MatExpr(b), 1); }
MatExpr operator+(const MatExpr &a, const Mat &b) { return add_expr(a, MatExpr(b), 1); }
MatExpr operator+(const Mat &a, const MatExpr &b) { return add_expr(MatExpr(a), b, 1); }
MatExpr operator+(const MatExpr &a, const MatExpr &b) { return add_expr(a, b, 1); }
MatExpr operator-(const Mat &a, const Mat &b) { return add_expr(MatExpr(a), MatExpr(b), -1); }
MatExpr operator-(const MatExpr &a, const Mat &b) { return add_expr(a, MatExpr(b), -1); }
MatExpr operator-(const Mat &a, const MatExpr &b) { return add_expr(MatExpr(a), b, -1); }
MatExpr operator-(const MatExpr &a, const MatExpr &b) { return add_expr(a, b, -1); }
MatExpr operator-(const Mat &a) { return scale_expr(MatExpr(a), -1); }
MatExpr operator-(const MatExpr &a) { return scale_expr(a, -1); }

double Mat::dot(const Mat &m) const
{
    assert(total() == m.total() && type() == m.type());
    double s = 0;
    const int R = rows, Cc = cols;
    for (int r = 0; r < R; r++)
        for (int c = 0; c < Cc; c++) {
            const int i = r * Cc + c;
            s += get_elem(*this, r, c) * get_elem(m, i / m.cols, i % m.cols);
        }
    return s;
}

Mat Mat::cross(const Mat &m) const
{
    assert(total() == 3 && m.total() == 3 && depth() == CV_32F);
    const float a0 = at<float>(0), a1 = at<float>(1), a2 = at<float>(2), b0 = m.at<float>(0), b1 = m.at<float>(1), b2 = m.at<float>(2);
    Mat d(rows, cols, type());
    d.at<float>(0) = a1 * b2 - a2 * b1; d.at<float>(1) = a2 * b0 - a0 * b2; d.at<float>(2) = a0 * b1 - a1 * b0;
    return d;
}

Mat Mat::mul(const Mat &m, double scale) const
{
    Mat d(rows, cols, type());
    for (int r = 0; r < rows; r++)
        for (int c = 0; c < cols; c++) {
            if (depth() == CV_32F) d.at<float>(r, c) = scale == 1 ? at<float>(r, c) * m.at<float>(r, c) : (float)(scale * at<float>(r, c) * m.at<float>(r, c));
            else if (depth() == CV_64F) d.at<double>(r, c) = scale * at<double>(r, c) * m.at<double>(r, c);
            else std::abort();
        }
    return d;
}

Mat Mat::inv(int) const
{
    // Gauss-Jordan with partial pivoting in double (used for 3x3 / 4x4 pose matrices only; not on the parity path)
    assert(rows == cols && (depth() == CV_32F || depth() == CV_64F));
    const int n = rows;
    std::vector<double> a((size_t)n * 2 * n, 0.0);
    for (int r = 0; r < n; r++) { for (int c = 0; c < n; c++) a[(size_t)r * 2 * n + c] = get_elem(*this, r, c); a[(size_t)r * 2 * n + n + r] = 1.0; }
    for (int i = 0; i < n; i++) {
        int p = i;
        for (int r = i + 1; r < n; r++) if (std::fabs(a[(size_t)r * 2 * n + i]) > std::fabs(a[(size_t)p * 2 * n + i])) p = r;
        if (a[(size_t)p * 2 * n + i] == 0) return Mat(MatInit{n, n, type(), 0});
        if (p != i) for (int c = 0; c < 2 * n; c++) std::swap(a[(size_t)p * 2 * n + c], a[(size_t)i * 2 * n + c]);
        const double piv = a[(size_t)i * 2 * n + i];
        for (int c = 0; c < 2 * n; c++) a[(size_t)i * 2 * n + c] /= piv;
        for (int r = 0; r < n; r++) if (r != i) { const double f = a[(size_t)r * 2 * n + i]; if (f != 0) for (int c = 0; c < 2 * n; c++) a[(size_t)r * 2 * n + c] -= f * a[(size_t)i * 2 * n + c]; }
    }
    Mat d(n, n, type());
    for (int r = 0; r < n; r++) for (int c = 0; c < n; c++) { if (depth() == CV_32F) d.at<float>(r, c) = (float)a[(size_t)r * 2 * n + n + c]; else d.at<double>(r, c) = a[(size_t)r * 2 * n + n + c]; }
    return d;
}

// cv::norm: accumulates in double for every depth
double norm(InputArray a_, int normType)
{
    const Mat a = a_.getMat();
    double s = 0;
    for (int r = 0; r < a.rows; r++)
        for (int c = 0; c < a.cols; c++) {
            const double v = get_elem(a, r, c);
            if (normType == NORM_L2) s += v * v; else if (normType == NORM_L1) s += std::fabs(v); else s = std::max(s, std::fabs(v));
        }
    return normType == NORM_L2 ? std::sqrt(s) : s;
}
double norm(InputArray a_, InputArray b_, int normType)
{
    const Mat a = a_.getMat(), b = b_.getMat();
    assert(a.rows == b.rows && a.cols == b.cols && a.type() == b.type());
    double s = 0;
    for (int r = 0; r < a.rows; r++)
        for (int c = 0; c < a.cols; c++) {
            const double v = get_elem(a, r, c) - get_elem(b, r, c);
            if (normType == NORM_L2) s += v * v; else if (normType == NORM_L1) s += std::fabs(v); else s = std::max(s, std::fabs(v));
        }
    return normType == NORM_L2 ? std::sqrt(s) : s;
}

// ------------------------------------------------------------------------------------------------------------------
// cv::fastAtan2 (degrees) -- every operation rounded to fp32, no FMA (this file is compiled with -ffp-contract=off)
// ------------------------------------------------------------------------------------------------------------------
float fastAtan2(float y, float x)
{
    const float k = (float)(180.0 / CV_PI);
    const float p1 = 0.9997878412794807f * k, p3 = -0.3258083974640975f * k, p5 = 0.1555786518463281f * k, p7 = -0.04432655554792128f * k;
    const float ax = std::fabs(x), ay = std::fabs(y);
    float a, c, c2;
    if (ax >= ay) { c = ay / (ax + (float)DBL_EPSILON); c2 = c * c; a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c; }
    else { c = ax / (ay + (float)DBL_EPSILON); c2 = c * c; a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c; }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

// ------------------------------------------------------------------------------------------------------------------
// cv::FAST, FAST-9/16 with optional 3x3 non-maximum suppression, on the ROI a Mat header describes.
// corner  <=>  S > threshold, where S = max over both polarities and the 16 contiguous 9-arcs of the minimum difference
// along the arc; response = S - 1 (cornerScore<16> starts at `threshold` and returns the last value that still passes,
// minus one); NMS keeps a corner whose response is strictly above its 8 neighbours' (non-corners count as 0); pixels
// within 3 of the ROI edge are never reported; output is row-major.  KeyPoint(x, y, 7.f, -1, response).
// ------------------------------------------------------------------------------------------------------------------
typedef uint8_t vu8 __attribute__((vector_size(32), aligned(1), may_alias));
static const int ring_dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int ring_dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

static inline vu8 vmin(vu8 a, vu8 b) { return a < b ? a : b; }
static inline vu8 vmax(vu8 a, vu8 b) { return a > b ? a : b; }

// S (saturated to 0 when no arc is one-sided) for 32 consecutive pixels starting at p
__attribute__((target_clones("avx2", "default")))
static void fast_score_row(const uint8_t *row, int stride, int x0, int x1, uint8_t *out)
{
    for (int x = x0; x < x1; x += 32) {
        const uint8_t *p = row + x;
        const vu8 v = *(const vu8 *)p;
        vu8 d[16], e[16];
        for (int k = 0; k < 16; k++) {
            const vu8 q = *(const vu8 *)(p + ring_dy[k] * stride + ring_dx[k]);
            d[k] = vmax(q, v) - v;             // ring - centre, saturated at 0
            e[k] = vmax(v, q) - q;             // centre - ring, saturated at 0
        }
        vu8 best = {0};
        for (int pol = 0; pol < 2; pol++) {
            vu8 *a = pol ? e : d;
            vu8 m2[16], m4[16], m8[16];
            for (int k = 0; k < 16; k++) m2[k] = vmin(a[k], a[(k + 1) & 15]);
            for (int k = 0; k < 16; k++) m4[k] = vmin(m2[k], m2[(k + 2) & 15]);
            for (int k = 0; k < 16; k++) m8[k] = vmin(m4[k], m4[(k + 4) & 15]);
            for (int k = 0; k < 16; k++) best = vmax(best, vmin(m8[k], a[(k + 8) & 15]));
        }
        *(vu8 *)(out + x) = best;
    }
}

void FAST(InputArray image_, std::vector<KeyPoint> &keypoints, int threshold, bool nms)
{
    const Mat img = image_.getMat();
    keypoints.clear();
    assert(img.type() == CV_8UC1);
    const int w = img.cols, h = img.rows;
    if (w < 7 || h < 7) return;
    // The ROI is copied into a padded scratch (row pitch and tail sized for whole 32-byte vectors): vector loads may not
    // run past the parent image, and hySLAM's cells are ~36 px wide.
    const int pitch = ((w + 31) & ~31) + 64;
    // malloc, not operator new: the buffers outlive the call (oracle/_ref may run calls under a scoped arena allocator)
    struct Scratch { uint8_t *p = nullptr; size_t n = 0; ~Scratch() { std::free(p); }
                     uint8_t *get(size_t need) { if (need > n) { std::free(p); p = (uint8_t *)std::malloc(need); n = need; if (!p) std::abort(); } return p; } };
    thread_local Scratch scratch_buf, score_buf;
    const size_t bytes = (size_t)pitch * (h + 1) + 64;
    uint8_t *scratch = scratch_buf.get(bytes), *score = score_buf.get(bytes);
    std::memset(score, 0, bytes);
    for (int y = 0; y < h; y++) { uint8_t *row = &scratch[(size_t)y * pitch]; std::memset(row, 0, 32); std::memcpy(row + 32, img.ptr(y), (size_t)w); std::memset(row + 32 + w, 0, (size_t)pitch - 32 - w); }
    for (int y = 3; y < h - 3; y++) fast_score_row(&scratch[(size_t)y * pitch + 32], pitch, 3, w - 3, &score[(size_t)y * pitch + 32]);
    for (int y = 3; y < h - 3; y++) {
        uint8_t *s = &score[(size_t)y * pitch + 32];
        for (int x = 0; x < 3; x++) s[x] = 0;
        for (int x = w - 3; x < pitch - 32; x++) s[x] = 0;       // vector tail wrote past the detect region
        for (int x = 3; x < w - 3; x++) s[x] = s[x] > threshold ? (uint8_t)(s[x] - 1) : 0;   // response, 0 = no corner
    }
    for (int y = 3; y < h - 3; y++) {
        const uint8_t *s = &score[(size_t)y * pitch + 32];
        for (int x = 3; x < w - 3; x++) {
            const int v = s[x];
            if (!v) continue;
            if (nms) {
                const uint8_t *u = s - pitch, *d = s + pitch;
                if (!(v > s[x - 1] && v > s[x + 1] && v > u[x - 1] && v > u[x] && v > u[x + 1] && v > d[x - 1] && v > d[x] && v > d[x + 1])) continue;
            }
            keypoints.push_back(KeyPoint((float)x, (float)y, 7.f, -1, (float)v));
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// cv::resize, INTER_LINEAR, 8UC1.  Q11 coefficients; vertical combine ((b*(h>>4))>>16 ... +2)>>2; an exact 2:1
// reduction silently takes OpenCV's INTER_AREA 2x2 fast path.
// ------------------------------------------------------------------------------------------------------------------
static void linear_coeffs(int dst_n, int src_n, std::vector<int> &ofs, std::vector<short> &c0, std::vector<short> &c1, bool vertical)
{
    const double inv_scale = (double)dst_n / src_n, scale = 1.0 / inv_scale;
    ofs.resize(dst_n); c0.resize(dst_n); c1.resize(dst_n);
    for (int d = 0; d < dst_n; d++) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = cvFloor(f);
        f -= s;
        if (!vertical) {                        // horizontal taps outside the row drop the fraction; vertical rows are clipped instead
            if (s < 0) { s = 0; f = 0.f; }
            if (s >= src_n - 1) { s = src_n - 1; f = 0.f; }
        }
        ofs[d] = s;
        c0[d] = (short)cvRound((1.f - f) * 2048.f);
        c1[d] = (short)cvRound(f * 2048.f);
    }
}

void resize(InputArray src_, OutputArray dst_, Size dsize, double fx, double fy, int interpolation)
{
    const Mat src = src_.getMat();
    assert(src.type() == CV_8UC1 && interpolation == INTER_LINEAR);
    if (dsize.width == 0 || dsize.height == 0) dsize = Size(cvRound(src.cols * fx), cvRound(src.rows * fy));
    dst_.create(dsize.height, dsize.width, src.type());
    Mat dst = dst_.ref();
    const int sw = src.cols, sh = src.rows, dw = dst.cols, dh = dst.rows;
    if (sw == dw && sh == dh) { src.copyTo(dst); return; }
    if (sw == 2 * dw && sh == 2 * dh) {
        for (int y = 0; y < dh; y++) {
            const uchar *p = src.ptr(2 * y), *q = src.ptr(2 * y + 1);
            uchar *o = dst.ptr(y);
            for (int x = 0; x < dw; x++) o[x] = (uchar)((p[2 * x] + p[2 * x + 1] + q[2 * x] + q[2 * x + 1] + 2) >> 2);
        }
        return;
    }
    std::vector<int> xo, yo;
    std::vector<short> xa0, xa1, yb0, yb1;
    linear_coeffs(dw, sw, xo, xa0, xa1, false);
    linear_coeffs(dh, sh, yo, yb0, yb1, true);
    std::vector<int> xo1(dw);
    for (int x = 0; x < dw; x++) xo1[x] = std::min(xo[x] + 1, sw - 1);
    std::vector<int> hbuf[2] = {std::vector<int>(dw), std::vector<int>(dw)};
    int have[2] = {-1, -1};                     // which source row each horizontal buffer holds
    for (int y = 0; y < dh; y++) {
        const int sy[2] = {std::min(std::max(yo[y], 0), sh - 1), std::min(std::max(yo[y] + 1, 0), sh - 1)};
        int slot[2];
        // reuse the horizontal pass of a source row when the previous output row already made it
        for (int k = 0; k < 2; k++) {
            if (have[0] == sy[k]) slot[k] = 0; else if (have[1] == sy[k]) slot[k] = 1; else slot[k] = -1;
        }
        for (int k = 0; k < 2; k++) {
            if (slot[k] >= 0) continue;
            int s = (slot[1 - k] == 0) ? 1 : 0;
            if (k == 1 && sy[1] == sy[0]) { slot[1] = slot[0]; continue; }
            const uchar *r = src.ptr(sy[k]);
            int *hb = hbuf[s].data();
            for (int x = 0; x < dw; x++) hb[x] = r[xo[x]] * xa0[x] + r[xo1[x]] * xa1[x];
            have[s] = sy[k]; slot[k] = s;
        }
        const int *h0 = hbuf[slot[0]].data(), *h1 = hbuf[slot[1]].data();
        const int b0 = yb0[y], b1 = yb1[y];
        uchar *o = dst.ptr(y);
        for (int x = 0; x < dw; x++) {
            const int v = (((b0 * (h0[x] >> 4)) >> 16) + ((b1 * (h1[x] >> 4)) >> 16) + 2) >> 2;
            o[x] = (uchar)(v < 0 ? 0 : v > 255 ? 255 : v);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// cv::GaussianBlur 7x7, sigma 2, 8UC1: OpenCV's fixed-point separable path, kernel [18,34,48,56,48,34,18]/256, horizontal
// pass exact (Q8 in 16 bits), vertical pass in 32 bits, out = (v + 32768) >> 16.  Borders: REFLECT_101.
// ------------------------------------------------------------------------------------------------------------------
static inline int reflect101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
    return i;
}

void GaussianBlur(InputArray src_, OutputArray dst_, Size ksize, double sigmaX, double sigmaY, int borderType)
{
    const Mat src = src_.getMat();
    assert(src.type() == CV_8UC1 && ksize.width == 7 && ksize.height == 7 && sigmaX == 2 && (sigmaY == 2 || sigmaY == 0));
    assert((borderType & ~BORDER_ISOLATED) == BORDER_REFLECT_101);
    (void)borderType; (void)sigmaY;
    const int w = src.cols, h = src.rows;
    std::vector<uint16_t> tmp((size_t)w * h);
    std::vector<uint8_t> padded((size_t)w + 6);
    for (int y = 0; y < h; y++) {
        const uchar *r = src.ptr(y);
        for (int x = -3; x < w + 3; x++) padded[x + 3] = r[reflect101(x, w)];
        const uint8_t *p = padded.data();
        uint16_t *t = &tmp[(size_t)y * w];
        for (int x = 0; x < w; x++)
            t[x] = (uint16_t)(18 * (p[x] + p[x + 6]) + 34 * (p[x + 1] + p[x + 5]) + 48 * (p[x + 2] + p[x + 4]) + 56 * p[x + 3]);
    }
    dst_.create(h, w, src.type());              // in place when dst is src (ORBExtractor.cpp:537): tmp already holds the rows
    Mat dst = dst_.ref();
    for (int y = 0; y < h; y++) {
        const uint16_t *rr[7];
        for (int k = 0; k < 7; k++) rr[k] = &tmp[(size_t)reflect101(y + k - 3, h) * w];
        uchar *o = dst.ptr(y);
        for (int x = 0; x < w; x++) {
            const uint32_t acc = 18u * ((uint32_t)rr[0][x] + rr[6][x]) + 34u * ((uint32_t)rr[1][x] + rr[5][x]) + 48u * ((uint32_t)rr[2][x] + rr[4][x]) + 56u * rr[3][x];
            o[x] = (uchar)((acc + 32768u) >> 16);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// cv::copyMakeBorder, 8-bit: REFLECT_101 / REPLICATE / CONSTANT.  `src` may be a view INSIDE `dst` (ORBExtractor.cpp:579:
// the pyramid level is a ROI of `temp`).  Without BORDER_ISOLATED OpenCV extends a ROI with the parent's real pixels;
// the reference passes whole images there (level 0) and BORDER_ISOLATED for the ROI case, so both reduce to extrapolation.
// ------------------------------------------------------------------------------------------------------------------
void copyMakeBorder(InputArray src_, OutputArray dst_, int top, int bottom, int left, int right, int borderType, const Scalar &value)
{
    const Mat src = src_.getMat();
    assert(src.depth() == CV_8U);
    const int bt = borderType & ~BORDER_ISOLATED;
    const int w = src.cols, h = src.rows, es = (int)src.elemSize();
    dst_.create(h + top + bottom, w + left + right, src.type());
    Mat dst = dst_.ref();
    auto map = [&](int i, int n) -> int {
        if (bt == BORDER_REFLECT_101) return reflect101(i, n);
        if (bt == BORDER_REPLICATE) return i < 0 ? 0 : (i >= n ? n - 1 : i);
        return (i < 0 || i >= n) ? -1 : i;
    };
    // centre first (memmove: src may alias), then left/right of the centre rows, then whole rows above / below
    for (int y = 0; y < h; y++) {
        uchar *d = dst.ptr(y + top) + (size_t)left * es;
        if (d != src.ptr(y)) std::memmove(d, src.ptr(y), (size_t)w * es);
    }
    for (int y = 0; y < h; y++) {
        uchar *row = dst.ptr(y + top);
        const uchar *c = row + (size_t)left * es;
        for (int x = -left; x < w + right; x++) {
            if (x >= 0 && x < w) continue;
            const int sx = map(x, w);
            for (int b = 0; b < es; b++) row[(size_t)(x + left) * es + b] = sx < 0 ? (uchar)value.val[b] : c[(size_t)sx * es + b];
        }
    }
    const size_t rb = (size_t)dst.cols * es;
    for (int y = -top; y < h + bottom; y++) {
        if (y >= 0 && y < h) continue;
        const int sy = map(y, h);
        if (sy < 0) { for (size_t i = 0; i < rb; i++) dst.ptr(y + top)[i] = (uchar)value.val[i % es]; }
        else std::memcpy(dst.ptr(y + top), dst.ptr(sy + top), rb);
    }
}

// cv::cvtColor to gray, 8-bit: (R*9798 + G*19235 + B*3735 + 16384) >> 15  (imgproc/color.cpp RGB2Gray<uchar>, Q15)
void cvtColor(InputArray src_, OutputArray dst_, int code, int)
{
    const Mat src = src_.getMat();
    assert(src.depth() == CV_8U && (src.channels() == 3 || src.channels() == 4));
    const bool rgb = (code == COLOR_RGB2GRAY || code == COLOR_RGBA2GRAY);
    const int cn = src.channels();
    Mat out(src.rows, src.cols, CV_8UC1);
    for (int y = 0; y < src.rows; y++) {
        const uchar *p = src.ptr(y);
        uchar *o = out.ptr(y);
        for (int x = 0; x < src.cols; x++, p += cn) {
            const int r = rgb ? p[0] : p[2], g = p[1], b = rgb ? p[2] : p[0];
            o[x] = (uchar)((r * 9798 + g * 19235 + b * 3735 + 16384) >> 15);
        }
    }
    out.copyTo(dst_);
}

}  // namespace cv
