// ref_glue.cpp -- C entry points over the REFERENCE's own translation units (oracle/_ref/libhyslam_ref.so).
// TEST INFRASTRUCTURE (oracle/), not product code: used by tests/ as the parity pin and by bench.py's CPU arm.
//
// The library is built by oracle/Makefile from these files of /root/reference, compiled UNMODIFIED where they lie:
//   src/features/ORBExtractor.cpp, src/features/Stereomatcher.cpp, src/features/FeatureExtractor.cpp,
//   src/features/low_level/{ORBFinder,DescriptorDistance,FeatureDescriptor,FeatureFinder}.cpp,
//   src/core/{FeatureViews,FeatureExtractorSettings}.cpp
// against oracle/cvshim (a minimal OpenCV-compatible header + the five OpenCV algorithms the path calls, each pinned to
// cv2).  This file only constructs the reference's objects the way the reference does and copies results out:
//   extractor   = ORBFactory::getExtractor(settings)            (ORBFactory.cpp:37-40; that TU itself needs DBoW2)
//   stereo      = ImageProcessing::ProcessStereoImage           (src/main/ImageProcessing.cpp:85-103)
//
// Quadtree tie policy.  ORBExtractor.cpp:324 sorts (count, ExtractorNode*) pairs: among nodes of equal size the heap
// ADDRESS decides which is split first.  `arena = 1` runs the extraction under a monotonic, never-reusing operator new,
// so that address order == creation order (the canonical policy of SURVEY A.4); `arena = 0` uses glibc malloc, i.e.
// what a stock hySLAM build does.  tests/test_oracle_vs_ref.py reports how often the two differ.
#include <ORBExtractor.h>
#include <ORBFinder.h>
#include <Stereomatcher.h>
#include <FeatureViews.h>
#include <Camera.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>
#include <thread>
#include <vector>

// ---------------------------------------------------------------------------------------------------------------
// scoped monotonic arena behind operator new (this library is linked -Bsymbolic: only its own allocations come here)
// ---------------------------------------------------------------------------------------------------------------
namespace {
struct Arena {
    char *base = nullptr; size_t cap = 0, used = 0; bool active = false;
    ~Arena() { std::free(base); }
};
thread_local Arena g_arena;
inline bool in_arena(const void *p) { return g_arena.base && (const char *)p >= g_arena.base && (const char *)p < g_arena.base + g_arena.cap; }
void *arena_or_malloc(size_t n)
{
    if (g_arena.active) {
        const size_t need = (n + 15) & ~(size_t)15;
        if (g_arena.used + need <= g_arena.cap) { void *p = g_arena.base + g_arena.used; g_arena.used += need; return p; }
        std::abort();                               // never fall back silently: the tie policy would change
    }
    void *p = std::malloc(n ? n : 1);
    if (!p) throw std::bad_alloc();
    return p;
}
struct ArenaScope {
    explicit ArenaScope(bool on)
    {
        if (!on) return;
        if (!g_arena.base) { g_arena.cap = (size_t)1 << 31; g_arena.base = (char *)std::malloc(g_arena.cap); if (!g_arena.base) std::abort(); }
        g_arena.used = 0; g_arena.active = true;
    }
    ~ArenaScope() { g_arena.active = false; }
};
}  // namespace
// hidden: only this library's own (-Bsymbolic) allocations come here; nothing that links against the library may bind to them
#define REF_HIDDEN __attribute__((visibility("hidden")))
REF_HIDDEN void *operator new(size_t n) { return arena_or_malloc(n); }
REF_HIDDEN void *operator new[](size_t n) { return arena_or_malloc(n); }
REF_HIDDEN void operator delete(void *p) noexcept { if (p && !in_arena(p)) std::free(p); }
REF_HIDDEN void operator delete[](void *p) noexcept { if (p && !in_arena(p)) std::free(p); }
REF_HIDDEN void operator delete(void *p, size_t) noexcept { if (p && !in_arena(p)) std::free(p); }
REF_HIDDEN void operator delete[](void *p, size_t) noexcept { if (p && !in_arena(p)) std::free(p); }

extern "C" {

#define REF_API __attribute__((visibility("default")))

struct ref_keypoint { float x, y, size, angle, response; int32_t octave, class_id; };   // == cv::KeyPoint
struct ref_params { int32_t nfeatures; float scale_factor; int32_t nlevels, cell_px, ini_th, min_th; };
struct ref_stereo_params { float mbf, fx; int32_t n_rows; float th_high, th_low, size_ref; };

static HYSLAM::FeatureExtractorSettings to_settings(const ref_params *p)
{
    HYSLAM::FeatureExtractorSettings s;
    s.nFeatures = p->nfeatures; s.fScaleFactor = p->scale_factor; s.nLevels = p->nlevels; s.N_CELLS = p->cell_px;
    s.init_threshold = p->ini_th; s.min_threshold = p->min_th;
    return s;
}

static std::shared_ptr<HYSLAM::FeatureExtractor> make_extractor(const ref_params *p)
{
    // ORBFactory::getExtractor(FeatureExtractorSettings), ORBFactory.cpp:37-40
    std::shared_ptr<HYSLAM::DescriptorDistance> dist_func = std::make_shared<HYSLAM::ORBDistance>();
    return std::make_shared<HYSLAM::ORBExtractor>(std::make_unique<HYSLAM::ORBFinder>(20.0, true), dist_func, to_settings(p));
}

// One frame through HYSLAM::ORBExtractor::operator().  level_images (optional): nlevels pointers that receive the
// pyramid levels (dense, level width as pitch) for stage-wise comparison.  Returns 0, or -2 if cap is too small.
REF_API int ref_extract(const ref_params *p, const uint8_t *image, int w, int h, int stride, int arena,
                        ref_keypoint *kps, uint8_t *desc, int cap, int32_t *n_out, uint8_t **level_images)
{
    int rc = 0;
    *n_out = 0;
    {
        ArenaScope scope(arena != 0);
        {
            std::shared_ptr<HYSLAM::FeatureExtractor> ex = make_extractor(p);
            cv::Mat img(h, w, CV_8UC1, (void *)image, (size_t)stride);
            std::vector<cv::KeyPoint> keypoints;
            std::vector<HYSLAM::FeatureDescriptor> descriptors;
            (*ex)(img, cv::Mat(), keypoints, descriptors);
            const int n = (int)keypoints.size();
            *n_out = n;
            if (n > cap || (int)descriptors.size() != n) rc = -2;
            else {
                static_assert(sizeof(cv::KeyPoint) == sizeof(ref_keypoint), "cv::KeyPoint layout");
                if (n) std::memcpy(kps, keypoints.data(), sizeof(ref_keypoint) * (size_t)n);
                for (int i = 0; i < n; i++) {
                    const cv::Mat d = descriptors[i].rawDescriptor();
                    std::memcpy(desc + 32 * (size_t)i, d.ptr(), 32);
                }
            }
            if (level_images) {
                HYSLAM::ORBExtractor *orb = static_cast<HYSLAM::ORBExtractor *>(ex.get());
                for (int l = 0; l < p->nlevels; l++) {
                    const cv::Mat &m = orb->mvImagePyramid[l];
                    if (level_images[l]) for (int y = 0; y < m.rows; y++) std::memcpy(level_images[l] + (size_t)y * m.cols, m.ptr(y), (size_t)m.cols);
                }
            }
        }                                           // everything allocated inside is destroyed before the arena is reused
    }
    return rc;
}

REF_API int ref_scale_tables(const ref_params *p, float *scale, float *inv_scale, float *sigma2, float *inv_sigma2)
{
    std::shared_ptr<HYSLAM::FeatureExtractor> ex = make_extractor(p);
    const std::vector<float> a = ex->GetScaleFactors(), b = ex->GetInverseScaleFactors(), c = ex->GetScaleSigmaSquares(), d = ex->GetInverseScaleSigmaSquares();
    for (int i = 0; i < p->nlevels; i++) { scale[i] = a[i]; inv_scale[i] = b[i]; sigma2[i] = c[i]; inv_sigma2[i] = d[i]; }
    return ex->GetLevels();
}

// HYSLAM::ORBDistance::distance through FeatureDescriptor::distance (DescriptorDistance.cpp:9-25, FeatureDescriptor.cpp:14-18)
REF_API float ref_hamming(const uint8_t *a, const uint8_t *b)
{
    std::shared_ptr<HYSLAM::DescriptorDistance> dist_func = std::make_shared<HYSLAM::ORBDistance>();
    cv::Mat ma(1, 32, CV_8UC1, (void *)a), mb(1, 32, CV_8UC1, (void *)b);
    HYSLAM::FeatureDescriptor da(ma, dist_func), db(mb, dist_func);
    return da.distance(db);
}

// C4 CPU arms (bench.py c4_match): a brute-force scan of `nq` query descriptors over `nt` targets with the best / second-best
// bookkeeping of BestMatchBoWCriterion (MatchCriteria.cpp:601-635).  lean = 0: every distance through the reference's own
// FeatureDescriptor::distance -> ORBDistance::distance (FeatureDescriptor.cpp:14-18, DescriptorDistance.cpp:9-25: a cv::Mat header copy
// per call, the bit-trick popcount); lean = 1: contiguous descriptors and the hardware popcount.  Returns the number of distances.
#if defined(__x86_64__)
__attribute__((target("popcnt")))
#endif
REF_API long ref_bf_scan(const uint8_t *q, int nq, const uint8_t *t, int nt, int lean, int32_t *best_idx, int32_t *best, int32_t *second)
{
    if (lean) {
        for (int i = 0; i < nq; i++) {
            const uint64_t *a = (const uint64_t *)(q + 32 * (size_t)i);
            int b1 = 256 + 1, b2 = 256 + 1, bi = -1;
            for (int j = 0; j < nt; j++) {
                const uint64_t *b = (const uint64_t *)(t + 32 * (size_t)j);
                const int d = __builtin_popcountll(a[0] ^ b[0]) + __builtin_popcountll(a[1] ^ b[1]) + __builtin_popcountll(a[2] ^ b[2]) + __builtin_popcountll(a[3] ^ b[3]);
                if (d < b1) { b2 = b1; b1 = d; bi = j; } else if (d < b2) b2 = d;
            }
            best_idx[i] = bi; best[i] = b1; second[i] = b2;
        }
        return (long)nq * nt;
    }
    std::shared_ptr<HYSLAM::DescriptorDistance> dist_func = std::make_shared<HYSLAM::ORBDistance>();
    std::vector<HYSLAM::FeatureDescriptor> dq, dt;
    dq.reserve(nq); dt.reserve(nt);
    for (int i = 0; i < nq; i++) dq.push_back(HYSLAM::FeatureDescriptor(cv::Mat(1, 32, CV_8UC1, (void *)(q + 32 * (size_t)i)), dist_func));
    for (int j = 0; j < nt; j++) dt.push_back(HYSLAM::FeatureDescriptor(cv::Mat(1, 32, CV_8UC1, (void *)(t + 32 * (size_t)j)), dist_func));
    for (int i = 0; i < nq; i++) {
        float b1 = 1e30f, b2 = 1e30f; int bi = -1;
        for (int j = 0; j < nt; j++) {
            const float d = dq[i].distance(dt[j]);
            if (d < b1) { b2 = b1; b1 = d; bi = j; } else if (d < b2) b2 = d;
        }
        best_idx[i] = bi; best[i] = (int32_t)b1; second[i] = b2 < 1e29f ? (int32_t)b2 : 257;
    }
    return (long)nq * nt;
}

static std::vector<HYSLAM::FeatureDescriptor> wrap_descriptors(const uint8_t *d, int n, std::shared_ptr<HYSLAM::DescriptorDistance> dist_func)
{
    std::vector<HYSLAM::FeatureDescriptor> out;
    out.reserve(n);
    for (int i = 0; i < n; i++) out.push_back(HYSLAM::FeatureDescriptor(cv::Mat(1, 32, CV_8UC1, (void *)(d + 32 * (size_t)i)), dist_func));
    return out;
}

// HYSLAM::Stereomatcher as ImageProcessing::ProcessStereoImage drives it (ImageProcessing.cpp:85-103): FeatureViews from
// the two keypoint / descriptor sets, a default FeatureExtractorSettings (only size_ref matters), Camera{K, mbf, mnMaxY}.
// An empty match list makes the reference read v[0] of an empty vector (Stereomatcher.cpp:143); that case is reported
// as rc 1 with all outputs -1 instead of being executed.
REF_API int ref_stereo_match(const ref_stereo_params *sp, const ref_keypoint *kl, const uint8_t *dl, int nl,
                             const ref_keypoint *kr, const uint8_t *dr, int nr, float *uR, float *depth)
{
    std::shared_ptr<HYSLAM::DescriptorDistance> dist_func = std::make_shared<HYSLAM::ORBDistance>();
    std::vector<cv::KeyPoint> keysL(nl), keysR(nr);
    if (nl) std::memcpy(keysL.data(), kl, sizeof(ref_keypoint) * (size_t)nl);
    if (nr) std::memcpy(keysR.data(), kr, sizeof(ref_keypoint) * (size_t)nr);
    HYSLAM::FeatureExtractorSettings orb_params;                    // default-constructed, as at ImageProcessing.cpp:85
    orb_params.size_ref = sp->size_ref;
    HYSLAM::FeatureViews views(keysL, keysR, wrap_descriptors(dl, nl, dist_func), wrap_descriptors(dr, nr, dist_func), orb_params);
    HYSLAM::Camera cam;
    cam.K = cv::Mat::eye(3, 3, CV_32F);
    cam.K.at<float>(0, 0) = sp->fx;
    cam.mbf = sp->mbf;
    cam.mnMaxY = (float)sp->n_rows;
    cam.sensor = 1;
    HYSLAM::FeatureMatcherSettings ms;
    ms.TH_HIGH = sp->th_high; ms.TH_LOW = sp->th_low;
    HYSLAM::Stereomatcher matcher(views, cam, ms);
    // probe for the empty-match case on a copy of the logic's precondition: run, but guard the undefined read
    for (int i = 0; i < nl; i++) { uR[i] = -1.f; depth[i] = -1.f; }
    if (nl == 0 || nr == 0) return 1;
    matcher.computeStereoMatches();
    std::vector<float> u, d;
    matcher.getData(u, d);
    for (int i = 0; i < nl; i++) { uR[i] = u[i]; depth[i] = d[i]; }
    return 0;
}

// Throughput legs for bench.py: n_pairs stereo pairs, extract L + R + stereo match, the reference's threading shape
// (ImageProcessing.cpp:82-84: the left extractor on a transient std::thread, the right one on the caller) when
// `threads_per_pair` == 2, everything on the caller when 1.  Images are dense (stride == w).  Returns total keypoints.
REF_API long ref_process_stereo_pairs(const ref_params *p, const ref_stereo_params *sp, const uint8_t *left, const uint8_t *right,
                                      int n_pairs, int w, int h, int threads_per_pair, long *n_matches_out)
{
    std::shared_ptr<HYSLAM::FeatureExtractor> exL = make_extractor(p), exR = make_extractor(p);   // ImageProcessing.cpp:31-32
    std::shared_ptr<HYSLAM::DescriptorDistance> dist_func = std::make_shared<HYSLAM::ORBDistance>();
    long total = 0, matches = 0;
    for (int i = 0; i < n_pairs; i++) {
        cv::Mat imL(h, w, CV_8UC1, (void *)(left + (size_t)i * w * h)), imR(h, w, CV_8UC1, (void *)(right + (size_t)i * w * h));
        std::vector<cv::KeyPoint> kL, kR;
        std::vector<HYSLAM::FeatureDescriptor> dL, dR;
        if (threads_per_pair >= 2) {
            std::thread t([&]() { (*exL)(imL, cv::Mat(), kL, dL); });
            (*exR)(imR, cv::Mat(), kR, dR);
            t.join();
        } else {
            (*exL)(imL, cv::Mat(), kL, dL);
            (*exR)(imR, cv::Mat(), kR, dR);
        }
        total += (long)kL.size() + (long)kR.size();
        if (sp && !kL.empty() && !kR.empty()) {
            HYSLAM::FeatureExtractorSettings orb_params;
            orb_params.size_ref = sp->size_ref;
            HYSLAM::FeatureViews views(kL, kR, dL, dR, orb_params);
            HYSLAM::Camera cam;
            cam.K = cv::Mat::eye(3, 3, CV_32F);
            cam.K.at<float>(0, 0) = sp->fx;
            cam.mbf = sp->mbf; cam.mnMaxY = (float)sp->n_rows; cam.sensor = 1;
            HYSLAM::FeatureMatcherSettings ms;
            ms.TH_HIGH = sp->th_high; ms.TH_LOW = sp->th_low;
            HYSLAM::Stereomatcher matcher(views, cam, ms);
            matcher.computeStereoMatches();
            std::vector<float> u, d;
            matcher.getData(u, d);
            for (float v : u) matches += v >= 0;
        }
    }
    if (n_matches_out) *n_matches_out = matches;
    return total;
}

// ---- cvshim primitives exposed for pinning against cv2 (tests/test_cvshim_vs_cv2.py) ----
REF_API int cvshim_fast(const uint8_t *img, int w, int h, int stride, int threshold, int nms, float *xs, float *ys, float *resp, int cap)
{
    std::vector<cv::KeyPoint> k;
    cv::FAST(cv::Mat(h, w, CV_8UC1, (void *)img, (size_t)stride), k, threshold, nms != 0);
    if ((int)k.size() > cap) return -2;
    for (size_t i = 0; i < k.size(); i++) { xs[i] = k[i].pt.x; ys[i] = k[i].pt.y; resp[i] = k[i].response; }
    return (int)k.size();
}
REF_API void cvshim_resize(const uint8_t *src, int sw, int sh, int sstride, uint8_t *dst, int dw, int dh, int dstride)
{
    cv::Mat d(dh, dw, CV_8UC1, dst, (size_t)dstride);
    cv::resize(cv::Mat(sh, sw, CV_8UC1, (void *)src, (size_t)sstride), d, cv::Size(dw, dh), 0, 0, cv::INTER_LINEAR);
}
REF_API void cvshim_blur(const uint8_t *src, int w, int h, int sstride, uint8_t *dst, int dstride)
{
    cv::Mat d(h, w, CV_8UC1, dst, (size_t)dstride);
    cv::GaussianBlur(cv::Mat(h, w, CV_8UC1, (void *)src, (size_t)sstride), d, cv::Size(7, 7), 2, 2, cv::BORDER_REFLECT_101);
}
REF_API void cvshim_border(const uint8_t *src, int w, int h, int sstride, uint8_t *dst, int dstride, int b, int border_type)
{
    cv::Mat d(h + 2 * b, w + 2 * b, CV_8UC1, dst, (size_t)dstride);
    cv::copyMakeBorder(cv::Mat(h, w, CV_8UC1, (void *)src, (size_t)sstride), d, b, b, b, b, border_type);
}
REF_API float cvshim_fast_atan2(float y, float x) { return cv::fastAtan2(y, x); }
REF_API void cvshim_gemm(const float *A, int m, int k, int ta, const float *B, int n, const float *C, float *D)
{
    cv::Mat a(ta ? k : m, ta ? m : k, CV_32F, (void *)A), b(k, n, CV_32F, (void *)B);
    cv::Mat d;
    if (C) { cv::Mat c(m, n, CV_32F, (void *)C); d = ta ? cv::Mat(a.t() * b + c) : cv::Mat(a * b + c); }
    else d = ta ? cv::Mat(a.t() * b) : cv::Mat(a * b);
    for (int i = 0; i < m; i++) for (int j = 0; j < n; j++) D[i * n + j] = d.at<float>(i, j);
}
REF_API double cvshim_norm(const float *a, int n) { return cv::norm(cv::Mat(n, 1, CV_32F, (void *)a)); }

}  // extern "C"
