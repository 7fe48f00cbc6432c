// hyorb_hyslam.hpp -- the C++ host side of the drop-in: hySLAM's own feature interfaces implemented over the C ABI of
// libhyorb (include/hyorb.h).  Header-only, C++14 like hySLAM; include it from a hySLAM translation unit (its include
// path supplies FeatureExtractor.h, ORBFactory.h, FeatureViews.h, Camera.h, FeatureMatcher.h and OpenCV).
//
//   HYSLAM::CudaORBExtractor   : FeatureExtractor    replaces HYSLAM::ORBExtractor        (src/features/ORBExtractor.h:63-116)
//   HYSLAM::CudaORBFactory     : ORBFactory          replaces ORBFactory::getExtractor    (src/features/ORBFactory.cpp:32-40)
//   HYSLAM::CudaStereomatcher                        same public surface as Stereomatcher (src/features/Stereomatcher.h:25-51)
//   HYSLAM::cuda_marshal::...                        FeatureViews / vector<FeatureDescriptor>  <->  SoA arrays of the ABI
//
// The only line of hySLAM that changes is the factory choice in System.cc:77-85 (`new CudaORBFactory(path)` for the
// YAML key `Features: ORB`); ImageProcessing::ProcessStereoImage (src/main/ImageProcessing.cpp:69-116) keeps its code
// and swaps `Stereomatcher` for `CudaStereomatcher`.  tests/cpp compiles this header against test doubles of the hySLAM
// headers and runs ProcessStereoImage's call sequence on a GPU (tests/test_gpu_cpp_shim.py).
#pragma once
#include <FeatureExtractor.h>
#include <FeatureMatcher.h>
#include <FeatureViews.h>
#include <ORBFactory.h>
#include <Camera.h>

#include <cassert>
#include <cstdint>
#include <cstring>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "hyorb.h"

namespace HYSLAM {

static_assert(sizeof(cv::KeyPoint) == sizeof(hyorb_keypoint), "hyorb_keypoint mirrors the layout of cv::KeyPoint");

namespace cuda_marshal {
// vector<FeatureDescriptor> (one cloned 1x32 cv::Mat per keypoint, FeatureDescriptor.cpp:14-18) -> n x 32 bytes
inline std::vector<uint8_t> packDescriptors(const std::vector<FeatureDescriptor> &d)
{
    std::vector<uint8_t> out(d.size() * HYORB_DESC_BYTES);
    for (size_t i = 0; i < d.size(); i++) {
        const cv::Mat m = d[i].rawDescriptor();
        assert(m.rows * m.cols == HYORB_DESC_BYTES);
        std::memcpy(out.data() + i * HYORB_DESC_BYTES, m.data, HYORB_DESC_BYTES);
    }
    return out;
}
// n x 32 bytes -> vector<FeatureDescriptor>, appended like ORBExtractor::operator() does (ORBExtractor.cpp:558-561)
inline void appendDescriptors(const uint8_t *rows, int n, const std::shared_ptr<DescriptorDistance> &dist, std::vector<FeatureDescriptor> &out)
{
    out.reserve(out.size() + n);
    for (int i = 0; i < n; i++)
        out.push_back(FeatureDescriptor(cv::Mat(1, HYORB_DESC_BYTES, CV_8U, const_cast<uint8_t *>(rows) + (size_t)i * HYORB_DESC_BYTES), dist));
}
inline const hyorb_keypoint *asAbi(const std::vector<cv::KeyPoint> &k) { return reinterpret_cast<const hyorb_keypoint *>(k.data()); }
inline void check(int rc) { if (rc != HYORB_OK) throw std::runtime_error(std::string("libhyorb: ") + hyorb_last_error()); }

// hySLAM constructs a Stereomatcher (and criteria objects) per frame; a matcher handle owns a stream and device buffers whose
// creation costs several hundred microseconds, so the per-frame objects of this header borrow handles from a process-wide
// free list instead of creating them (handles are returned on destruction and live until the process exits).
class MatcherPool {
public:
    static hyorb_matcher *acquire(int device)
    {
        {
            std::lock_guard<std::mutex> lock(mutex());
            auto &f = free_list();
            for (size_t i = 0; i < f.size(); i++)
                if (f[i].first == device) { hyorb_matcher *m = f[i].second; f.erase(f.begin() + (long)i); return m; }
        }
        hyorb_matcher *m = nullptr;
        check(hyorb_matcher_create(device, nullptr, &m));
        return m;
    }
    static void release(int device, hyorb_matcher *m)
    {
        if (!m) return;
        std::lock_guard<std::mutex> lock(mutex());
        free_list().push_back(std::make_pair(device, m));
    }
private:
    static std::mutex &mutex() { static std::mutex mu; return mu; }
    static std::vector<std::pair<int, hyorb_matcher *>> &free_list() { static std::vector<std::pair<int, hyorb_matcher *>> f; return f; }
};
}  // namespace cuda_marshal

// ---------------------------------------------------------------------------------------------------------------
class CudaORBExtractor : public FeatureExtractor {
public:
    CudaORBExtractor(std::shared_ptr<DescriptorDistance> dist, FeatureExtractorSettings s, int device = 0)
        : dist_func(dist), nlevels(s.nLevels), scale(s.fScaleFactor)
    {
        hyorb_extractor_params p{s.nFeatures, s.fScaleFactor, s.nLevels, s.N_CELLS, s.init_threshold, s.min_threshold, 0};
        cuda_marshal::check(hyorb_extractor_create(&p, device, nullptr, &h));
        sf.resize(nlevels); inv.resize(nlevels); s2.resize(nlevels); is2.resize(nlevels);
        cuda_marshal::check(hyorb_extractor_get_scales(h, sf.data(), inv.data(), s2.data(), is2.data(), nullptr));
        cap = 4 * s.nFeatures + 1024;          // the quadtree may return more than nFeatures (ORBExtractor.cpp:309-312)
        kp.resize(cap); desc.resize((size_t)cap * HYORB_DESC_BYTES);
    }
    // FeatureExtractor (FeatureExtractor.h:25-37) declares no virtual destructor: own these objects through the
    // std::shared_ptr the factory returns (make_shared remembers the concrete type), as hySLAM does everywhere.
    ~CudaORBExtractor() { hyorb_extractor_destroy(h); }
    CudaORBExtractor(const CudaORBExtractor &) = delete;
    CudaORBExtractor &operator=(const CudaORBExtractor &) = delete;

    // Contract of ORBExtractor::operator() (ORBExtractor.cpp:496-562): the mask is ignored, the image must be CV_8UC1, an
    // empty image returns silently, `keypoints` is cleared, descriptors are appended.
    void operator()(cv::InputArray image, cv::InputArray /*mask*/, std::vector<cv::KeyPoint> &keypoints,
                    std::vector<FeatureDescriptor> &descriptors) override
    {
        if (image.empty()) return;
        const cv::Mat im = image.getMat();
        assert(im.type() == CV_8UC1);
        int n = 0;
        cuda_marshal::check(hyorb_extract_host(h, im.data, im.cols, im.rows, (int)im.step, kp.data(), desc.data(), cap, &n));
        const cv::KeyPoint *k = reinterpret_cast<const cv::KeyPoint *>(kp.data());
        keypoints.assign(k, k + n);
        cuda_marshal::appendDescriptors(desc.data(), n, dist_func, descriptors);
    }
    // ImageProcessing::PreProcessImg (ImageProcessing.cpp:118-138) + operator() in one call: the camera frame (CV_8UC1 / C3 / C4;
    // mbRGB = cam_data.RGB; fscale = cam_data.scale, 1.0f or 0.5f as in the reference's configurations) is uploaded as it is,
    // scaled and converted to gray on the device; `gray` receives the frame the reference keeps as track_data.image (:109).
    void extractFromCameraFrame(const cv::Mat &frame, bool mbRGB, float fscale, cv::Mat &gray, std::vector<cv::KeyPoint> &keypoints,
                                std::vector<FeatureDescriptor> &descriptors)
    {
        if (frame.empty()) return;
        if (fscale != 1.0f && fscale != 0.5f) throw std::runtime_error("CudaORBExtractor: camera scale must be 1.0 or 0.5");
        const int half = fscale == 0.5f ? 1 : 0;
        int gw = 0, gh = 0, n = 0;
        cuda_marshal::check(hyorb_preprocess_size(frame.cols, frame.rows, half, &gw, &gh));
        gray.create(gh, gw, CV_8UC1);
        cuda_marshal::check(hyorb_extract_color_host(h, frame.data, frame.cols, frame.rows, (int)frame.step, frame.channels(), mbRGB ? 1 : 0, half,
                                                     gray.data, (int)gray.step, kp.data(), desc.data(), cap, &n));
        const cv::KeyPoint *k = reinterpret_cast<const cv::KeyPoint *>(kp.data());
        keypoints.assign(k, k + n);
        cuda_marshal::appendDescriptors(desc.data(), n, dist_func, descriptors);
    }
    int GetLevels() override { return nlevels; }
    float GetScaleFactor() override { return scale; }
    std::vector<float> GetScaleFactors() override { return sf; }
    std::vector<float> GetInverseScaleFactors() override { return inv; }
    std::vector<float> GetScaleSigmaSquares() override { return s2; }
    std::vector<float> GetInverseScaleSigmaSquares() override { return is2; }

private:
    hyorb_extractor *h = nullptr;
    std::shared_ptr<DescriptorDistance> dist_func;
    int nlevels, cap = 0;
    float scale;
    std::vector<float> sf, inv, s2, is2;
    std::vector<hyorb_keypoint> kp;
    std::vector<uint8_t> desc;
};

// ---------------------------------------------------------------------------------------------------------------
// Keeps ORBFactory's settings, vocabulary and distance function; only the extractor objects change.
class CudaORBFactory : public ORBFactory {
public:
    CudaORBFactory() : ORBFactory() {}
    explicit CudaORBFactory(std::string settings_path, int device_ = 0) : ORBFactory(settings_path), device(device_) {}
    // getExtractor(std::string camera_type) is NOT overridden: ORBFactory's own version (ORBFactory.cpp:32-35) re-reads the
    // ORB.<camera_type>.Extractor / Matcher block of the settings file (SLAM and Imaging cameras differ: 1000 features / 1.2
    // vs 3000 / 1.4 in config/slam_feature_config.yaml), updates the settings ImageProcessing.cpp:34-36 reads back through
    // getFeatureExtractorSettings(), and then dispatches virtually to the overload below.
    using ORBFactory::getExtractor;
    std::shared_ptr<FeatureExtractor> getExtractor(FeatureExtractorSettings s) override
    {
        return std::make_shared<CudaORBExtractor>(getDistanceFunc(), s, device);
    }

private:
    int device = 0;
};

// ---------------------------------------------------------------------------------------------------------------
// Stereomatcher(views, camera, settings); computeStereoMatches(); getData(...) -- Stereomatcher.h:25-51.
class CudaStereomatcher {
public:
    CudaStereomatcher(FeatureViews views, Camera cam_data, FeatureMatcherSettings settings, int device = 0)
        : mvKeys(views.getKeys()), mvKeysRight(views.getKeysR()), descL(cuda_marshal::packDescriptors(views.getDescriptors())),
          descR(cuda_marshal::packDescriptors(views.getDescriptorsR())), N((int)mvKeys.size())
    {
        const FeatureExtractorSettings orb_params = views.getOrbParams();       // default-constructed at the call site: only size_ref matters
        sp.mbf = cam_data.mbf; sp.fx = cam_data.fx(); sp.n_rows = (int)cam_data.mnMaxY;
        sp.th_high = settings.TH_HIGH; sp.th_low = settings.TH_LOW; sp.size_ref = orb_params.size_ref;
        dev = device;
        m = cuda_marshal::MatcherPool::acquire(device);
    }
    ~CudaStereomatcher() { cuda_marshal::MatcherPool::release(dev, m); }
    CudaStereomatcher(const CudaStereomatcher &) = delete;
    CudaStereomatcher &operator=(const CudaStereomatcher &) = delete;

    void computeStereoMatches()
    {
        mvuRight.assign(N, -1.0f);
        mvDepth.assign(N, -1.0f);
        if (N == 0) return;
        cuda_marshal::check(hyorb_stereo_match_host(m, &sp, cuda_marshal::asAbi(mvKeys), descL.data(), N, cuda_marshal::asAbi(mvKeysRight), descR.data(),
                                                    (int)mvKeysRight.size(), mvuRight.data(), mvDepth.data(), nullptr, nullptr));
    }
    void getData(std::vector<float> &mvuRight_, std::vector<float> &mvDepth_) { mvuRight_ = mvuRight; mvDepth_ = mvDepth; }
    void getData(FeatureViews &views) { views.setuRs(mvuRight); views.setDepths(mvDepth); }

private:
    std::vector<cv::KeyPoint> mvKeys, mvKeysRight;
    std::vector<uint8_t> descL, descR;
    hyorb_stereo_params sp;
    hyorb_matcher *m = nullptr;
    std::vector<float> mvuRight, mvDepth;
    int N, dev = 0;
};

// ---------------------------------------------------------------------------------------------------------------
// ImageProcessing::ProcessStereoImage (ImageProcessing.cpp:69-116) as ONE device call: extractor_left(imL), extractor_right(imR)
// and Stereomatcher(...).computeStereoMatches() run back to back on the GPU (hyorb_process_stereo_batch_host with one pair), so
// the keypoints and descriptors are not downloaded, repacked and uploaded again between extraction and association and no
// second host thread is needed.  Returns the FeatureViews the reference builds at :98-103 (keys, right keys, descriptors, uR,
// depth); results are identical to the three-object path above.
class CudaStereoFrontEnd {
public:
    CudaStereoFrontEnd(std::shared_ptr<DescriptorDistance> dist, FeatureExtractorSettings s, int device = 0) : dist_func(dist)
    {
        hyorb_extractor_params p{s.nFeatures, s.fScaleFactor, s.nLevels, s.N_CELLS, s.init_threshold, s.min_threshold, 0};
        cuda_marshal::check(hyorb_extractor_create(&p, device, nullptr, &h));
        cap = 4 * s.nFeatures + 1024;
        kp.resize((size_t)2 * cap); desc.resize((size_t)2 * cap * HYORB_DESC_BYTES); uR.resize(cap); depth.resize(cap);
    }
    ~CudaStereoFrontEnd() { hyorb_extractor_destroy(h); }
    CudaStereoFrontEnd(const CudaStereoFrontEnd &) = delete;
    CudaStereoFrontEnd &operator=(const CudaStereoFrontEnd &) = delete;

    FeatureViews processStereoImage(const cv::Mat &imGrayLeft, const cv::Mat &imGrayRight, const Camera &cam_data, const FeatureMatcherSettings &settings,
                                    FeatureExtractorSettings orb_params = FeatureExtractorSettings())
    {
        assert(imGrayLeft.type() == CV_8UC1 && imGrayRight.type() == CV_8UC1 && imGrayLeft.rows == imGrayRight.rows && imGrayLeft.cols == imGrayRight.cols);
        const int w = imGrayLeft.cols, hgt = imGrayLeft.rows;
        pair.resize((size_t)2 * w * hgt);            // left | right, dense rows: the ABI takes one base pointer and an image stride
        for (int y = 0; y < hgt; y++) {
            std::memcpy(pair.data() + (size_t)y * w, imGrayLeft.ptr<unsigned char>(y), (size_t)w);
            std::memcpy(pair.data() + (size_t)(hgt + y) * w, imGrayRight.ptr<unsigned char>(y), (size_t)w);
        }
        hyorb_stereo_params sp;
        sp.mbf = cam_data.mbf; sp.fx = cam_data.fx(); sp.n_rows = (int)cam_data.mnMaxY;
        sp.th_high = settings.TH_HIGH; sp.th_low = settings.TH_LOW; sp.size_ref = orb_params.size_ref;
        int32_t counts[2] = {0, 0};
        cuda_marshal::check(hyorb_process_stereo_batch_host(h, &sp, pair.data(), 1, w, hgt, w, (size_t)w * hgt, kp.data(), desc.data(), cap, counts,
                                                            uR.data(), depth.data()));
        const cv::KeyPoint *kl = reinterpret_cast<const cv::KeyPoint *>(kp.data()), *kr = kl + cap;
        std::vector<FeatureDescriptor> dl, dr;
        cuda_marshal::appendDescriptors(desc.data(), counts[0], dist_func, dl);
        cuda_marshal::appendDescriptors(desc.data() + (size_t)cap * HYORB_DESC_BYTES, counts[1], dist_func, dr);
        FeatureViews views(std::vector<cv::KeyPoint>(kl, kl + counts[0]), std::vector<cv::KeyPoint>(kr, kr + counts[1]), dl, dr, orb_params);
        views.setuRs(std::vector<float>(uR.begin(), uR.begin() + counts[0]));
        views.setDepths(std::vector<float>(depth.begin(), depth.begin() + counts[0]));
        return views;
    }

private:
    hyorb_extractor *h = nullptr;
    std::shared_ptr<DescriptorDistance> dist_func;
    int cap = 0;
    std::vector<hyorb_keypoint> kp;
    std::vector<uint8_t> desc, pair;
    std::vector<float> uR, depth;
};

// ---------------------------------------------------------------------------------------------------------------
// The Hamming scans of FeatureMatcher (SearchForTriangulation / SearchByBoW inner loops, FeatureMatcher.cc:281-345, with
// BestMatchBoWCriterion, MatchCriteria.cpp:601-635) over explicit candidate lists; projection and map bookkeeping stay
// with hySLAM's FeatureMatcher.
class CudaDescriptorScan {
public:
    explicit CudaDescriptorScan(int device = 0) : dev(device) { m = cuda_marshal::MatcherPool::acquire(device); }
    ~CudaDescriptorScan() { cuda_marshal::MatcherPool::release(dev, m); }
    CudaDescriptorScan(const CudaDescriptorScan &) = delete;
    CudaDescriptorScan &operator=(const CudaDescriptorScan &) = delete;

    struct Result { std::vector<int32_t> best_idx; std::vector<uint16_t> best, second; std::vector<uint8_t> accepted; };

    // rule: HYORB_RULE_LANDMARK / HYORB_RULE_BOW / HYORB_RULE_MONOINIT; cand_off/cand_idx = CSR candidate lists (nullptr:
    // every target in index order)
    Result scan(const std::vector<FeatureDescriptor> &queries, const std::vector<FeatureDescriptor> &targets, const int32_t *cand_off,
                const int32_t *cand_idx, int rule, float thr, float ratio)
    {
        const std::vector<uint8_t> q = cuda_marshal::packDescriptors(queries), t = cuda_marshal::packDescriptors(targets);
        const int nq = (int)queries.size(), nt = (int)targets.size();
        Result r;
        r.best_idx.assign(nq, -1); r.best.assign(nq, 65535); r.second.assign(nq, 65535); r.accepted.assign(nq, 0);
        cuda_marshal::check(hyorb_match_csr_host(m, q.data(), nq, t.data(), nt, cand_off, cand_idx, rule, thr, ratio, r.best_idx.data(), r.best.data(),
                                                 r.second.data(), r.accepted.data()));
        return r;
    }

    // FeatureMatcher::SearchByProjection up to its association loop (FeatureMatcher.cc:57-213): the caller fills `pr` from the
    // frame (mRcw, mtcw, GetCameraCenter(), camera K / mbf / sensor / image bounds) and one hyorb_landmark per candidate MapPoint
    // (GetWorldPos, getSize, Get{Min,Max}DistanceInvariance, frame.hasAssociation(lm)), in the iteration order of its
    // std::map<MapPoint*, ...>; `flags` = HYORB_SBP_* picks the variant, lm_prev_angle is needed with HYORB_SBP_ROTATION.
    // Result::accepted[i] != 0  <=>  the reference would call frame.associateLandMark(best_idx[i], landmark i, true).
    Result searchByProjection(const hyorb_projection &pr, const std::vector<hyorb_landmark> &landmarks, const std::vector<FeatureDescriptor> &lm_desc,
                              const std::vector<float> *lm_prev_angle, const FeatureViews &views, const std::vector<uint8_t> *already_matched,
                              float th, float thr, float ratio, unsigned flags)
    {
        const std::vector<cv::KeyPoint> keys = views.getKeys();
        const std::vector<uint8_t> q = cuda_marshal::packDescriptors(lm_desc), t = cuda_marshal::packDescriptors(views.getDescriptors());
        const std::vector<float> uR = views.getuRs();
        const int n = (int)landmarks.size(), nt = (int)keys.size();
        Result r;
        r.best_idx.assign(n, -1); r.best.assign(n, 65535); r.second.assign(n, 65535); r.accepted.assign(n, 0);
        cuda_marshal::check(hyorb_search_by_projection_ex_host(m, &pr, landmarks.data(), q.data(), lm_prev_angle ? lm_prev_angle->data() : nullptr, n,
                                                               cuda_marshal::asAbi(keys), t.data(), uR.empty() ? nullptr : uR.data(),
                                                               already_matched ? already_matched->data() : nullptr, nt, th,
                                                               views.getOrbParams().size_ref, thr, ratio, flags, r.best_idx.data(), r.best.data(),
                                                               r.second.data(), r.accepted.data(), nullptr));
        return r;
    }

    // FeatureMatcher::SearchForTriangulation (FeatureMatcher.cc:373-402) up to RotationConsistencyBoW: cand_off / cand_idx = for every
    // eligible feature of pKF1 (index criteria applied by the caller) the features of pKF2 under the same vocabulary node, in
    // FeatureVector order; F12 = the 3x3 CV_32F fundamental matrix; EpipolarConsistencyBoWCriterion + BestMatchBoWCriterion(TH_LOW, 1.0)
    // run on the device.  accepted[i] != 0  <=>  matches_internal gets (i, best_idx[i]).
    Result searchForTriangulation(const FeatureViews &views1, const FeatureViews &views2, const int32_t *cand_off, const int32_t *cand_idx,
                                  const cv::Mat &F12, float TH_LOW)
    {
        const std::vector<cv::KeyPoint> k1 = views1.getKeys(), k2 = views2.getKeys();
        const std::vector<uint8_t> q = cuda_marshal::packDescriptors(views1.getDescriptors()), t = cuda_marshal::packDescriptors(views2.getDescriptors());
        float F[9];
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) F[3 * r + c] = F12.at<float>(r, c);
        FeatureExtractorSettings orb_params = views2.getOrbParams();
        const int nq = (int)k1.size(), nt = (int)k2.size();
        Result r;
        r.best_idx.assign(nq, -1); r.best.assign(nq, 65535); r.second.assign(nq, 65535); r.accepted.assign(nq, 0);
        cuda_marshal::check(hyorb_match_csr_epipolar_host(m, cuda_marshal::asAbi(k1), q.data(), nq, cuda_marshal::asAbi(k2), t.data(), nt, cand_off, cand_idx,
                                                          F, orb_params.sigma_ref, orb_params.size_ref, HYORB_RULE_BOW, TH_LOW, 1.0f, r.best_idx.data(),
                                                          r.best.data(), r.second.data(), r.accepted.data()));
        return r;
    }

    // MapPointDBEntry::_computeDistinctiveDescriptor_ (MapPointDB.cpp:127-171) for many landmarks: observations[l] = descriptors of landmark l
    std::vector<int32_t> distinctiveDescriptors(const std::vector<std::vector<FeatureDescriptor>> &observations)
    {
        std::vector<int32_t> off(1, 0), best(observations.size(), -1), median(observations.size(), -1);
        std::vector<uint8_t> all;
        for (const auto &obs : observations) {
            const std::vector<uint8_t> d = cuda_marshal::packDescriptors(obs);
            all.insert(all.end(), d.begin(), d.end());
            off.push_back((int32_t)(all.size() / HYORB_DESC_BYTES));
        }
        cuda_marshal::check(hyorb_distinctive_descriptor_host(m, all.data(), off.data(), (int)observations.size(), best.data(), median.data()));
        return best;
    }

private:
    hyorb_matcher *m = nullptr;
    int dev = 0;
};

}  // namespace HYSLAM
