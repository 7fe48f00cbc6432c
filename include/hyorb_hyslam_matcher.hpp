// hyorb_hyslam_matcher.hpp -- HYSLAM::CudaFeatureMatcher: the public surface of hySLAM's FeatureMatcher (src/features/FeatureMatcher.h:105-176,
// same names, same argument lists, same observable results) with the descriptor / window / projection scans running on the GPU through
// the C ABI of libhyorb (include/hyorb.h).  Header-only, C++14; include it from a hySLAM translation unit next to hyorb_hyslam.hpp.
//
// FeatureMatcher is a concrete class created by the non-virtual FeatureFactory::getFeatureMatcher() (FeatureFactory.cpp:7-9), so the drop-in
// is a link-time / source-level substitution: `using FeatureMatcher = CudaFeatureMatcher;` at the call sites, or a one-line change of
// getFeatureMatcher().  What stays on the host, exactly as in the reference: reading Frame / KeyFrame / MapPoint state into flat arrays,
// the pointer-ordered bookkeeping around the scans (std::map<MapPoint*, ...> iteration, associateLandMark, fuse_matches.insert,
// set_difference of MapPoint* lists) and the DBoW2 FeatureVector walk.  What moves to the device: landmark projection and landmark
// criteria, grid window queries, every view criterion, all Hamming scans, the rotation histograms, the mono-initialisation matcher.
//
// tests/cpp/matcher_shim_test.cpp runs this class against the reference's REAL Frame / KeyFrame / MapPoint classes (oracle/_ref) and
// tests/test_gpu_cpp_matcher_shim.py compares every entry point with HYSLAM::FeatureMatcher itself on the same scenes.
//
// Not provided: Fuse(pKF, Scw, ...) -- its body is commented out in the reference (FeatureMatcher.cc:523-627 returns without a value);
// SearchByProjection(pKF, Scw, ...) and the two legacy KeyFrame-KeyFrame SearchByBoW overloads (only reached from LoopClosing, which the
// reference marks as dead code).
#pragma once
#include <FeatureMatcher.h>
#include <FeatureViews.h>
#include <Frame.h>
#include <KeyFrame.h>
#include <MapPoint.h>
#include <Camera.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <iterator>
#include <map>
#include <set>
#include <vector>

#include "hyorb_hyslam.hpp"

namespace HYSLAM {

class CudaFeatureMatcher {
public:
    explicit CudaFeatureMatcher(float nnratio = 0.6, bool checkOri = true, int device = 0) : mfNNratio(nnratio), mbCheckOrientation(checkOri), dev(device)
    {
        m = cuda_marshal::MatcherPool::acquire(device);
    }
    explicit CudaFeatureMatcher(FeatureMatcherSettings settings, int device = 0)
        : mfNNratio(settings.nnratio), mbCheckOrientation(settings.checkOri), TH_LOW(settings.TH_LOW), TH_HIGH(settings.TH_HIGH), dev(device)
    {
        m = cuda_marshal::MatcherPool::acquire(device);
    }
    ~CudaFeatureMatcher() { cuda_marshal::MatcherPool::release(dev, m); }
    CudaFeatureMatcher(const CudaFeatureMatcher &) = delete;
    CudaFeatureMatcher &operator=(const CudaFeatureMatcher &) = delete;

    // ---- FeatureMatcher.h:114 -- track the local map (FeatureMatcher.cc:123-143)
    int SearchByProjection(Frame &F, const std::vector<MapPoint *> &vpMapPoints, const float th = 3)
    {
        return searchByProjection(F, vpMapPoints, th, HYORB_SBP_DISTANCE | HYORB_SBP_STEREO, TH_HIGH, mfNNratio, nullptr);
    }
    // ---- FeatureMatcher.h:118 -- track from the previous frame (FeatureMatcher.cc:145-176)
    int SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, const float th, const bool /*bMono*/)
    {
        const std::vector<MapPoint *> vpMapPoints = LastFrame.replicatemvpMapPoints();
        return searchByProjection(CurrentFrame, vpMapPoints, th, HYORB_SBP_STEREO | HYORB_SBP_ROTATION, TH_HIGH, mfNNratio, &LastFrame);
    }
    // ---- FeatureMatcher.h:122 -- relocalisation (FeatureMatcher.cc:180-213).  The reference lists RotationConsistencyCriterion here too, but
    // never hands it a previous frame, so the criterion returns its input (MatchCriteria.cpp:368-370): no rotation test in effect.
    int SearchByProjection(Frame &CurrentFrame, KeyFrame *pKF, const std::set<MapPoint *> &sAlreadyFound, const float th, const int ORBdist)
    {
        std::vector<MapPoint *> landmarks_pKF = pKF->getAssociatedLandMarks();
        std::vector<MapPoint *> found(sAlreadyFound.begin(), sAlreadyFound.end());
        std::sort(landmarks_pKF.begin(), landmarks_pKF.end());
        std::sort(found.begin(), found.end());
        std::vector<MapPoint *> vpMapPoints;
        std::set_difference(landmarks_pKF.begin(), landmarks_pKF.end(), found.begin(), found.end(), std::inserter(vpMapPoints, vpMapPoints.begin()));
        return searchByProjection(CurrentFrame, vpMapPoints, th, HYORB_SBP_DISTANCE, (float)ORBdist, 1.00f, nullptr);
    }

    // ---- FeatureMatcher.h:131 -- relocalisation / loop detection against a frame (FeatureMatcher.cc:216-280)
    int SearchByBoW(KeyFrame *pKF, Frame &F, std::map<size_t, MapPoint *> &matches)
    {
        const FeatureViews &KFviews = pKF->getViews();
        const FeatureViews &Fviews = F.getViews();
        std::vector<uint8_t> ok1(KFviews.numViews(), 0), ok2(Fviews.numViews(), 1);
        for (int i = 0; i < KFviews.numViews(); i++) ok1[i] = hasGoodLandmark(pKF, i);                    // PreviouslyMatchedIndexCriterion(true)
        std::vector<std::pair<size_t, size_t>> pairs = searchByBoW(pKF->mFeatVec, F.mFeatVec, KFviews, Fviews, ok1, ok2, nullptr, TH_LOW, mfNNratio);
        for (const auto &p : pairs) matches[p.second] = pKF->hasAssociation((int)p.first);               // ascending idx_kf, later entries overwrite
        return (int)pairs.size();
    }

    // ---- FeatureMatcher.h:136 -- new landmarks for triangulation (FeatureMatcher.cc:373-402)
    int SearchForTriangulation(KeyFrame *pKF1, KeyFrame *pKF2, cv::Mat F12, std::vector<std::pair<size_t, size_t>> &vMatchedPairs, const bool bOnlyStereo)
    {
        const FeatureViews &v1 = pKF1->getViews();
        const FeatureViews &v2 = pKF2->getViews();
        std::vector<uint8_t> ok1(v1.numViews()), ok2(v2.numViews());
        const bool st1 = bOnlyStereo && pKF1->getCamera().sensor != 0, st2 = bOnlyStereo && pKF2->getCamera().sensor != 0;   // StereoIndexCriterion skips mono cameras
        for (int i = 0; i < v1.numViews(); i++) ok1[i] = !hasGoodLandmark(pKF1, i) && (!st1 || v1.uR(i) >= 0);               // PreviouslyMatchedIndexCriterion(false)
        for (int i = 0; i < v2.numViews(); i++) ok2[i] = !hasGoodLandmark(pKF2, i) && (!st2 || v2.uR(i) >= 0);
        std::vector<std::pair<size_t, size_t>> pairs = searchByBoW(pKF1->mFeatVec, pKF2->mFeatVec, v1, v2, ok1, ok2, &F12, TH_LOW, 1.000f);
        for (const auto &p : pairs) vMatchedPairs.push_back(p);
        return (int)pairs.size();
    }

    // ---- FeatureMatcher.h:140 -- monocular map initialisation (FeatureMatcher.cc:404-462)
    int SearchForInitialization(Frame &F1, Frame &F2, std::vector<cv::Point2f> &vbPrevMatched, std::vector<int> &vnMatches12, int windowSize = 10)
    {
        const FeatureViews &v1 = F1.getViews();
        const FeatureViews &v2 = F2.getViews();
        const std::vector<cv::KeyPoint> k1 = v1.getKeys(), k2 = v2.getKeys();
        const std::vector<uint8_t> d1 = cuda_marshal::packDescriptors(v1.getDescriptors()), d2 = cuda_marshal::packDescriptors(v2.getDescriptors());
        const int n1 = (int)k1.size(), n2 = (int)k2.size();
        std::vector<float> prev(2 * (size_t)n1);
        for (int i = 0; i < n1; i++) { prev[2 * i] = vbPrevMatched[i].x; prev[2 * i + 1] = vbPrevMatched[i].y; }
        std::vector<int32_t> m12(n1, -1);
        int32_t nm = 0;
        const Camera &cam = F2.getCamera();
        hyorb_bounds b{cam.mnMinX, cam.mnMaxX, cam.mnMinY, cam.mnMaxY};
        cuda_marshal::check(hyorb_search_for_initialization_host(m, cuda_marshal::asAbi(k1), d1.data(), n1, cuda_marshal::asAbi(k2), d2.data(), n2, b, prev.data(),
                                                                 windowSize, TH_LOW, mfNNratio, m12.data(), &nm));
        vnMatches12 = std::vector<int>(n1, -1);
        for (int i = 0; i < n1; i++)
            if (m12[i] >= 0) { vnMatches12[i] = m12[i]; vbPrevMatched[i] = cv::Point2f(prev[2 * i], prev[2 * i + 1]); }
        return nm;
    }

    // ---- FeatureMatcher.h:143 -- duplicated landmarks (FeatureMatcher.cc:464-521)
    int Fuse(KeyFrame *pKF, const std::vector<MapPoint *> &vpMapPoints, std::map<std::size_t, MapPoint *> &fuse_matches, const float th = 3.0,
             const float reprojection_err = 5.99)
    {
        std::vector<MapPoint *> cand;
        for (MapPoint *lm : vpMapPoints)                                     // pre-screen, :480-487
            if (lm && !(lm->isBad() || lm->IsInKeyFrame(pKF) || lm->Protected())) cand.push_back(lm);
        const int n = (int)cand.size();
        if (n == 0) return (int)fuse_matches.size();
        const FeatureViews &views = pKF->getViews();
        const FeatureExtractorSettings orb_params = views.orbParams();
        const hyorb_projection pr = projectionOf(pKF->GetRotation(), pKF->GetTranslation(), pKF->GetCameraCenter(), pKF->getCamera());
        std::vector<hyorb_landmark> lms(n);
        std::vector<float> normals(3 * (size_t)n);
        std::vector<uint8_t> lm_desc((size_t)n * HYORB_DESC_BYTES);
        for (int i = 0; i < n; i++) {
            lms[i] = landmarkOf(cand[i], pKF->hasAssociation(cand[i]));
            const cv::Mat nrm = cand[i]->GetNormal();
            for (int k = 0; k < 3; k++) normals[3 * i + k] = nrm.at<float>(k);
            packOne(cand[i]->GetDescriptor(), lm_desc.data() + (size_t)i * HYORB_DESC_BYTES);
        }
        const std::vector<cv::KeyPoint> keys = views.getKeys();
        const std::vector<uint8_t> t_desc = cuda_marshal::packDescriptors(views.getDescriptors());
        const std::vector<float> uR = viewURs(views);
        std::vector<int32_t> bi(n, -1);
        std::vector<uint16_t> b(n), s(n);
        std::vector<uint8_t> acc(n, 0);
        cuda_marshal::check(hyorb_fuse_host(m, &pr, lms.data(), normals.data(), lm_desc.data(), n, cuda_marshal::asAbi(keys), t_desc.data(),
                                            uR.empty() ? nullptr : uR.data(), (int)keys.size(), th, orb_params.size_ref, orb_params.sigma_ref, reprojection_err,
                                            std::cos(1.047f) /* ViewingAngleCriterion(1.047): cos(float) */, TH_LOW, 1.000f, bi.data(), b.data(), s.data(), acc.data(),
                                            nullptr));
        for (int i = 0; i < n; i++)
            if (acc[i]) fuse_matches.insert(std::make_pair((size_t)bi[i], cand[i]));
        return (int)fuse_matches.size();
    }

    // ---- FeatureMatcher.h:150 -- loop closing (FeatureMatcher.cc:739-937)
    int SearchBySim3(KeyFrame *pKF1, KeyFrame *pKF2, std::vector<MapPoint *> &vpMatches12, const float &s12, const cv::Mat &R12, const cv::Mat &t12, const float th)
    {
        const cv::Mat R1w = pKF1->GetRotation(), t1w = pKF1->GetTranslation(), R2w = pKF2->GetRotation(), t2w = pKF2->GetTranslation();
        const cv::Mat sR12 = s12 * R12;                                      // the reference's own cv::Mat expressions, :753-756
        const cv::Mat sR21 = (1.0 / s12) * R12.t();
        const cv::Mat t21 = -sR21 * t12;
        const std::vector<MapPoint *> vpMapPoints1 = pKF1->GetMapPointMatches(), vpMapPoints2 = pKF2->GetMapPointMatches();
        const int N1 = (int)vpMapPoints1.size(), N2 = (int)vpMapPoints2.size();
        std::vector<bool> matched1(N1, false), matched2(N2, false);
        for (int i = 0; i < N1; i++) {
            MapPoint *pMP = vpMatches12[i];
            if (pMP) {
                matched1[i] = true;
                const int idx2 = pMP->GetIndexInKeyFrame(pKF2);
                if (idx2 >= 0 && idx2 < N2) matched2[idx2] = true;
            }
        }
        const std::vector<int> vnMatch1 = sim3Direction(pKF1, vpMapPoints1, matched1, R1w, t1w, sR21, t21, pKF2, th);
        const std::vector<int> vnMatch2 = sim3Direction(pKF2, vpMapPoints2, matched2, R2w, t2w, sR12, t12, pKF1, th);
        int nFound = 0;
        for (int i1 = 0; i1 < N1; i1++) {
            const int idx2 = vnMatch1[i1];
            if (idx2 >= 0 && vnMatch2[idx2] == i1) { vpMatches12[i1] = vpMapPoints2[idx2]; nFound++; }
        }
        return nFound;
    }

    static const int HISTO_LENGTH = 30;

protected:
    float mfNNratio;
    bool mbCheckOrientation;
    float TH_LOW = 50.0f;        // FeatureMatcherSettings defaults (FeatureMatcher.h:98-103); the (nnratio, checkOri) constructor of the reference leaves them unset
    float TH_HIGH = 100.0f;

private:
    hyorb_matcher *m = nullptr;
    int dev = 0;

    static void copy3x3(const cv::Mat &M, float *out) { for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) out[3 * r + c] = M.at<float>(r, c); }
    static void copy3(const cv::Mat &v, float *out) { for (int k = 0; k < 3; k++) out[k] = v.at<float>(k); }
    static hyorb_projection projectionOf(const cv::Mat &Rcw, const cv::Mat &tcw, const cv::Mat &Ow, const Camera &cam)
    {
        hyorb_projection pr;
        copy3x3(Rcw, pr.Rcw); copy3(tcw, pr.tcw); copy3(Ow, pr.Ow); copy3x3(cam.K, pr.K);
        pr.mbf = cam.mbf; pr.stereo = cam.sensor == 1 ? 1 : 0;
        pr.bounds.min_x = cam.mnMinX; pr.bounds.max_x = cam.mnMaxX; pr.bounds.min_y = cam.mnMinY; pr.bounds.max_y = cam.mnMaxY;
        return pr;
    }
    static hyorb_landmark landmarkOf(MapPoint *lm, int assoc_idx)
    {
        hyorb_landmark L;
        const cv::Mat P = lm->GetWorldPos();
        copy3(P, L.Pw);
        L.size = lm->getSize(); L.min_dist = lm->GetMinDistanceInvariance(); L.max_dist = lm->GetMaxDistanceInvariance(); L.assoc_idx = assoc_idx;
        return L;
    }
    static void packOne(const FeatureDescriptor &d, uint8_t *out)
    {
        const cv::Mat r = d.rawDescriptor();
        std::memcpy(out, r.data, HYORB_DESC_BYTES);
    }
    // FeatureViews::uR(i) for every view: empty for monocular views (the ABI then treats every uR as -1)
    static std::vector<float> viewURs(const FeatureViews &views)
    {
        if (!views.isStereo()) return std::vector<float>();
        return views.getuRs();
    }
    static uint8_t hasGoodLandmark(KeyFrame *pKF, int idx)                   // PreviouslyMatchedIndexCriterion's test, MatchCriteria.cpp:556-566
    {
        MapPoint *lm = pKF->hasAssociation(idx);
        return (uint8_t)(lm && !lm->isBad());
    }

    // _SearchByProjection_ (FeatureMatcher.cc:57-121) for the three Frame variants
    int searchByProjection(Frame &frame, const std::vector<MapPoint *> &landmarks, float th, unsigned flags, float thr, float ratio, const Frame *prev)
    {
        // the reference keeps its matches in a std::map<MapPoint*, ...>: landmarks in ascending pointer order, one entry per landmark
        std::vector<MapPoint *> lmv;
        for (MapPoint *lm : landmarks) if (lm) lmv.push_back(lm);
        std::sort(lmv.begin(), lmv.end());
        lmv.erase(std::unique(lmv.begin(), lmv.end()), lmv.end());
        const int n = (int)lmv.size();
        if (n == 0) return 0;
        const FeatureViews &views = frame.getViews();
        const FeatureExtractorSettings orb_params = views.orbParams();
        const cv::Mat Tcw = frame.mTcw;
        const hyorb_projection pr = projectionOf(Tcw.rowRange(0, 3).colRange(0, 3), Tcw.rowRange(0, 3).col(3), frame.GetCameraCenter(), frame.getCamera());
        std::vector<hyorb_landmark> lms(n);
        std::vector<uint8_t> lm_desc((size_t)n * HYORB_DESC_BYTES);
        std::vector<float> prev_angle;
        if (prev) prev_angle.resize(n);
        for (int i = 0; i < n; i++) {
            lms[i] = landmarkOf(lmv[i], frame.hasAssociation(lmv[i]));
            packOne(lmv[i]->GetDescriptor(), lm_desc.data() + (size_t)i * HYORB_DESC_BYTES);
            if (prev) prev_angle[i] = prev->getViews().keypt(prev->hasAssociation(lmv[i])).angle;       // MatchCriteria.cpp:381
        }
        const std::vector<cv::KeyPoint> keys = views.getKeys();
        const int nt = (int)keys.size();
        const std::vector<uint8_t> t_desc = cuda_marshal::packDescriptors(views.getDescriptors());
        const std::vector<float> uR = viewURs(views);
        std::vector<uint8_t> t_matched(nt, 0);                               // PreviouslyMatchedCriterionCore, MatchCriteria.cpp:124-144
        for (int i = 0; i < nt; i++) { MapPoint *p = frame.hasAssociation(i); t_matched[i] = (uint8_t)(p && p->Observations() > 0); }
        std::vector<int32_t> bi(n, -1);
        std::vector<uint16_t> b(n), s(n);
        std::vector<uint8_t> acc(n, 0);
        if (pr.stereo && (flags & HYORB_SBP_STEREO) && uR.empty()) flags &= ~(unsigned)HYORB_SBP_STEREO;
        cuda_marshal::check(hyorb_search_by_projection_ex_host(m, &pr, lms.data(), lm_desc.data(), prev ? prev_angle.data() : nullptr, n, cuda_marshal::asAbi(keys),
                                                               t_desc.data(), uR.empty() ? nullptr : uR.data(), t_matched.data(), nt, th, orb_params.size_ref, thr,
                                                               ratio, flags, bi.data(), b.data(), s.data(), acc.data(), nullptr));
        int nmatches = 0;
        for (int i = 0; i < n; i++)                                          // association loop in map (= pointer) order, :111-116
            if (acc[i]) { frame.associateLandMark(bi[i], lmv[i], true); nmatches++; }
        return nmatches;
    }

    // _SearchByBoW_ (FeatureMatcher.cc:281-345): candidate lists from the two DBoW2 FeatureVectors, scan (+ epipolar gate) on the device,
    // RotationConsistencyBoW on the device.  Returns the surviving (idx1, idx2) pairs in idx1 order.
    std::vector<std::pair<size_t, size_t>> searchByBoW(const DBoW2::FeatureVector &fv1, const DBoW2::FeatureVector &fv2, const FeatureViews &v1,
                                                        const FeatureViews &v2, const std::vector<uint8_t> &ok1, const std::vector<uint8_t> &ok2,
                                                        const cv::Mat *F12, float thr, float ratio)
    {
        const int n1 = v1.numViews(), n2 = v2.numViews();
        std::vector<std::vector<int32_t>> cand(n1);
        auto f1it = fv1.begin(), f2it = fv2.begin();
        while (f1it != fv1.end() && f2it != fv2.end()) {                     // merge walk of the two node maps
            if (f1it->first == f2it->first) {
                std::vector<int32_t> idx2;
                for (unsigned int j : f2it->second) if (ok2[j]) idx2.push_back((int32_t)j);
                for (unsigned int i : f1it->second) if (ok1[i] && cand[i].empty()) cand[i] = idx2;      // matches_internal.insert keeps the first node's match
                ++f1it; ++f2it;
            } else if (f1it->first < f2it->first) f1it = fv1.lower_bound(f2it->first);
            else f2it = fv2.lower_bound(f1it->first);
        }
        std::vector<int32_t> off(n1 + 1, 0), idx;
        for (int i = 0; i < n1; i++) { idx.insert(idx.end(), cand[i].begin(), cand[i].end()); off[i + 1] = (int32_t)idx.size(); }
        if (idx.empty()) idx.push_back(0);
        const std::vector<cv::KeyPoint> k1 = v1.getKeys(), k2 = v2.getKeys();
        const std::vector<uint8_t> d1 = cuda_marshal::packDescriptors(v1.getDescriptors()), d2 = cuda_marshal::packDescriptors(v2.getDescriptors());
        std::vector<int32_t> bi(n1, -1);
        std::vector<uint16_t> b(n1), s(n1);
        std::vector<uint8_t> acc(n1, 0);
        if (F12) {
            float Fm[9];
            copy3x3(*F12, Fm);
            const FeatureExtractorSettings orb2 = v2.orbParams();
            cuda_marshal::check(hyorb_match_csr_epipolar_host(m, cuda_marshal::asAbi(k1), d1.data(), n1, cuda_marshal::asAbi(k2), d2.data(), n2, off.data(), idx.data(),
                                                              Fm, orb2.sigma_ref, orb2.size_ref, HYORB_RULE_BOW, thr, ratio, bi.data(), b.data(), s.data(), acc.data()));
        } else {
            cuda_marshal::check(hyorb_match_csr_host(m, d1.data(), n1, d2.data(), n2, off.data(), idx.data(), HYORB_RULE_BOW, thr, ratio, bi.data(), b.data(), s.data(),
                                                     acc.data()));
        }
        std::vector<size_t> i1s, i2s;
        std::vector<float> a_prev, a_curr;
        for (int i = 0; i < n1; i++)
            if (acc[i]) { i1s.push_back(i); i2s.push_back(bi[i]); a_curr.push_back(k1[i].angle); a_prev.push_back(k2[bi[i]].angle); }    // rot = angle(views2) - angle(views1)
        std::vector<uint8_t> keep(i1s.size(), 0);
        if (!i1s.empty()) cuda_marshal::check(hyorb_rotation_consistency_host(m, a_prev.data(), a_curr.data(), (int)i1s.size(), keep.data()));
        std::vector<std::pair<size_t, size_t>> out;
        for (size_t k = 0; k < i1s.size(); k++) if (keep[k]) out.push_back(std::make_pair(i1s[k], i2s[k]));
        return out;
    }

    // one direction of SearchBySim3 (:783-845): vnMatch per feature of keyframe A
    std::vector<int> sim3Direction(KeyFrame *pKFa, const std::vector<MapPoint *> &mpsA, const std::vector<bool> &matchedA, const cv::Mat &Ra, const cv::Mat &ta,
                                   const cv::Mat &sRba, const cv::Mat &tba, KeyFrame *pKFb, float th)
    {
        const int NA = (int)mpsA.size();
        std::vector<int> vnMatch(NA, -1), feat;
        for (int i = 0; i < NA; i++) { MapPoint *p = mpsA[i]; if (p && !matchedA[i] && !p->isBad()) feat.push_back(i); }
        const int n = (int)feat.size();
        if (n == 0) return vnMatch;
        const FeatureViews &vb = pKFb->getViews();
        const FeatureExtractorSettings orb_b = vb.orbParams();
        const hyorb_projection prb = projectionOf(pKFb->GetRotation(), pKFb->GetTranslation(), pKFb->GetCameraCenter(), pKFb->getCamera());
        std::vector<hyorb_landmark> lms(n);
        std::vector<uint8_t> lm_desc((size_t)n * HYORB_DESC_BYTES);
        for (int j = 0; j < n; j++) {
            MapPoint *p = mpsA[feat[j]];
            lms[j] = landmarkOf(p, pKFb->hasAssociation(p));
            packOne(p->GetDescriptor(), lm_desc.data() + (size_t)j * HYORB_DESC_BYTES);
        }
        float R_a[9], t_a[3], sR[9], tb[3];
        copy3x3(Ra, R_a); copy3(ta, t_a); copy3x3(sRba, sR); copy3(tba, tb);
        const std::vector<cv::KeyPoint> kb = vb.getKeys();
        const std::vector<uint8_t> db = cuda_marshal::packDescriptors(vb.getDescriptors());
        std::vector<int32_t> bi(n, -1);
        std::vector<uint16_t> b(n);
        std::vector<uint8_t> acc(n, 0);
        cuda_marshal::check(hyorb_search_by_sim3_host(m, R_a, t_a, sR, tb, &prb, lms.data(), lm_desc.data(), n, cuda_marshal::asAbi(kb), db.data(), (int)kb.size(), th,
                                                      orb_b.size_ref, TH_HIGH, bi.data(), b.data(), acc.data(), nullptr));
        for (int j = 0; j < n; j++) if (acc[j]) vnMatch[feat[j]] = bi[j];
        return vnMatch;
    }
};

}  // namespace HYSLAM
