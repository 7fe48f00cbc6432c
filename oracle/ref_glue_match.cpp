// ref_glue_match.cpp -- C entry points over the REFERENCE's own matcher translation units (part of oracle/_ref).
// TEST INFRASTRUCTURE (oracle/), not product code.
//
// Compiled UNMODIFIED from /root/reference by oracle/Makefile: src/features/{FeatureMatcher.cc,MatchCriteria.cpp},
// src/core/{Frame.cc,KeyFrame.cc,MapPoint.cc,Camera.cpp,LandMarkMatches.cpp,SensorData.cpp} (+ the extractor-side files of
// ref_glue.cpp).  This file only BUILDS the reference's objects from flat arrays (a "scene": MapPoints, Frames, KeyFrames)
// and calls the reference's public matcher entry points (FeatureMatcher.h:105-176), then reads the observable results back.
//
// Pointer order.  _SearchByProjection_ keeps its matches in a std::map<MapPoint*, ...> (FeatureMatcher.cc:64), so the
// association loop (:111-116) and RotationConsistencyCriterion's re-keying (MatchCriteria.cpp:363-401) run in POINTER order;
// the relocalisation variant std::sort()s MapPoint* (:188-189).  The scene places all MapPoints in ONE contiguous block, so
// pointer order == landmark id order -- the canonical policy of this project (same idea as the extractor's monotonic arena).
#include <FeatureMatcher.h>
#include <MatchCriteria.h>
#include <Frame.h>
#include <KeyFrame.h>
#include <MapPoint.h>
#include <Camera.h>
#include <FeatureViews.h>
#include <SensorData.h>
#include <GenUtils.h>
#include <Converter.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <new>
#include <set>
#include <vector>

#include "ref_scene.hpp"

// Which matcher class the entry points below drive, and their symbol prefix: the reference's FeatureMatcher as refm_* (oracle/_ref), or --
// compiled a second time by tests/cpp/matcher_shim_test.cpp -- the C++ drop-in CudaFeatureMatcher as shimm_* on the SAME scene objects.
#ifndef REFM_MATCHER
#define REFM_MATCHER FeatureMatcher
#endif
#ifndef REFM_NAME
#define REFM_NAME(x) refm_##x
#endif

#ifndef REFM_ENTRY_POINTS_ONLY
// ---- link stubs for the three symbols the matcher TUs reference from files that cannot be compiled here -------------
// util/GenUtils.cpp and util/Converter.cc need Eigen / g2o.  GenUtils::Epipole is four lines over the reference's own
// KeyFrame::ProjectLandMark (GenUtils.cpp:11-18); the two Converter functions only run when a frame carries IMU data
// (Frame::CalcRelativeQuat, Frame.cc:344-356), which the scenes built here never do.
namespace HYSLAM {
int GenUtils::Epipole(KeyFrame *pKF1, KeyFrame *pKF2, float &ex, float &ey)
{
    cv::Mat uv_ur;
    pKF2->ProjectLandMark(pKF1->GetCameraCenter(), uv_ur);
    ex = uv_ur.at<float>(0, 0);
    ey = uv_ur.at<float>(1, 0);
    return 0;
}
Eigen::Quaterniond Converter::toQuatEigen(const std::vector<double> q) { return Eigen::Quaterniond(q[0], q[1], q[2], q[3]); }
std::vector<double> Converter::toQuatStdvec(Eigen::Quaterniond q) { return std::vector<double>{q.x(), q.y(), q.z(), q.w()}; }
}  // namespace HYSLAM

using namespace HYSLAM;

extern "C" {
#define REF_API __attribute__((visibility("default")))

}

using namespace refm;

extern "C" {

REF_API void *refm_scene_create(int max_mappoints)
{
    Scene *s = new Scene();
    s->cap = max_mappoints;
    s->mp = (MapPoint *)std::malloc(sizeof(MapPoint) * (size_t)(max_mappoints > 0 ? max_mappoints : 1));
    return s;
}
REF_API void refm_scene_destroy(void *h) { delete (Scene *)h; }

// MapPoints: world position, normal, size (world units), RAW mfMinDistance / mfMaxDistance (the getters apply 0.8 / 1.2),
// descriptor, bad flag, protection count.  Returns the id of the first one added.
REF_API int refm_add_mappoints(void *h, int n, const float *Pw, const float *normal, const float *size, const float *min_dist, const float *max_dist,
                               const uint8_t *desc, const uint8_t *bad, const int32_t *n_protected)
{
    Scene *s = (Scene *)h;
    if (s->n_mp + n > s->cap) return -1;
    const int first = s->n_mp;
    for (int i = 0; i < n; i++) {
        MapPoint *p = new (s->mp + s->n_mp) MapPoint(mat_from(Pw + 3 * i, 3, 1));
        s->n_mp++;
        if (normal) p->setNormal(mat_from(normal + 3 * i, 3, 1));
        p->setSize(size ? size[i] : 0.f);
        p->setMinDistanceInvariance(min_dist ? min_dist[i] : 0.f);
        p->setMaxDistanceInvariance(max_dist ? max_dist[i] : 0.f);
        p->setDescriptor(FeatureDescriptor(cv::Mat(1, 32, CV_8UC1, (void *)(desc + 32 * (size_t)i)), s->dist));
        if (bad && bad[i]) p->setBad();
        if (n_protected && n_protected[i] > 0) p->SetProtection(n_protected[i], 0);
    }
    return first;
}

// A Frame (as_keyframe = 0) or a KeyFrame built from it (KeyFrame::KeyFrame(Frame&), KeyFrame.cc:33-72).  assoc[i] = id of
// the MapPoint associated with keypoint i, or -1.  For a KeyFrame every associated MapPoint also gets the observation
// (KeyFrame*, i) and nObs >= 1, the state hySLAM's map keeps (MapPointDB) and the criteria read (IsInKeyFrame, Observations()).
REF_API int refm_add_frame(void *h, const refm_frame_desc *d, const int32_t *assoc, int as_keyframe)
{
    Scene *s = (Scene *)h;
    const int n = d->n;
    std::vector<cv::KeyPoint> keys(n);
    if (n) std::memcpy(keys.data(), d->kps, sizeof(refm_keypoint) * (size_t)n);
    std::vector<FeatureDescriptor> descs;
    descs.reserve(n);
    for (int i = 0; i < n; i++) descs.push_back(FeatureDescriptor(cv::Mat(1, 32, CV_8UC1, (void *)(d->desc + 32 * (size_t)i)), s->dist));
    FeatureExtractorSettings orb;
    orb.nFeatures = n; orb.fScaleFactor = 1.2f; orb.nLevels = 8; orb.init_threshold = 20; orb.min_threshold = 4; orb.N_CELLS = 30;
    orb.size_ref = d->size_ref; orb.sigma_ref = d->sigma_ref;
    const bool stereo = d->uR != nullptr;
    FeatureViews views = stereo
        ? FeatureViews(keys, std::vector<cv::KeyPoint>(), std::vector<float>(d->uR, d->uR + n), std::vector<float>(d->depth, d->depth + n), descs,
                       std::vector<FeatureDescriptor>(), orb)
        : FeatureViews(keys, descs, orb);
    Camera cam;
    cam.K = mat_from(d->K, 3, 3);
    cam.mbf = d->mbf; cam.sensor = d->sensor; cam.RGB = 0; cam.thDepth = 40.f; cam.fps = 10.f; cam.scale = 1.f;
    cam.mnMinX = d->min_x; cam.mnMaxX = d->max_x; cam.mnMinY = d->min_y; cam.mnMaxY = d->max_y;
    std::unique_ptr<Frame> f(new Frame(0.0, views, nullptr, cam, std::string(), SensorData(), stereo));
    f->SetPose(mat_from(d->Tcw, 4, 4));
    if (assoc)
        for (int i = 0; i < n; i++)
            if (assoc[i] >= 0) { if (assoc[i] >= s->n_mp) return -1; f->associateLandMark(i, s->mp + assoc[i], true); }
    std::unique_ptr<KeyFrame> kf;
    if (as_keyframe) {
        kf.reset(new KeyFrame(*f));
        if (assoc)
            for (int i = 0; i < n; i++)
                if (assoc[i] >= 0) {
                    MapPoint *p = s->mp + assoc[i];
                    std::map<KeyFrame *, size_t> obs = p->GetObservations();
                    obs[kf.get()] = (size_t)i;
                    p->setObservations(obs);
                    p->setNObs((int)obs.size());
                }
    }
    s->frames.push_back(std::move(f));
    s->keyframes.push_back(std::move(kf));
    return (int)s->frames.size() - 1;
}

// DBoW2::FeatureVector of a frame / keyframe from a node id per feature (features added in ascending index order, as
// DBoW2's transform does)
REF_API int refm_set_feature_nodes(void *h, int frame, const int32_t *node_of, int n)
{
    Scene *s = (Scene *)h;
    DBoW2::FeatureVector fv;
    for (int i = 0; i < n; i++) if (node_of[i] >= 0) fv.addFeature((DBoW2::NodeId)node_of[i], (unsigned)i);
    s->frames[frame]->mFeatVec = fv;
    if (s->keyframes[frame]) s->keyframes[frame]->mFeatVec = fv;
    return 0;
}

REF_API int refm_set_observation_count(void *h, int mp, int n_obs) { ((Scene *)h)->mp[mp].setNObs(n_obs); return 0; }

// association state of a frame / keyframe: out[i] = MapPoint id of keypoint i or -1
REF_API int refm_get_assoc(void *h, int frame, int32_t *out, int n)
{
    Scene *s = (Scene *)h;
    for (int i = 0; i < n; i++) out[i] = s->id_of(s->keyframes[frame] ? s->keyframes[frame]->hasAssociation(i) : s->frames[frame]->hasAssociation(i));
    return 0;
}

// camera centre as the reference derives it from the pose (Frame::UpdatePoseMatrices, Frame.cc:161-168: mOw = -mRcw.t()*mtcw)
REF_API int refm_camera_center(void *h, int frame, float *Ow)
{
    Scene *s = (Scene *)h;
    const cv::Mat c = s->keyframes[frame] ? s->keyframes[frame]->GetCameraCenter() : s->frames[frame]->GetCameraCenter();
    for (int i = 0; i < 3; i++) Ow[i] = c.at<float>(i);
    return 0;
}

// Frame::GetFeaturesInAreaNEW / KeyFrame::GetFeaturesInArea (Frame.cc:416-457, KeyFrame.cc:329-369)
REF_API int refm_features_in_area(void *h, int frame, float x, float y, float r, int32_t *out, int cap)
{
    Scene *s = (Scene *)h;
    const std::vector<size_t> v = s->keyframes[frame] ? s->keyframes[frame]->GetFeaturesInArea(x, y, r) : s->frames[frame]->GetFeaturesInAreaNEW(x, y, r);
    if ((int)v.size() > cap) return -2;
    for (size_t i = 0; i < v.size(); i++) out[i] = (int32_t)v[i];
    return (int)v.size();
}

// Frame::ProjectLandMark + landMarkSizePixels for one MapPoint: uv_ur[3], size in pixels; returns the visibility flag
REF_API int refm_project(void *h, int frame, int mp, float *uv_ur, float *size_px)
{
    Scene *s = (Scene *)h;
    cv::Mat uv;
    const bool ok = s->frames[frame]->ProjectLandMark(s->mp + mp, uv);
    for (int i = 0; i < 3; i++) uv_ur[i] = uv.at<float>(i, 0);
    *size_px = s->frames[frame]->landMarkSizePixels(s->mp + mp);
    return ok ? 1 : 0;
}

}  // extern "C"
#endif  // REFM_ENTRY_POINTS_ONLY

#ifdef REFM_ENTRY_POINTS_ONLY
using namespace refm;
#define REF_API __attribute__((visibility("default")))
#endif

extern "C" {

// FeatureMatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th)   (FeatureMatcher.cc:123-143), lm_ids[i] < 0 = null entry
REF_API int REFM_NAME(search_by_projection)(void *h, int frame, const int32_t *lm_ids, int n, float th, const refm_settings *st)
{
    Scene *s = (Scene *)h;
    std::vector<MapPoint *> lms(n);
    for (int i = 0; i < n; i++) lms[i] = lm_ids[i] >= 0 ? s->mp + lm_ids[i] : nullptr;
    REFM_MATCHER m(to_settings(st));
    return m.SearchByProjection(*s->frames[frame], lms, th);
}

// FeatureMatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, th, bMono)   (:145-176)
REF_API int REFM_NAME(search_by_projection_motion)(void *h, int cur, int last, float th, int mono, const refm_settings *st)
{
    Scene *s = (Scene *)h;
    REFM_MATCHER m(to_settings(st));
    return m.SearchByProjection(*s->frames[cur], *s->frames[last], th, mono != 0);
}

// FeatureMatcher::SearchByProjection(Frame &CurrentFrame, KeyFrame*, const set<MapPoint*> &sAlreadyFound, th, ORBdist)   (:180-213)
REF_API int REFM_NAME(search_by_projection_reloc)(void *h, int cur, int kf, const int32_t *found, int n_found, float th, int orb_dist, const refm_settings *st)
{
    Scene *s = (Scene *)h;
    std::set<MapPoint *> already;
    for (int i = 0; i < n_found; i++) already.insert(s->mp + found[i]);
    REFM_MATCHER m(to_settings(st));
    return m.SearchByProjection(*s->frames[cur], s->keyframes[kf].get(), already, th, orb_dist);
}

// FeatureMatcher::Fuse(KeyFrame*, const vector<MapPoint*>&, map<size_t, MapPoint*> &fuse_matches, th, reprojection_err)   (:464-521)
// out_idx / out_lm: the (keypoint index, MapPoint id) pairs of fuse_matches in map order
REF_API int REFM_NAME(fuse)(void *h, int kf, const int32_t *lm_ids, int n, float th, float reproj_err, const refm_settings *st, int32_t *out_idx, int32_t *out_lm, int cap)
{
    Scene *s = (Scene *)h;
    std::vector<MapPoint *> lms(n);
    for (int i = 0; i < n; i++) lms[i] = lm_ids[i] >= 0 ? s->mp + lm_ids[i] : nullptr;
    REFM_MATCHER m(to_settings(st));
    std::map<std::size_t, MapPoint *> fm;
    m.Fuse(s->keyframes[kf].get(), lms, fm, th, reproj_err);
    if ((int)fm.size() > cap) return -2;
    int k = 0;
    for (auto &e : fm) { out_idx[k] = (int32_t)e.first; out_lm[k] = s->id_of(e.second); k++; }
    return k;
}

// FeatureMatcher::SearchForInitialization(Frame &F1, Frame &F2, vector<Point2f> &vbPrevMatched, vector<int> &vnMatches12, windowSize)   (:404-462)
REF_API int REFM_NAME(search_for_initialization)(void *h, int f1, int f2, float *prev_matched_xy /* n1 x 2, in/out */, int32_t *matches12 /* n1 */, int window,
                                           const refm_settings *st)
{
    Scene *s = (Scene *)h;
    const int n1 = s->frames[f1]->N;
    std::vector<cv::Point2f> prev(n1);
    for (int i = 0; i < n1; i++) prev[i] = cv::Point2f(prev_matched_xy[2 * i], prev_matched_xy[2 * i + 1]);
    std::vector<int> m12;
    REFM_MATCHER m(to_settings(st));
    const int r = m.SearchForInitialization(*s->frames[f1], *s->frames[f2], prev, m12, window);
    for (int i = 0; i < n1; i++) { matches12[i] = m12[i]; prev_matched_xy[2 * i] = prev[i].x; prev_matched_xy[2 * i + 1] = prev[i].y; }
    return r;
}

// FeatureMatcher::SearchBySim3(pKF1, pKF2, vpMatches12, s12, R12, t12, th)   (:739-937); matches12: MapPoint ids (in: already matched, out)
REF_API int REFM_NAME(search_by_sim3)(void *h, int kf1, int kf2, int32_t *matches12, int n1, float s12, const float *R12, const float *t12, float th,
                                const refm_settings *st)
{
    Scene *s = (Scene *)h;
    std::vector<MapPoint *> m12(n1);
    for (int i = 0; i < n1; i++) m12[i] = matches12[i] >= 0 ? s->mp + matches12[i] : nullptr;
    REFM_MATCHER m(to_settings(st));
    const int r = m.SearchBySim3(s->keyframes[kf1].get(), s->keyframes[kf2].get(), m12, s12, mat_from(R12, 3, 3), mat_from(t12, 3, 1), th);
    for (int i = 0; i < n1; i++) matches12[i] = s->id_of(m12[i]);
    return r;
}

// FeatureMatcher::SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo)   (:373-402)
REF_API int REFM_NAME(search_for_triangulation)(void *h, int kf1, int kf2, const float *F12, int only_stereo, const refm_settings *st, int32_t *out_i1, int32_t *out_i2, int cap)
{
    Scene *s = (Scene *)h;
    std::vector<std::pair<size_t, size_t>> pairs;
    REFM_MATCHER m(to_settings(st));
    m.SearchForTriangulation(s->keyframes[kf1].get(), s->keyframes[kf2].get(), mat_from(F12, 3, 3), pairs, only_stereo != 0);
    if ((int)pairs.size() > cap) return -2;
    for (size_t i = 0; i < pairs.size(); i++) { out_i1[i] = (int32_t)pairs[i].first; out_i2[i] = (int32_t)pairs[i].second; }
    return (int)pairs.size();
}

// FeatureMatcher::SearchByBoW(KeyFrame *pKF, Frame &F, map<size_t, MapPoint*> &matches)   (:216-280): (F keypoint index, MapPoint id) pairs
REF_API int REFM_NAME(search_by_bow)(void *h, int kf, int frame, const refm_settings *st, int32_t *out_idx, int32_t *out_lm, int cap)
{
    Scene *s = (Scene *)h;
    std::map<size_t, MapPoint *> matches;
    REFM_MATCHER m(to_settings(st));
    m.SearchByBoW(s->keyframes[kf].get(), *s->frames[frame], matches);
    if ((int)matches.size() > cap) return -2;
    int k = 0;
    for (auto &e : matches) { out_idx[k] = (int32_t)e.first; out_lm[k] = s->id_of(e.second); k++; }
    return k;
}

}  // extern "C"
