// ref_scene.hpp -- the "scene" behind oracle/ref_glue_match.cpp: MapPoints, Frames and KeyFrames of the REFERENCE'S OWN classes, built
// from flat arrays.  TEST INFRASTRUCTURE.  Shared with tests/cpp/matcher_shim_test.cpp, which runs the C++ drop-in matcher
// (include/hyorb_hyslam_matcher.hpp) on the very same objects.
#pragma once
#include <FeatureMatcher.h>
#include <Frame.h>
#include <KeyFrame.h>
#include <MapPoint.h>
#include <cstdlib>
#include <memory>
#include <vector>

extern "C" {
struct refm_keypoint { float x, y, size, angle, response; int32_t octave, class_id; };
struct refm_frame_desc {
    int32_t n;
    const refm_keypoint *kps;
    const uint8_t *desc;            // n x 32
    const float *uR, *depth;        // null = monocular views
    float K[9];                     // row-major 3x3
    float mbf;
    int32_t sensor;                 // Camera::sensor: 0 mono, 1 stereo
    float min_x, max_x, min_y, max_y;
    float Tcw[16];                  // row-major 4x4
    float size_ref, sigma_ref;      // FeatureExtractorSettings of the views
};
struct refm_settings { float nnratio, th_high, th_low; int32_t check_ori; };
}

namespace refm {
using namespace HYSLAM;
struct Scene {
    int cap = 0, n_mp = 0;
    MapPoint *mp = nullptr;         // ONE block: id order == pointer order
    std::shared_ptr<DescriptorDistance> dist = std::make_shared<ORBDistance>();
    std::vector<std::unique_ptr<Frame>> frames;
    std::vector<std::unique_ptr<KeyFrame>> keyframes;       // same index as frames; null when the entry is a plain Frame
    ~Scene()
    {
        keyframes.clear(); frames.clear();
        for (int i = 0; i < n_mp; i++) mp[i].~MapPoint();
        std::free(mp);
    }
    int id_of(MapPoint *p) const { return p ? (int)(p - mp) : -1; }
};

inline FeatureMatcherSettings to_settings(const refm_settings *s)
{
    FeatureMatcherSettings m;
    m.nnratio = s->nnratio; m.TH_HIGH = s->th_high; m.TH_LOW = s->th_low; m.checkOri = s->check_ori != 0;
    return m;
}
inline cv::Mat mat_from(const float *v, int r, int c)
{
    cv::Mat m(r, c, CV_32F);
    for (int i = 0; i < r; i++) for (int j = 0; j < c; j++) m.at<float>(i, j) = v[i * c + j];
    return m;
}
}  // namespace refm
