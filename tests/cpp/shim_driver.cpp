// Runs the C++ drop-in (include/hyorb_hyslam.hpp) through the call sequence of ImageProcessing::ProcessStereoImage
// (hySLAM src/main/ImageProcessing.cpp:69-116) and dumps what it produced; tests/test_gpu_cpp_shim.py compares the
// dump with the CPU oracle.  Compiled against the REAL hySLAM headers of /root/reference (+ oracle/cvshim for OpenCV) when that tree is present, else against
// the test doubles in tests/cpp/mock_hyslam (tests/cpp/Makefile).
//   shim_driver left.raw right.raw W H nFeatures mbf fx out.bin
#include <hyorb_hyslam.hpp>

#include <cstdio>
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <thread>

using namespace HYSLAM;

static cv::Mat read_raw(const char *path, int w, int h)
{
    cv::Mat m(h, w, CV_8UC1);
    FILE *f = fopen(path, "rb");
    if (!f || fread(m.data, 1, (size_t)w * h, f) != (size_t)w * h) { fprintf(stderr, "cannot read %s\n", path); exit(2); }
    fclose(f);
    return m;
}
static void extractFeatures(FeatureExtractor *ex, cv::Mat &img, std::vector<cv::KeyPoint> &k, std::vector<FeatureDescriptor> &d)
{
    (*ex)(img, cv::Mat(), k, d);          // FeatureUtil::extractFeatures, called on a transient thread at ImageProcessing.cpp:82
}
template <typename T> static void put(FILE *f, const std::vector<T> &v) { if (!v.empty()) fwrite(v.data(), sizeof(T), v.size(), f); }

int main(int argc, char **argv)
{
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);      // hardware queues for the lanes' streams, before the CUDA context exists

    if (argc < 9) { fprintf(stderr, "usage: shim_driver left.raw right.raw W H nFeatures mbf fx out.bin\n"); return 2; }
    const int w = atoi(argv[3]), h = atoi(argv[4]), nf = atoi(argv[5]);
    const float mbf = (float)atof(argv[6]), fx = (float)atof(argv[7]);
    try {
        cv::Mat mImGray = read_raw(argv[1], w, h), imGrayRight = read_raw(argv[2], w, h);
        CudaORBFactory factory;                                   // System.cc:77-85 would pick this for `Features: ORB`
        FeatureExtractorSettings s = factory.getFeatureExtractorSettings();
        s.nFeatures = nf;
        std::shared_ptr<FeatureExtractor> extractor_left = factory.getExtractor(s), extractor_right = factory.getExtractor(s);

        // ImageProcessing.cpp:31-36 asks the factory per camera TYPE; the SLAM and Imaging blocks of the settings differ
        {
            std::shared_ptr<FeatureExtractor> imaging = factory.getExtractor(std::string("Imaging"));
            const FeatureExtractorSettings si = factory.getFeatureExtractorSettings();
            std::shared_ptr<FeatureExtractor> slam = factory.getExtractor(std::string("SLAM"));
            const FeatureExtractorSettings ss = factory.getFeatureExtractorSettings();
            if (!imaging || !slam || imaging->GetScaleFactor() != 1.4f || slam->GetScaleFactor() != 1.2f || si.nFeatures != 3000 || ss.nFeatures != 1000 ||
                !dynamic_cast<CudaORBExtractor *>(imaging.get())) { fprintf(stderr, "per-camera-type settings are not honoured\n"); return 1; }
            printf("camera types ok\n");
        }

        std::vector<cv::KeyPoint> mvKeys, mvKeysRight;
        std::vector<FeatureDescriptor> mDescriptors, mDescriptorsRight;
        std::thread orb_thread(extractFeatures, extractor_left.get(), std::ref(mImGray), std::ref(mvKeys), std::ref(mDescriptors));
        (*extractor_right)(imGrayRight, cv::Mat(), mvKeysRight, mDescriptorsRight);
        orb_thread.join();
        FeatureExtractorSettings orb_params = FeatureExtractorSettings();      // default-constructed, as at ImageProcessing.cpp:85 (only size_ref = 31 is read)

        FeatureViews LMviews(mvKeys, mvKeysRight, mDescriptors, mDescriptorsRight, orb_params);
        Camera cam;
        cam.K = cv::Mat(3, 3, CV_32F);
        for (int i = 0; i < 9; i++) cam.K.ptr<float>()[i] = 0.f;
        cam.K.at<float>(0, 0) = fx; cam.K.at<float>(1, 1) = fx; cam.K.at<float>(2, 2) = 1.f;
        cam.mbf = mbf; cam.mnMaxX = (float)w; cam.mnMaxY = (float)h;
        CudaStereomatcher stereomatch(LMviews, cam, factory.getFeatureMatcherSettings());
        stereomatch.computeStereoMatches();
        stereomatch.getData(LMviews);

        // SearchForTriangulation-style scan of the left descriptors against the right ones (all targets, BoW rule)
        CudaDescriptorScan scan;
        const FeatureMatcherSettings ms = factory.getFeatureMatcherSettings();
        CudaDescriptorScan::Result r = scan.scan(mDescriptors, mDescriptorsRight, nullptr, nullptr, HYORB_RULE_BOW, ms.TH_LOW, ms.nnratio);

        // representative descriptor of a few synthetic landmarks: landmark l observed by left keypoints 3l, 3l+1, 3l+2 and right keypoint l
        std::vector<std::vector<FeatureDescriptor>> observations;
        for (size_t l = 0; 3 * l + 2 < mDescriptors.size() && l < mDescriptorsRight.size() && l < 50; l++)
            observations.push_back({mDescriptors[3 * l], mDescriptors[3 * l + 1], mDescriptors[3 * l + 2], mDescriptorsRight[l]});
        const std::vector<int32_t> distinctive = scan.distinctiveDescriptors(observations);

        // PreProcessImg + extraction on a camera frame: a BGR frame with B = G = R = the left gray image converts back to exactly
        // that image ((g * 32768 + 16384) >> 15 = g), so the result must equal the left extraction above
        {
            cv::Mat bgr(h, w, CV_8UC3);
            for (int y = 0; y < h; y++)
                for (int x = 0; x < w; x++)
                    for (int c = 0; c < 3; c++) bgr.ptr<unsigned char>(y)[3 * x + c] = mImGray.at<unsigned char>(y, x);
            cv::Mat gray2;
            std::vector<cv::KeyPoint> k2;
            std::vector<FeatureDescriptor> d2;
            static_cast<CudaORBExtractor *>(extractor_left.get())->extractFromCameraFrame(bgr, false, 1.0f, gray2, k2, d2);
            bool same = gray2.rows == h && gray2.cols == w && k2.size() == mvKeys.size();
            for (int y = 0; same && y < h; y++) same = memcmp(gray2.ptr<unsigned char>(y), mImGray.ptr<unsigned char>(y), (size_t)w) == 0;
            same = same && (k2.empty() || memcmp(k2.data(), mvKeys.data(), k2.size() * sizeof(cv::KeyPoint)) == 0);
            same = same && cuda_marshal::packDescriptors(d2) == cuda_marshal::packDescriptors(mDescriptors);
            if (!same) { fprintf(stderr, "camera-frame path differs from the gray path\n"); return 1; }
            printf("camera frame ok\n");
        }

        // the whole of ProcessStereoImage as one device call must give what the three-object path above gave
        {
            CudaStereoFrontEnd front(factory.getDistanceFunc(), s);
            FeatureViews fused = front.processStereoImage(mImGray, imGrayRight, cam, factory.getFeatureMatcherSettings(), orb_params);
            const std::vector<cv::KeyPoint> fk = fused.getKeys(), fkr = fused.getKeysR();
            bool same = fk.size() == mvKeys.size() && fkr.size() == mvKeysRight.size();
            same = same && (fk.empty() || memcmp(fk.data(), mvKeys.data(), fk.size() * sizeof(cv::KeyPoint)) == 0);
            same = same && (fkr.empty() || memcmp(fkr.data(), mvKeysRight.data(), fkr.size() * sizeof(cv::KeyPoint)) == 0);
            same = same && cuda_marshal::packDescriptors(fused.getDescriptors()) == cuda_marshal::packDescriptors(mDescriptors);
            same = same && cuda_marshal::packDescriptors(fused.getDescriptorsR()) == cuda_marshal::packDescriptors(mDescriptorsRight);
            const std::vector<float> fu = fused.getuRs(), fd = fused.getDepths(), u0 = LMviews.getuRs(), d0 = LMviews.getDepths();
            same = same && fu.size() == u0.size() && fd.size() == d0.size();
            same = same && (fu.empty() || (memcmp(fu.data(), u0.data(), fu.size() * sizeof(float)) == 0 && memcmp(fd.data(), d0.data(), fd.size() * sizeof(float)) == 0));
            if (!same) { fprintf(stderr, "fused stereo front end differs from the three-object path\n"); return 1; }
            printf("fused stereo ok\n");
            // per-pair latency of both forms as hySLAM would drive them (informational)
            const int reps = 20;
            auto t0 = std::chrono::steady_clock::now();
            for (int r = 0; r < reps; r++) {
                std::vector<cv::KeyPoint> k1, k2;
                std::vector<FeatureDescriptor> d1, d2;
                std::thread th(extractFeatures, extractor_left.get(), std::ref(mImGray), std::ref(k1), std::ref(d1));
                (*extractor_right)(imGrayRight, cv::Mat(), k2, d2);
                th.join();
                FeatureViews v(k1, k2, d1, d2, orb_params);
                CudaStereomatcher sm(v, cam, factory.getFeatureMatcherSettings());
                sm.computeStereoMatches();
                sm.getData(v);
            }
            auto t1 = std::chrono::steady_clock::now();
            for (int r = 0; r < reps; r++) (void)front.processStereoImage(mImGray, imGrayRight, cam, factory.getFeatureMatcherSettings(), orb_params);
            auto t2 = std::chrono::steady_clock::now();
            printf("latency per pair: three objects %.3f ms, fused %.3f ms\n", std::chrono::duration<double, std::milli>(t1 - t0).count() / reps,
                   std::chrono::duration<double, std::milli>(t2 - t1).count() / reps);
        }

        // SearchForTriangulation-style scan: the first (up to) 300 left keypoints against ALL right keypoints as one "node", behind the
        // epipolar gate of a rectified pair (F12 = [0 0 0; 0 0 -1; 0 1 0]: the line of (x1, y1) is y2 = y1)
        const int n_tri = (int)std::min<size_t>(300, mvKeys.size());
        std::vector<int32_t> tri_off(mvKeys.size() + 1, 0), tri_idx;
        for (size_t i = 0; i < mvKeys.size(); i++) {
            if ((int)i < n_tri) for (size_t j = 0; j < mvKeysRight.size(); j++) tri_idx.push_back((int32_t)j);
            tri_off[i + 1] = (int32_t)tri_idx.size();
        }
        if (tri_idx.empty()) tri_idx.push_back(0);
        cv::Mat F12(3, 3, CV_32F);
        for (int i = 0; i < 9; i++) F12.ptr<float>()[i] = 0.f;
        F12.at<float>(1, 2) = -1.f; F12.at<float>(2, 1) = 1.f;
        FeatureViews viewsL(mvKeys, mDescriptors, orb_params), viewsR(mvKeysRight, mDescriptorsRight, orb_params);
        CudaDescriptorScan::Result tri = scan.searchForTriangulation(viewsL, viewsR, tri_off.data(), tri_idx.data(), F12, ms.TH_LOW);

        FILE *f = fopen(argv[8], "wb");
        if (!f) { fprintf(stderr, "cannot write %s\n", argv[8]); return 2; }
        const int32_t hdr[4] = {(int32_t)mvKeys.size(), (int32_t)mvKeysRight.size(), extractor_left->GetLevels(), (int32_t)sizeof(cv::KeyPoint)};
        fwrite(hdr, sizeof(hdr), 1, f);
        put(f, mvKeys);
        put(f, cuda_marshal::packDescriptors(mDescriptors));
        put(f, mvKeysRight);
        put(f, cuda_marshal::packDescriptors(mDescriptorsRight));
        put(f, LMviews.getuRs());
        put(f, LMviews.getDepths());
        put(f, r.best_idx); put(f, r.best); put(f, r.second); put(f, r.accepted);
        put(f, extractor_left->GetScaleFactors());
        const int32_t nd = (int32_t)distinctive.size();
        fwrite(&nd, sizeof(nd), 1, f);
        put(f, distinctive);
        const int32_t ntri = n_tri;
        fwrite(&ntri, sizeof(ntri), 1, f);
        put(f, tri.best_idx); put(f, tri.best); put(f, tri.second); put(f, tri.accepted);
        fclose(f);
        // a FeatureDescriptor built by the shim behaves like the reference's (ORBDistance through the stored functor)
        if (mDescriptors.size() > 1) printf("distance(desc0, desc1) = %.0f\n", mDescriptors[0].distance(mDescriptors[1]));
        printf("shim ok: %zu + %zu keypoints\n", mvKeys.size(), mvKeysRight.size());
    } catch (const std::exception &e) {
        fprintf(stderr, "shim_driver: %s\n", e.what());
        return 1;
    }
    return 0;
}
