/*
 * hyorb.h -- C ABI of libhyorb: a B200-native (sm_100a) ORB front end that is a drop-in for the
 * ORB extraction / stereo association / descriptor matching path of bmhopkinson/hyslam.
 *
 * Each entry point cites the reference interface it replaces (paths relative to the hySLAM tree).
 * Rules of the boundary:
 *   - plain pointers and sizes only; no C++ / OpenCV / torch types cross it;
 *   - every function returns an int status (HYORB_OK == 0, negative == error); nothing throws;
 *     hyorb_last_error() returns a thread-local message for the last failing call on this thread;
 *   - a handle is single-threaded; distinct handles are independent (own CUDA stream + workspace);
 *     there is no global mutable state (the reference mutates a shared factory);
 *   - capacity overflow is reported (HYORB_ECAPACITY), never truncated silently;
 *   - there is NO CPU fallback: without a CUDA device every compute call fails with HYORB_ECUDA.
 *   - "_host" entry points take host memory and stage H2D/D2H themselves (synchronous);
 *     "_device" entry points take device pointers, enqueue on the handle's stream and do not sync.
 */
#ifndef HYORB_H
#define HYORB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define HYORB_API
#else
#define HYORB_API __attribute__((visibility("default")))
#endif

enum {
    HYORB_OK = 0,
    HYORB_EINVAL = -1,       /* bad argument */
    HYORB_ECAPACITY = -2,    /* an output or internal buffer was too small (nothing is truncated silently) */
    HYORB_EUNSUPPORTED = -3, /* shape / parameter outside what the kernels implement (see DESIGN.md limits) */
    HYORB_ENOMEM = -4,       /* host or device allocation failed */
    HYORB_ECUDA = -5         /* CUDA runtime error, or no CUDA device */
};

#define HYORB_MAX_LEVELS 16
#define HYORB_DESC_BYTES 32
#define HYORB_GRID_COLS 64 /* FRAME_GRID_COLS, src/core/Frame.h:70 */
#define HYORB_GRID_ROWS 48 /* FRAME_GRID_ROWS, src/core/Frame.h:69 */

/* Same memory layout as cv::KeyPoint (28 bytes): a std::vector<cv::KeyPoint>::data() can be passed. */
typedef struct hyorb_keypoint {
    float x, y;      /* pt, in level-0 pixel coordinates (ORBExtractor.cpp:546-552) */
    float size;      /* (float)(int)(31 * scale[octave])  (ORBExtractor.cpp:478)   */
    float angle;     /* degrees, cv::fastAtan2 of the intensity centroid (ORBFinder.cpp:16-43) */
    float response;  /* FAST score */
    int32_t octave;
    int32_t class_id; /* always -1 */
} hyorb_keypoint;

/* Mirrors HYSLAM::FeatureExtractorSettings (src/core/FeatureExtractorSettings.h:19-33) as parsed by
 * ORBFactory::LoadSettings (src/features/ORBFactory.cpp:55-71). */
typedef struct hyorb_extractor_params {
    int32_t nfeatures;   /* N_Features */
    float scale_factor;  /* scale_factor (float in the reference) */
    int32_t nlevels;     /* N_Levels, 1..HYORB_MAX_LEVELS */
    int32_t cell_px;     /* N_Cells: really the FAST cell edge in px (ORBExtractor.cpp:409) */
    int32_t ini_th;      /* threshold_init -- accepted and ignored: the reference's FAST threshold is stuck */
    int32_t min_th;      /* threshold_min  -- at 20 (ORBFinder.cpp:58-60, ORBFinder.h:92)                   */
    uint32_t flags;      /* reserved, must be 0 (== reproduce the reference exactly) */
} hyorb_extractor_params;

typedef struct hyorb_extractor hyorb_extractor;

/* ----- ORB extraction: replaces HYSLAM::ORBExtractor (src/features/ORBExtractor.h:63-116) ----- */

/* ORBExtractor::ORBExtractor (ORBExtractor.cpp:76-119).  `device` is a CUDA ordinal.
 * `cuda_stream` is a cudaStream_t to enqueue on, or NULL to let the handle create its own. */
HYORB_API int hyorb_extractor_create(const hyorb_extractor_params *params, int device, void *cuda_stream,
                                     hyorb_extractor **out);
HYORB_API int hyorb_extractor_destroy(hyorb_extractor *h);

/* FeatureExtractor::GetLevels / GetScaleFactors / GetInverseScaleFactors / GetScaleSigmaSquares /
 * GetInverseScaleSigmaSquares (src/features/FeatureExtractor.h:29-35).  Each array: nlevels floats
 * (NULL to skip).  quota: per-level feature quota mnFeaturesPerLevel (ORBExtractor.cpp:104-117). */
HYORB_API int hyorb_extractor_get_levels(const hyorb_extractor *h);
HYORB_API int hyorb_extractor_get_scales(const hyorb_extractor *h, float *scale, float *inv_scale,
                                         float *sigma2, float *inv_sigma2, int32_t *quota);

/* ORBExtractor::operator()(image, mask, keypoints, descriptors) (ORBExtractor.cpp:496-562) on ONE
 * 8-bit gray HOST image (`stride` bytes per row).  The mask is ignored by the reference and has no
 * parameter here.  Writes *n keypoints (levels concatenated 0..L-1, reference order) and n*32
 * descriptor bytes.  An empty image (NULL / w<=0 / h<=0) returns HYORB_OK with *n = 0 (:499-500). */
HYORB_API int hyorb_extract_host(hyorb_extractor *h, const uint8_t *gray, int width, int height, int stride,
                                 hyorb_keypoint *kps, uint8_t *desc, int capacity, int *n);

/* Throughput form of the same operator over n_images same-sized images (the offline / batched path:
 * one launch sequence covers every level of every image).  Image i starts at images + i*image_stride.
 * Outputs: kps[i*capacity + k], desc[(i*capacity + k)*32], counts[i]. */
HYORB_API int hyorb_extract_batch_host(hyorb_extractor *h, const uint8_t *images, int n_images, int width, int height,
                                       int stride, size_t image_stride, hyorb_keypoint *kps, uint8_t *desc,
                                       int capacity, int32_t *counts);
/* Same with DEVICE pointers for images and outputs; asynchronous on the handle's stream.  Errors that are
 * only known on the device (capacity overflow) are reported by hyorb_extractor_sync(). */
HYORB_API int hyorb_extract_batch_device(hyorb_extractor *h, const uint8_t *d_images, int n_images, int width,
                                         int height, int stride, size_t image_stride, hyorb_keypoint *d_kps,
                                         uint8_t *d_desc, int capacity, int32_t *d_counts);
/* cudaStreamSynchronize + deferred device-side status of the launches since the last sync. */
HYORB_API int hyorb_extractor_sync(hyorb_extractor *h);

/* Stage outputs of the LAST extract call, for stage-by-stage parity tests (host destination buffers).
 * what: */
enum {
    HYORB_DBG_PYRAMID = 0,    /* level bytes, tightly packed w*h (ORBExtractor::mvImagePyramid without borders) */
    HYORB_DBG_BLURRED = 1,    /* GaussianBlur'ed level (ORBExtractor.cpp:536-537), w*h bytes */
    HYORB_DBG_CANDIDATES = 2, /* per-cell FAST output before distribution: int32 triples (x, y, response), lattice
                                 coordinates (pixel - 16), in the reference's order (cell row-major, row-major inside) */
    HYORB_DBG_LEVEL_COUNT = 3 /* one int32: keypoints kept on this level after DistributeOctTree */
};
HYORB_API int hyorb_extractor_level_size(hyorb_extractor *h, int width, int height, int level, int *lw, int *lh);
/* Upper bound of the keypoints one width x height image can produce with this handle's quotas: sum over levels of
 * max(quota + 3, 4 * roots) -- ORBExtractor::DistributeOctTree (ORBExtractor.cpp:107-290) stops at the first node count
 * >= quota and one split adds at most 3 nodes.  Host-batch downloads copy min(capacity, bound) entries per image while
 * the batch runs (anything beyond is fetched afterwards), so this is also the D2H size of a large batch.  Returns the
 * bound (> 0) or a negative status. */
HYORB_API int hyorb_extractor_keypoint_bound(hyorb_extractor *h, int width, int height);
/* returns the number of bytes written (>= 0) or a negative status */
HYORB_API long hyorb_extractor_debug_read(hyorb_extractor *h, int image_index, int what, int level, void *dst,
                                          size_t dst_bytes);
/* number of kernels this handle has launched since creation (bench.py's gpu_launches) */
HYORB_API long hyorb_extractor_launch_count(const hyorb_extractor *h);

/* ----- Stereo association: replaces HYSLAM::Stereomatcher (src/features/Stereomatcher.h:25-51) ----- */

/* Camera + settings actually read by Stereomatcher::computeStereoMatches (Stereomatcher.cpp:36-156). */
typedef struct hyorb_stereo_params {
    float mbf;       /* Camera::mbf (stereo baseline * fx) */
    float fx;        /* Camera::fx() */
    int32_t n_rows;  /* (int)camera.mnMaxY */
    float th_high;   /* FeatureMatcherSettings::TH_HIGH, default 100 (FeatureMatcher.h:98-103) */
    float th_low;    /* FeatureMatcherSettings::TH_LOW,  default 50 */
    float size_ref;  /* FeatureExtractorSettings::size_ref = 31 (FeatureExtractorSettings.h:27) */
} hyorb_stereo_params;

/* ImageProcessing::PreProcessImg (src/main/ImageProcessing.cpp:118-138): cv::resize(img, img, Size(), fscale, fscale)
 * followed by cvtColor(RGB|BGR[A] -> GRAY), for the scales the reference's camera configurations use: half_scale = 0
 * (fscale 1.0, copy) or 1 (fscale 0.5 = OpenCV's INTER_AREA 2x2 box path; needs 2*cvRound(src*0.5) <= src, else
 * HYORB_EUNSUPPORTED).  channels: 1, 3 or 4 interleaved bytes per pixel (alpha ignored); rgb_order: cam_data.RGB (1 = RGB[A],
 * 0 = BGR[A]).  hyorb_preprocess_size gives the gray frame's size.  hyorb_preprocess_device converts n_images device
 * frames (asynchronously on the handle's stream, like the other *_device entry points; its output can be fed to them
 * directly).  hyorb_extract_color_host = PreProcessImg + ORBExtractor::operator() on one HOST camera frame; gray_out
 * (optional, gray_out_stride bytes per row) receives the gray frame the reference keeps as track_data.image (:109). */
HYORB_API int hyorb_preprocess_size(int width, int height, int half_scale, int *out_width, int *out_height);
HYORB_API int hyorb_preprocess_device(hyorb_extractor *h, const uint8_t *d_src, int n_images, int width, int height, int stride,
                                      size_t image_stride, int channels, int rgb_order, int half_scale, uint8_t *d_gray,
                                      int gray_stride, size_t gray_image_stride);
HYORB_API int hyorb_extract_color_host(hyorb_extractor *h, const uint8_t *image, int width, int height, int stride, int channels,
                                       int rgb_order, int half_scale, uint8_t *gray_out, int gray_out_stride,
                                       hyorb_keypoint *kps, uint8_t *desc, int capacity, int *n);

/* ImageProcessing::ProcessStereoImage (src/main/ImageProcessing.cpp:69-116) over n_pairs stereo pairs in one call:
 * extractor_left(imL), extractor_right(imR) (:82-83), then Stereomatcher(...).computeStereoMatches() (:101-102).
 * Images are interleaved: image 2p = left, 2p+1 = right of pair p.  kps/desc/counts as hyorb_extract_batch_*
 * (2*n_pairs images); uR/depth are [n_pairs][capacity], indexed by left keypoint, -1 where unmatched. */
HYORB_API int hyorb_process_stereo_batch_host(hyorb_extractor *h, const hyorb_stereo_params *sp, const uint8_t *images,
                                              int n_pairs, int width, int height, int stride, size_t image_stride,
                                              hyorb_keypoint *kps, uint8_t *desc, int capacity, int32_t *counts, float *uR,
                                              float *depth);
HYORB_API int hyorb_process_stereo_batch_device(hyorb_extractor *h, const hyorb_stereo_params *sp, const uint8_t *d_images,
                                                int n_pairs, int width, int height, int stride, size_t image_stride,
                                                hyorb_keypoint *d_kps, uint8_t *d_desc, int capacity, int32_t *d_counts,
                                                float *d_uR, float *d_depth);

/* Per-stage device time of the calls since the last reset, from CUDA events recorded on the handle's stream around
 * each stage (the reference only has commented-out stopwatches, ImageProcessing.cpp:70,112-114).  Synchronises.
 * ms[HYORB_N_STAGES]: accumulated milliseconds per stage; *calls: number of profiled calls. */
enum { HYORB_STAGE_PYRAMID = 0, HYORB_STAGE_FAST = 1, HYORB_STAGE_QUADTREE = 2, HYORB_STAGE_BLUR = 3,
       HYORB_STAGE_DESCRIBE = 4, HYORB_STAGE_STEREO = 5, HYORB_N_STAGES = 6 };
HYORB_API int hyorb_extractor_set_profiling(hyorb_extractor *h, int enable);
HYORB_API int hyorb_extractor_stage_times(hyorb_extractor *h, double *ms, long *calls, int reset);

/* How a batch call is pipelined on the device (no counterpart in the reference, which is one frame per call).
 * lanes: the batch is cut into this many independent sub-batches, each on its own stream (1..16), separately for the
 * device-pointer and the host-buffer entry points; side_blur: 0 = every stage of a lane in stream order, 1 = the blur
 * on a side stream next to FAST + quadtree, 2 = next to the quadtree only (default).  A value < 0 keeps the current
 * setting.  lanes = 1, side_blur = 0 serialises the kernels, which is what makes hyorb_extractor_stage_times() report
 * each kernel's own duration. */
HYORB_API int hyorb_extractor_set_pipelining(hyorb_extractor *h, int device_lanes, int host_lanes, int side_blur);


typedef struct hyorb_matcher hyorb_matcher; /* workspace + stream for the stereo / matching entry points */
HYORB_API int hyorb_matcher_create(int device, void *cuda_stream, hyorb_matcher **out);
HYORB_API int hyorb_matcher_destroy(hyorb_matcher *m);
HYORB_API int hyorb_matcher_sync(hyorb_matcher *m);
HYORB_API long hyorb_matcher_launch_count(const hyorb_matcher *m);

/* Stereomatcher::computeStereoMatches + getData(uR, depth): uR[i], depth[i] for every left keypoint,
 * -1 where unmatched.  best_r / best_dist (optional, may be NULL): matched right index / Hamming distance. */
HYORB_API int hyorb_stereo_match_host(hyorb_matcher *m, const hyorb_stereo_params *sp,
                                      const hyorb_keypoint *kps_l, const uint8_t *desc_l, int n_l,
                                      const hyorb_keypoint *kps_r, const uint8_t *desc_r, int n_r,
                                      float *uR, float *depth, int32_t *best_r, int32_t *best_dist);
/* Batched device form over n_pairs stereo pairs laid out as hyorb_extract_batch_device leaves them:
 * image 2p = left, 2p+1 = right of pair p; kps/desc/counts strided by `capacity`.  Outputs are
 * [n_pairs][capacity]. */
HYORB_API int hyorb_stereo_match_batch_device(hyorb_matcher *m, const hyorb_stereo_params *sp, int n_pairs,
                                              const hyorb_keypoint *d_kps, const uint8_t *d_desc,
                                              const int32_t *d_counts, int capacity, float *d_uR, float *d_depth,
                                              int32_t *d_best_r, int32_t *d_best_dist);

/* ----- Descriptor matching: the inner loops of HYSLAM::FeatureMatcher + MatchCriteria ----- */

/* Acceptance rule applied to (best, second) after the scan of BestScoreCriterionCore
 * (src/features/MatchCriteria.cpp:248-280): */
enum {
    HYORB_RULE_LANDMARK = 0, /* BestScoreCriterion::apply   (:214-246): best <= thr && !(best > ratio*second) */
    HYORB_RULE_BOW = 1,      /* BestMatchBoWCriterion::apply (:601-635): best <  thr &&   best < ratio*second */
    HYORB_RULE_MONOINIT = 2  /* MonoInitBestScore::apply    (:486-523): best <= thr &&   best < second*ratio */
};

/* ORBDistance::distance (src/features/low_level/DescriptorDistance.cpp:9-25) scanned over candidate lists.
 * cand_off[nq+1] / cand_idx: CSR lists of target indices per query (a DBoW2 node's feature list,
 * FeatureMatcher.cc:281-345); both NULL = every target in index order (C4: brute force).
 * Outputs per query: best_idx (-1 if no candidate), best / second Hamming distance (65535 = none,
 * i.e. the reference's FLT_MAX), accepted (0/1). */
HYORB_API int hyorb_match_csr_host(hyorb_matcher *m, const uint8_t *q_desc, int nq, const uint8_t *t_desc, int nt,
                                   const int32_t *cand_off, const int32_t *cand_idx, int rule, float thr, float ratio,
                                   int32_t *best_idx, uint16_t *best, uint16_t *second, uint8_t *accepted);
/* Device-pointer brute force (cand lists NULL), asynchronous. */
HYORB_API int hyorb_match_bruteforce_device(hyorb_matcher *m, const uint8_t *d_q_desc, int nq, const uint8_t *d_t_desc,
                                            int nt, int rule, float thr, float ratio, int32_t *d_best_idx,
                                            uint16_t *d_best, uint16_t *d_second, uint8_t *d_accepted);

typedef struct hyorb_bounds { float min_x, max_x, min_y, max_y; } hyorb_bounds; /* Frame::mnMinX.. (Frame.cc:62-66) */

/* Frame::AssignFeaturesToGrid + PosInGrid (src/core/Frame.cc:137-153, 459-469): 64 x 48 grid, round().
 * cell_off[64*48+1], cell_idx[n]: CSR, cell id = ix*48 + iy, insertion order inside a cell. */
HYORB_API int hyorb_grid_build_host(hyorb_matcher *m, const hyorb_keypoint *kps, int n, const hyorb_bounds *b,
                                    int32_t *cell_off, int32_t *cell_idx);

/* One query per landmark of FeatureMatcher::_SearchByProjection_ (FeatureMatcher.cc:57-121). */
typedef struct hyorb_window_query {
    float u, v, r;          /* projected position and search radius (FeatureMatcher.cc:88-92) */
    float size_lo, size_hi; /* FeatureSizeCriterion bounds: keep size_lo < kp.size < size_hi (MatchCriteria.cpp:350-360) */
    float ur, ur_radius;    /* StereoConsistencyCriterion (:149-177); ur_radius < 0 = mono camera, skip */
} hyorb_window_query;

/* Frame::GetFeaturesInAreaNEW (Frame.cc:416-457) -> PreviouslyMatchedCriterion (MatchCriteria.cpp:124-144)
 * -> FeatureSizeCriterion -> StereoConsistencyCriterion -> BestScoreCriterion, for nq queries against one
 * frame's keypoints.  t_uR / t_matched may be NULL.  Builds the grid internally. */
HYORB_API int hyorb_match_window_host(hyorb_matcher *m, const hyorb_keypoint *t_kps, const uint8_t *t_desc,
                                      const float *t_uR, const uint8_t *t_matched, int nt, const hyorb_bounds *b,
                                      const hyorb_window_query *queries, const uint8_t *q_desc, int nq, float thr,
                                      float ratio, int32_t *best_idx, uint16_t *best, uint16_t *second,
                                      uint8_t *accepted);

/* Pose + camera of the frame landmarks are projected into, and one landmark, for the per-landmark front half of
 * FeatureMatcher::SearchByProjection(Frame&, landmarks, th) (FeatureMatcher.cc:123-143 -> _SearchByProjection_ :57-121). */
typedef struct hyorb_projection {
    float Rcw[9], tcw[3];   /* Frame::mRcw, mtcw, row-major */
    float Ow[3];            /* Frame::GetCameraCenter() */
    float K[9];             /* Camera::K, row-major */
    float mbf;              /* Camera::mbf */
    int32_t stereo;         /* Camera::sensor == 1: ur = u - mbf/z and the stereo consistency test apply */
    hyorb_bounds bounds;    /* Camera::mnMinX .. mnMaxY */
} hyorb_projection;
typedef struct hyorb_landmark {
    float Pw[3];                /* MapPoint::GetWorldPos() */
    float size;                 /* MapPoint::getSize(), world units */
    float min_dist, max_dist;   /* GetMinDistanceInvariance() / GetMaxDistanceInvariance() */
    int32_t assoc_idx;          /* Frame::hasAssociation(lm): keypoint of this frame already associated with it, or -1 */
} hyorb_landmark;

/* ProjectionCriterion + DistanceCriterion (MatchCriteria.cpp:13-28, 46-77), Frame::ProjectLandMark / Camera::Project
 * (Frame.cc:176-180, Camera.cpp:116-153) and Frame::landMarkSizePixels (Frame.cc:296-317) for n landmarks:
 * queries[i] = the window query _SearchByProjection_ builds (radius = th * size_px / size_ref, FeatureSizeCriterion bounds
 * frac_smaller/frac_larger * size_px, stereo radius), passed[i] = both landmark criteria hold.  t_kps is only read for
 * landmarks with assoc_idx >= 0. */
HYORB_API int hyorb_project_landmarks_host(hyorb_matcher *m, const hyorb_projection *pr, const hyorb_landmark *lms, int n,
                                           const hyorb_keypoint *t_kps, int nt, float th, float size_ref, float frac_smaller,
                                           float frac_larger, hyorb_window_query *queries, uint8_t *passed);

/* FeatureMatcher::SearchByProjection(Frame&, landmarks, th) up to (not including) the pointer-ordered associateLandMark loop:
 * projection and landmark criteria as above, then per passing landmark GetFeaturesInAreaNEW -> PreviouslyMatchedCriterion ->
 * FeatureSizeCriterion(0.5, 1.5) -> StereoConsistencyCriterion(th) -> BestScoreCriterion(thr = TH_HIGH, ratio), all on the
 * device in one call.  accepted[i] = landmark i found a keypoint (best_idx[i]); landmarks that fail a landmark criterion
 * report best_idx = -1.  passed may be NULL. */
/* Which criteria a projection search applies (the variants of FeatureMatcher::SearchByProjection differ only in these and in
 * thr / ratio): local map (FeatureMatcher.cc:123-143) = DISTANCE | STEREO; motion model (:145-176) = STEREO | ROTATION;
 * relocalisation against a keyframe (:180-213) = DISTANCE | ROTATION with ratio 1.0. */
enum { HYORB_SBP_DISTANCE = 1,   /* DistanceCriterion (MatchCriteria.cpp:46-77) */
       HYORB_SBP_STEREO = 2,     /* StereoConsistencyCriterion(th) (:149-177); only has an effect for a stereo camera */
       HYORB_SBP_ROTATION = 4,   /* RotationConsistencyCriterion (:363-401) against lm_prev_angle */
       HYORB_SBP_VIEWANGLE = 8 };/* ViewingAngleCriterion (:84-110); hyorb_fuse_host only (needs the landmark normals) */

/* As hyorb_search_by_projection_host with the criteria chosen by `flags`.  HYORB_SBP_ROTATION needs lm_prev_angle[i] = angle of
 * the keypoint landmark i is associated with in the previous frame, and the landmarks listed in the reference's map order
 * (ascending MapPoint*): where several landmarks match the same keypoint the reference keeps the last one in that order. */
HYORB_API int hyorb_search_by_projection_ex_host(hyorb_matcher *m, const hyorb_projection *pr, const hyorb_landmark *lms,
                                                 const uint8_t *lm_desc, const float *lm_prev_angle, int n,
                                                 const hyorb_keypoint *t_kps, const uint8_t *t_desc, const float *t_uR,
                                                 const uint8_t *t_matched, int nt, float th, float size_ref, float thr,
                                                 float ratio, unsigned flags, int32_t *best_idx, uint16_t *best,
                                                 uint16_t *second, uint8_t *accepted, uint8_t *passed);

HYORB_API int hyorb_search_by_projection_host(hyorb_matcher *m, const hyorb_projection *pr, const hyorb_landmark *lms,
                                              const uint8_t *lm_desc, int n, const hyorb_keypoint *t_kps, const uint8_t *t_desc,
                                              const float *t_uR, const uint8_t *t_matched, int nt, float th, float size_ref,
                                              float thr, float ratio, int32_t *best_idx, uint16_t *best, uint16_t *second,
                                              uint8_t *accepted, uint8_t *passed);

/* FeatureMatcher::Fuse(KeyFrame*, landmarks, fuse_matches, th, reprojection_err) (FeatureMatcher.cc:464-521) up to the insertion into
 * fuse_matches.  The caller pre-screens the landmarks the way :480-487 does (not null / bad / already in the keyframe / protected) and
 * lists the survivors in vector order.  On the device: ProjectionCriterion + DistanceCriterion + ViewingAngleCriterion(max_angle)
 * (MatchCriteria.cpp:13-110; lm_normal = MapPoint::GetNormal(), cos_max_angle = cosf(max_angle) evaluated by the caller's libm, as
 * the reference's `cos(max_angle)` is), KeyFrame::ProjectLandMark / landMarkSizePixels, GetFeaturesInArea, FeatureSizeCriterion(0.5, 1.5),
 * ProjectionViewCriterion(reproj_err) (:282-333 over KeyFrame::ReprojectionError, KeyFrame.cc:548-575; sigma_ref / size_ref =
 * FeatureExtractorSettings of the keyframe's views; t_uR = NULL for monocular views) and BestScoreCriterion(thr = TH_LOW, ratio = 1.0).
 * accepted[i] != 0 <=> the reference would call fuse_matches.insert({best_idx[i], landmark i}); std::map::insert keeps the FIRST
 * landmark that reaches a keypoint. */
HYORB_API int hyorb_fuse_host(hyorb_matcher *m, const hyorb_projection *pr, const hyorb_landmark *lms, const float *lm_normal,
                              const uint8_t *lm_desc, int n, const hyorb_keypoint *t_kps, const uint8_t *t_desc, const float *t_uR, int nt,
                              float th, float size_ref, float sigma_ref, float reproj_err, float cos_max_angle, float thr, float ratio,
                              int32_t *best_idx, uint16_t *best, uint16_t *second, uint8_t *accepted, uint8_t *passed);

/* One direction of FeatureMatcher::SearchBySim3 (FeatureMatcher.cc:739-937; :783-845 is KF1 -> KF2, :848-910 the mirror image): the
 * landmarks of keyframe A (lms / lm_desc, those that pass :787-793) are carried into camera A with its pose (R_a, t_a), into camera B
 * with the similarity (sR_ba, t_ba) -- the caller forms sR21 = (1/s12) * R12.t(), t21 = -sR21 * t12 / sR12 = s12 * R12, t12 exactly as
 * :753-756 does -- projected with B's camera, range-checked against the landmark's distance invariance, and matched against B's
 * features inside th * B.landMarkSizePixels(lm) / size_ref (pr_b = B's own pose and camera; assoc_idx = B.hasAssociation(lm)) with no
 * view criteria: accepted[i] <=> best distance <= thr (TH_HIGH), best_idx[i] = vnMatch[i].  The caller runs both directions and keeps
 * the mutually agreeing pairs (:913-930). */
HYORB_API int hyorb_search_by_sim3_host(hyorb_matcher *m, const float *R_a, const float *t_a, const float *sR_ba, const float *t_ba,
                                        const hyorb_projection *pr_b, const hyorb_landmark *lms, const uint8_t *lm_desc, int n,
                                        const hyorb_keypoint *kps_b, const uint8_t *desc_b, int nb, float th, float size_ref, float thr,
                                        int32_t *best_idx, uint16_t *best, uint8_t *accepted, uint8_t *passed);

/* FeatureMatcher::SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize) (FeatureMatcher.cc:404-462) with
 * MonoInitScoreExceedsPrevious and MonoInitBestScore(thr = TH_LOW, ratio = nnratio) (MatchCriteria.cpp:486-549) and the rotation
 * histogram.  prev_xy (n1 x 2) is vbPrevMatched, updated in place for matched features; matches12[i1] = index in frame 2 or -1;
 * *n_matches = the reference's return value.  The reference's loop is sequential (a later feature may take a frame-2 feature from an
 * earlier one only with a smaller distance); the device solves the same triangular system by fixed-point passes (match.cu). */
HYORB_API int hyorb_search_for_initialization_host(hyorb_matcher *m, const hyorb_keypoint *k1, const uint8_t *d1, int n1,
                                                   const hyorb_keypoint *k2, const uint8_t *d2, int n2, hyorb_bounds bounds2,
                                                   float *prev_xy, int window, float thr, float ratio, int32_t *matches12,
                                                   int32_t *n_matches);

/* RotationConsistency + ComputeThreeMaxima (MatchCriteria.cpp:684-767): keep[i] = 1 if match i (angles of the
 * two matched keypoints, pairs in ascending current-index order) falls in one of the 3 dominant rotation bins. */
HYORB_API int hyorb_rotation_consistency_host(hyorb_matcher *m, const float *angle_prev, const float *angle_curr,
                                              int n, uint8_t *keep);

/* Bag-of-words vocabulary tree, resident on the device.  hySLAM reaches it through DBoW2 (ORBVocabulary::transform,
 * src/features/low_level/ORBVocabulary.cpp:31-42; Frame::ComputeBoW, src/core/Frame.cc:472-479, levelsup = 4); DBoW2 and its
 * vocabulary file are not part of the reference tree, so the tree is handed over explicitly: node 0 = root, children of node i =
 * child_idx[child_off[i] .. child_off[i+1]) in DBoW2's child order, node_desc = 32 bytes per node (the root's are ignored),
 * word_of[i] = word id of a leaf (-1 for inner nodes), weight_of[i] = its weight, L = depth of the leaves. */
typedef struct hyorb_vocabulary hyorb_vocabulary;
HYORB_API int hyorb_vocabulary_create(int device, int n_nodes, int L, const int32_t *child_off, const int32_t *child_idx,
                                      const uint8_t *node_desc, const int32_t *word_of, const float *weight_of, hyorb_vocabulary **out);
HYORB_API int hyorb_vocabulary_destroy(hyorb_vocabulary *v);

/* DBoW2::TemplatedVocabulary::transform per feature: descend the tree taking the child with the smallest Hamming distance at
 * every level (first child wins ties), word_id / weight = the leaf reached, node_id = the node passed at level L - levelsup
 * (the key of DBoW2::FeatureVector; 0 = root when L - levelsup <= 0). */
HYORB_API int hyorb_bow_transform_host(hyorb_matcher *m, const hyorb_vocabulary *v, const uint8_t *desc, int n, int levelsup,
                                       int32_t *word_id, int32_t *node_id, float *weight);

/* FeatureMatcher::_SearchByBoW_ / SearchForTriangulation (FeatureMatcher.cc:281-345, 373-402) with the candidate gating on the
 * device: both descriptor sets are quantised, the features of set 2 are grouped by node (index order inside a node, as
 * FeatureVector::addFeature appends them), and every feature of set 1 scans the set-2 features under its own node with the
 * acceptance rule `rule`.  mask1 / mask2 (may be NULL): the BoWIndexCriterion filters, 0 = feature not eligible.
 * node1 / node2 (may be NULL) return the node ids. */
HYORB_API int hyorb_search_by_bow_host(hyorb_matcher *m, const hyorb_vocabulary *v, const uint8_t *desc1, const uint8_t *mask1, int n1,
                                       const uint8_t *desc2, const uint8_t *mask2, int n2, int levelsup, int rule, float thr,
                                       float ratio, int32_t *node1, int32_t *node2, int32_t *best_idx, uint16_t *best,
                                       uint16_t *second, uint8_t *accepted);

/* FeatureMatcher::SearchForTriangulation (FeatureMatcher.cc:373-402) up to the rotation histogram: _SearchByBoW_ with
 * EpipolarConsistencyBoWCriterion(F12) (MatchCriteria.cpp:641-676: the candidate must lie within 3.84 * sigma2(kp2.size) of the
 * epipolar line x1' F12, sigma2 = FeatureExtractorSettings::determineSigma2 = sigma_ref * (size / size_ref)^2) in front of
 * BestMatchBoWCriterion(thr = TH_LOW, ratio = 1.0).  F12: 9 floats, row-major.  mask1 / mask2 carry the index criteria
 * (PreviouslyMatchedIndexCriterion(false), StereoIndexCriterion for bOnlyStereo).  RotationConsistencyBoW follows through
 * hyorb_rotation_consistency_host. */
HYORB_API int hyorb_search_for_triangulation_host(hyorb_matcher *m, const hyorb_vocabulary *v, const hyorb_keypoint *kps1,
                                                  const uint8_t *desc1, const uint8_t *mask1, int n1, const hyorb_keypoint *kps2,
                                                  const uint8_t *desc2, const uint8_t *mask2, int n2, int levelsup, const float *F12,
                                                  float sigma_ref, float size_ref, float thr, float ratio, int32_t *node1,
                                                  int32_t *node2, int32_t *best_idx, uint16_t *best, uint16_t *second,
                                                  uint8_t *accepted);

/* The same epipolar gate in front of an explicit candidate-list scan (hyorb_match_csr_host): for a hySLAM build that already
 * holds DBoW2 FeatureVectors, the node lists go in as CSR and nothing is re-quantised. */
HYORB_API int hyorb_match_csr_epipolar_host(hyorb_matcher *m, const hyorb_keypoint *q_kps, const uint8_t *q_desc, int nq,
                                            const hyorb_keypoint *t_kps, const uint8_t *t_desc, int nt, const int32_t *cand_off,
                                            const int32_t *cand_idx, const float *F12, float sigma_ref, float size_ref, int rule,
                                            float thr, float ratio, int32_t *best_idx, uint16_t *best, uint16_t *second,
                                            uint8_t *accepted);

/* Representative descriptor of a landmark: MapPointDBEntry::_computeDistinctiveDescriptor_ (src/core/MapPointDB.cpp:127-171).
 * desc: the observation descriptors of all landmarks back to back (32 bytes each); lm_off[n_landmarks + 1]: CSR offsets
 * (rows of landmark l = lm_off[l] .. lm_off[l+1]).  Per landmark: all-pairs Hamming distances, per-row order statistic
 * sorted[(int)(0.5*(N-1))], first row with the smallest one.  best_idx[l] is relative to the landmark's first row (-1
 * for an empty list: the reference returns without touching the landmark), best_median[l] that row's median. */
HYORB_API int hyorb_distinctive_descriptor_host(hyorb_matcher *m, const uint8_t *desc, const int32_t *lm_off, int n_landmarks,
                                                int32_t *best_idx, int32_t *best_median);

/* ----- misc ----- */
HYORB_API const char *hyorb_last_error(void);
HYORB_API const char *hyorb_version(void);
HYORB_API int hyorb_device_count(void); /* 0 when no CUDA device / driver */

#ifdef __cplusplus
}
#endif
#endif /* HYORB_H */
