"""Host-side mirrors of HYSLAM::Stereomatcher (src/features/Stereomatcher.h:25-51) and of the Hamming scans of
HYSLAM::FeatureMatcher / MatchCriteria (src/features/FeatureMatcher.h:105-176) over the C ABI."""
import ctypes as C

import numpy as np

from . import _ffi as F
from .settings import FeatureExtractorSettings, FeatureMatcherSettings


class _Handle:
    def __init__(self, device=0, stream=None):
        self._h = C.c_void_p()
        F.check(F.lib().hyorb_matcher_create(int(device), stream, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            F.lib().hyorb_matcher_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:       # interpreter teardown: module globals may already be gone
            pass

    def sync(self):
        F.check(F.lib().hyorb_matcher_sync(self._h))

    def launch_count(self):
        return int(F.lib().hyorb_matcher_launch_count(self._h))


class Stereomatcher(_Handle):
    """``Stereomatcher(views, camera, settings)`` -> ``computeStereoMatches()`` -> ``getData()`` as in
    ImageProcessing::ProcessStereoImage (src/main/ImageProcessing.cpp:100-103).  ``views`` is the tuple
    (kps_left, desc_left, kps_right, desc_right) the reference packs into a FeatureViews."""

    def __init__(self, views, camera, settings=None, extractor_settings=None, device=0, stream=None):
        super().__init__(device, stream)
        self.kl, self.dl, self.kr, self.dr = views
        ms = settings or FeatureMatcherSettings()
        es = extractor_settings or FeatureExtractorSettings()
        self.params = F.StereoParams(float(camera.mbf), float(camera.fx), int(camera.mnMaxY), float(ms.TH_HIGH), float(ms.TH_LOW),
                                     float(es.size_ref))
        self.mvuRight = self.mvDepth = self.best_r = self.best_dist = None

    def computeStereoMatches(self):
        kl, kr = np.ascontiguousarray(self.kl, F.KP_DTYPE), np.ascontiguousarray(self.kr, F.KP_DTYPE)
        dl, dr = np.ascontiguousarray(self.dl, np.uint8), np.ascontiguousarray(self.dr, np.uint8)
        nl = len(kl)
        self.mvuRight = np.full(nl, -1, np.float32)
        self.mvDepth = np.full(nl, -1, np.float32)
        self.best_r = np.full(nl, -1, np.int32)
        self.best_dist = np.full(nl, -1, np.int32)
        F.check(F.lib().hyorb_stereo_match_host(self._h, C.byref(self.params), F.ptr(kl), F.ptr(dl), nl, F.ptr(kr), F.ptr(dr), len(kr),
                                                F.ptr(self.mvuRight), F.ptr(self.mvDepth), F.ptr(self.best_r), F.ptr(self.best_dist)))

    def getData(self):
        return self.mvuRight, self.mvDepth

    def match_batch_device(self, n_pairs, d_kps, d_desc, d_counts, capacity, d_uR, d_depth, d_best_r=None, d_best_dist=None):
        F.check(F.lib().hyorb_stereo_match_batch_device(self._h, C.byref(self.params), n_pairs, d_kps, d_desc, d_counts, capacity,
                                                        d_uR, d_depth, d_best_r, d_best_dist))


class Vocabulary:
    """Device-resident bag-of-words tree (the role of HYSLAM::ORBVocabulary / DBoW2::TemplatedVocabulary,
    src/features/low_level/ORBVocabulary.cpp).  ``tree``: dict(L, child_off, child_idx, node_desc, word_of, weight_of); node 0 = root."""

    def __init__(self, tree, device=0):
        self.tree = tree
        self.L = int(tree["L"])
        co = np.ascontiguousarray(tree["child_off"], np.int32); ci = np.ascontiguousarray(tree["child_idx"], np.int32)
        nd = np.ascontiguousarray(tree["node_desc"], np.uint8); wo = np.ascontiguousarray(tree["word_of"], np.int32)
        wt = np.ascontiguousarray(tree["weight_of"], np.float32)
        self._h = C.c_void_p()
        F.check(F.lib().hyorb_vocabulary_create(int(device), len(co) - 1, self.L, F.ptr(co), F.ptr(ci), F.ptr(nd), F.ptr(wo), F.ptr(wt), C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            F.lib().hyorb_vocabulary_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:       # interpreter teardown: module globals may already be gone
            pass

    @staticmethod
    def random_tree(k, L, seed, shrink=0.0):
        """Stand-in vocabulary (the reference's ORBvoc file is not part of its tree): complete k-ary tree of depth L in breadth-first
        order with seeded random node descriptors; ``shrink`` removes that fraction of the children lists' tails to exercise
        ragged trees.  Leaves get consecutive word ids and positive weights."""
        rng = np.random.default_rng(seed)
        cnt_list, frontier, n_nodes = [], 1, 1          # breadth-first: cnt_list[i] = number of children of node i
        for d in range(L):
            nxt = 0
            for _ in range(frontier):
                kk = k if rng.random() >= shrink else max(1, int(rng.integers(1, k + 1)))
                cnt_list.append(kk); nxt += kk
            n_nodes += nxt
            frontier = nxt
        cnt = np.zeros(n_nodes, np.int32)
        cnt[: len(cnt_list)] = cnt_list
        off = np.zeros(n_nodes + 1, np.int32)
        off[1:] = np.cumsum(cnt)
        idx = np.arange(1, n_nodes, dtype=np.int32)        # breadth-first numbering: the children of consecutive parents are consecutive
        node_desc = rng.integers(0, 256, (n_nodes, 32), dtype=np.uint8)
        leaf = cnt == 0
        word_of = np.full(n_nodes, -1, np.int32); word_of[leaf] = np.arange(int(leaf.sum()), dtype=np.int32)
        weight_of = np.zeros(n_nodes, np.float32); weight_of[leaf] = rng.uniform(0.1, 5.0, int(leaf.sum())).astype(np.float32)
        return dict(L=L, child_off=off, child_idx=idx, node_desc=node_desc, word_of=word_of, weight_of=weight_of)


class FeatureMatcher(_Handle):
    """The data-parallel inner loops of HYSLAM::FeatureMatcher: candidate enumeration (grid window / explicit CSR
    lists), Hamming scan, best / second-best, acceptance rule, rotation histogram.  Landmark projection and map
    bookkeeping stay with the caller."""

    def __init__(self, settings=None, device=0, stream=None):
        super().__init__(device, stream)
        self.settings = settings or FeatureMatcherSettings()

    def _outs(self, nq):
        return (np.full(nq, -1, np.int32), np.full(nq, 65535, np.uint16), np.full(nq, 65535, np.uint16), np.zeros(nq, np.uint8))

    def match(self, q_desc, t_desc, cand_off=None, cand_idx=None, rule=F.RULE_BOW, thr=None, ratio=None):
        """Best / second-best scan of every query over its candidate list (BoW-node lists of
        FeatureMatcher::_SearchByBoW_, FeatureMatcher.cc:281-345) or, with no lists, over all targets in index order."""
        q = np.ascontiguousarray(q_desc, np.uint8); t = np.ascontiguousarray(t_desc, np.uint8)
        nq, nt = len(q), len(t)
        thr = float(self.settings.TH_LOW if thr is None else thr)
        ratio = float(self.settings.nnratio if ratio is None else ratio)
        if cand_off is not None:
            cand_off = np.ascontiguousarray(cand_off, np.int32); cand_idx = np.ascontiguousarray(cand_idx, np.int32)
            if len(cand_idx) == 0:
                cand_idx = np.zeros(1, np.int32)
        bi, b, s, acc = self._outs(nq)
        F.check(F.lib().hyorb_match_csr_host(self._h, F.ptr(q), nq, F.ptr(t), nt, F.ptr(cand_off), F.ptr(cand_idx), int(rule), thr, ratio,
                                             F.ptr(bi), F.ptr(b), F.ptr(s), F.ptr(acc)))
        return bi, b, s, acc

    def SearchForTriangulation(self, desc1, desc2, cand_off=None, cand_idx=None, ratio=None):
        """FeatureMatcher::SearchForTriangulation's scan (FeatureMatcher.cc:373-402): BestMatchBoWCriterion(TH_LOW, ratio)."""
        return self.match(desc1, desc2, cand_off, cand_idx, F.RULE_BOW, self.settings.TH_LOW, ratio)

    def grid_build(self, kps, bounds):
        kps = np.ascontiguousarray(kps, F.KP_DTYPE)
        off = np.zeros(F.GRID_COLS * F.GRID_ROWS + 1, np.int32); idx = np.zeros(max(len(kps), 1), np.int32)
        b = F.Bounds(*[float(v) for v in bounds])
        F.check(F.lib().hyorb_grid_build_host(self._h, F.ptr(kps), len(kps), C.byref(b), F.ptr(off), F.ptr(idx)))
        return off, idx[: off[-1]].copy()

    def SearchByProjection(self, t_kps, t_desc, bounds, queries, q_desc, t_uR=None, t_matched=None, thr=None, ratio=None):
        """Per-landmark window search of FeatureMatcher::_SearchByProjection_ (FeatureMatcher.cc:57-121)."""
        t_kps = np.ascontiguousarray(t_kps, F.KP_DTYPE); t_desc = np.ascontiguousarray(t_desc, np.uint8)
        queries = np.ascontiguousarray(queries, F.WQ_DTYPE); q_desc = np.ascontiguousarray(q_desc, np.uint8)
        t_uR = None if t_uR is None else np.ascontiguousarray(t_uR, np.float32)
        t_matched = None if t_matched is None else np.ascontiguousarray(t_matched, np.uint8)
        thr = float(self.settings.TH_HIGH if thr is None else thr)
        ratio = float(self.settings.nnratio if ratio is None else ratio)
        nq = len(queries)
        b = F.Bounds(*[float(v) for v in bounds])
        bi, bb, s, acc = self._outs(nq)
        F.check(F.lib().hyorb_match_window_host(self._h, F.ptr(t_kps), F.ptr(t_desc), F.ptr(t_uR), F.ptr(t_matched), len(t_kps), C.byref(b),
                                                F.ptr(queries), F.ptr(q_desc), nq, thr, ratio, F.ptr(bi), F.ptr(bb), F.ptr(s), F.ptr(acc)))
        return bi, bb, s, acc

    def RotationConsistency(self, angle_prev, angle_curr):
        a = np.ascontiguousarray(angle_prev, np.float32); c = np.ascontiguousarray(angle_curr, np.float32)
        keep = np.zeros(len(a), np.uint8)
        F.check(F.lib().hyorb_rotation_consistency_host(self._h, F.ptr(a), F.ptr(c), len(a), F.ptr(keep)))
        return keep

    @staticmethod
    def make_projection(Rcw, tcw, Ow, K, mbf, stereo, bounds):
        """hyorb_projection from the frame's pose (Frame::mRcw, mtcw, GetCameraCenter()) and camera (K, mbf, sensor, image bounds)"""
        pr = F.Projection()
        pr.Rcw[:] = [float(v) for v in np.asarray(Rcw, np.float32).reshape(9)]
        pr.tcw[:] = [float(v) for v in np.asarray(tcw, np.float32).reshape(3)]
        pr.Ow[:] = [float(v) for v in np.asarray(Ow, np.float32).reshape(3)]
        pr.K[:] = [float(v) for v in np.asarray(K, np.float32).reshape(9)]
        pr.mbf = float(mbf); pr.stereo = int(stereo)
        pr.bounds.min_x, pr.bounds.max_x, pr.bounds.min_y, pr.bounds.max_y = [float(v) for v in bounds]
        return pr

    def ProjectLandMarks(self, pr, landmarks, t_kps, th, size_ref=31.0, frac_smaller=0.5, frac_larger=1.5):
        """ProjectionCriterion + DistanceCriterion + ProjectLandMark + landMarkSizePixels for every landmark (MatchCriteria.cpp:13-77,
        Frame.cc:176-180, 296-317): (window queries [WQ_DTYPE], passed flags)."""
        lms = np.ascontiguousarray(landmarks, F.LM_DTYPE); t_kps = np.ascontiguousarray(t_kps, F.KP_DTYPE)
        n = len(lms)
        q = np.zeros(n, F.WQ_DTYPE); passed = np.zeros(n, np.uint8)
        F.check(F.lib().hyorb_project_landmarks_host(self._h, C.byref(pr), F.ptr(lms), n, F.ptr(t_kps), len(t_kps), float(th), float(size_ref),
                                                     float(frac_smaller), float(frac_larger), F.ptr(q), F.ptr(passed)))
        return q, passed

    def SearchByProjectionLandMarks(self, pr, landmarks, lm_desc, t_kps, t_desc, th, t_uR=None, t_matched=None, size_ref=31.0, thr=None, ratio=None,
                                    flags=F.SBP_DISTANCE | F.SBP_STEREO, lm_prev_angle=None):
        """FeatureMatcher::SearchByProjection up to the association loop, in one device call.  ``flags`` picks the variant: local map
        (FeatureMatcher.cc:123-143) = SBP_DISTANCE | SBP_STEREO (default); motion model (:145-176) = SBP_STEREO | SBP_ROTATION with
        ``lm_prev_angle``; relocalisation (:180-213) = SBP_DISTANCE | SBP_ROTATION, ratio 1.0.
        Returns (best_idx, best, second, accepted, passed) per landmark."""
        lms = np.ascontiguousarray(landmarks, F.LM_DTYPE); lm_desc = np.ascontiguousarray(lm_desc, np.uint8)
        t_kps = np.ascontiguousarray(t_kps, F.KP_DTYPE); t_desc = np.ascontiguousarray(t_desc, np.uint8)
        t_uR = None if t_uR is None else np.ascontiguousarray(t_uR, np.float32)
        t_matched = None if t_matched is None else np.ascontiguousarray(t_matched, np.uint8)
        thr = float(self.settings.TH_HIGH if thr is None else thr)
        ratio = float(self.settings.nnratio if ratio is None else ratio)
        n = len(lms)
        bi, b, s, acc = self._outs(n)
        passed = np.zeros(n, np.uint8)
        pa = None if lm_prev_angle is None else np.ascontiguousarray(lm_prev_angle, np.float32)
        F.check(F.lib().hyorb_search_by_projection_ex_host(self._h, C.byref(pr), F.ptr(lms), F.ptr(lm_desc), F.ptr(pa), n, F.ptr(t_kps), F.ptr(t_desc),
                                                           F.ptr(t_uR), F.ptr(t_matched), len(t_kps), float(th), float(size_ref), thr, ratio,
                                                           int(flags), F.ptr(bi), F.ptr(b), F.ptr(s), F.ptr(acc), F.ptr(passed)))
        return bi, b, s, acc, passed

    def Fuse(self, pr, landmarks, lm_normal, lm_desc, t_kps, t_desc, th=3.0, reprojection_err=5.99, t_uR=None, size_ref=31.0, sigma_ref=1.0,
             max_angle=1.047, thr=None, ratio=1.0):
        """FeatureMatcher::Fuse(pKF, landmarks, fuse_matches, th, reprojection_err) (FeatureMatcher.cc:464-521) up to the insertion into
        fuse_matches, for landmarks the caller has pre-screened like :480-487.  Returns (best_idx, best, second, accepted, passed);
        ``fuse_matches`` = first accepted landmark per keypoint (std::map::insert)."""
        lms = np.ascontiguousarray(landmarks, F.LM_DTYPE); lm_desc = np.ascontiguousarray(lm_desc, np.uint8)
        nrm = np.ascontiguousarray(lm_normal, np.float32).reshape(-1, 3)
        t_kps = np.ascontiguousarray(t_kps, F.KP_DTYPE); t_desc = np.ascontiguousarray(t_desc, np.uint8)
        t_uR = None if t_uR is None else np.ascontiguousarray(t_uR, np.float32)
        thr = float(self.settings.TH_LOW if thr is None else thr)
        n = len(lms)
        bi, b, s, acc = self._outs(n)
        passed = np.zeros(n, np.uint8)
        cos_max = float(np.cos(np.float32(max_angle), dtype=np.float32))      # the reference evaluates cos(float) = cosf on the host
        F.check(F.lib().hyorb_fuse_host(self._h, C.byref(pr), F.ptr(lms), F.ptr(nrm), F.ptr(lm_desc), n, F.ptr(t_kps), F.ptr(t_desc), F.ptr(t_uR), len(t_kps),
                                        float(th), float(size_ref), float(sigma_ref), float(reprojection_err), cos_max, thr, float(ratio),
                                        F.ptr(bi), F.ptr(b), F.ptr(s), F.ptr(acc), F.ptr(passed)))
        return bi, b, s, acc, passed

    def SearchBySim3Direction(self, R_a, t_a, sR_ba, t_ba, pr_b, landmarks, lm_desc, kps_b, desc_b, th, size_ref=31.0, thr=None):
        """One direction of FeatureMatcher::SearchBySim3 (FeatureMatcher.cc:783-845 / 848-910).  Returns (best_idx, best, accepted, passed)."""
        f = lambda a: np.ascontiguousarray(a, np.float32).reshape(-1)
        Ra, ta, sR, tb = f(R_a), f(t_a), f(sR_ba), f(t_ba)
        lms = np.ascontiguousarray(landmarks, F.LM_DTYPE); lm_desc = np.ascontiguousarray(lm_desc, np.uint8)
        kb = np.ascontiguousarray(kps_b, F.KP_DTYPE); db = np.ascontiguousarray(desc_b, np.uint8)
        thr = float(self.settings.TH_HIGH if thr is None else thr)
        n = len(lms)
        bi = np.full(n, -1, np.int32); b = np.full(n, 65535, np.uint16); acc = np.zeros(n, np.uint8); passed = np.zeros(n, np.uint8)
        F.check(F.lib().hyorb_search_by_sim3_host(self._h, F.ptr(Ra), F.ptr(ta), F.ptr(sR), F.ptr(tb), C.byref(pr_b), F.ptr(lms), F.ptr(lm_desc), n, F.ptr(kb),
                                                  F.ptr(db), len(kb), float(th), float(size_ref), thr, F.ptr(bi), F.ptr(b), F.ptr(acc), F.ptr(passed)))
        return bi, b, acc, passed

    def SearchForInitialization(self, k1, d1, k2, d2, bounds, prev_matched, window=100, thr=None, ratio=None):
        """FeatureMatcher::SearchForInitialization (FeatureMatcher.cc:404-462).  Returns (n_matches, vnMatches12, updated vbPrevMatched)."""
        k1 = np.ascontiguousarray(k1, F.KP_DTYPE); k2 = np.ascontiguousarray(k2, F.KP_DTYPE)
        d1 = np.ascontiguousarray(d1, np.uint8); d2 = np.ascontiguousarray(d2, np.uint8)
        pm = np.ascontiguousarray(prev_matched, np.float32).copy()
        thr = float(self.settings.TH_LOW if thr is None else thr)
        ratio = float(self.settings.nnratio if ratio is None else ratio)
        m12 = np.full(len(k1), -1, np.int32)
        nm = C.c_int32(0)
        b = F.Bounds(*[float(v) for v in bounds])
        F.check(F.lib().hyorb_search_for_initialization_host(self._h, F.ptr(k1), F.ptr(d1), len(k1), F.ptr(k2), F.ptr(d2), len(k2), b, F.ptr(pm), int(window),
                                                             thr, ratio, F.ptr(m12), C.addressof(nm)))
        return nm.value, m12, pm

    def BowTransform(self, vocab, desc, levelsup=4):
        """ORBVocabulary::transform (ORBVocabulary.cpp:31-42) per feature: (word_id, node_id at level L - levelsup, weight).
        ``feature_vector(node_id)`` / ``bow_vector(word_id, weight)`` assemble DBoW2's two containers from them."""
        desc = np.ascontiguousarray(desc, np.uint8)
        n = len(desc)
        w = np.full(n, -1, np.int32); nid = np.full(n, -1, np.int32); wt = np.zeros(n, np.float32)
        F.check(F.lib().hyorb_bow_transform_host(self._h, vocab._h, F.ptr(desc), n, int(levelsup), F.ptr(w), F.ptr(nid), F.ptr(wt)))
        return w, nid, wt

    @staticmethod
    def feature_vector(node_id):
        """DBoW2::FeatureVector as CSR: (sorted unique node ids, offsets, feature indices in index order inside a node)"""
        order = np.argsort(node_id, kind="stable").astype(np.int32)
        nodes, start = np.unique(np.asarray(node_id)[order], return_index=True)
        return nodes, np.append(start, len(order)).astype(np.int32), order

    @staticmethod
    def bow_vector(word_id, weight):
        """DBoW2::BowVector for TF_IDF weighting + L1 normalisation (the ORB vocabulary's settings): {word: weight}"""
        acc = {}
        for w, x in zip(word_id.tolist(), weight.tolist()):
            if x > 0:
                acc[w] = acc.get(w, 0.0) + x
        if acc:
            nd = float(len(acc))
            acc = {w: x / nd for w, x in acc.items()}
            norm = sum(abs(x) for x in acc.values())
            if norm > 0:
                acc = {w: x / norm for w, x in acc.items()}
        return acc

    def SearchByBoW(self, vocab, desc1, desc2, mask1=None, mask2=None, levelsup=4, rule=F.RULE_BOW, thr=None, ratio=None):
        """FeatureMatcher::_SearchByBoW_ / SearchForTriangulation (FeatureMatcher.cc:281-345, 373-402) with the DBoW2 gating on the
        device.  Returns (best_idx, best, second, accepted, node1, node2)."""
        d1 = np.ascontiguousarray(desc1, np.uint8); d2 = np.ascontiguousarray(desc2, np.uint8)
        m1 = None if mask1 is None else np.ascontiguousarray(mask1, np.uint8)
        m2 = None if mask2 is None else np.ascontiguousarray(mask2, np.uint8)
        thr = float(self.settings.TH_LOW if thr is None else thr)
        ratio = float(self.settings.nnratio if ratio is None else ratio)
        n1, n2 = len(d1), len(d2)
        bi, b, s, acc = self._outs(n1)
        node1 = np.full(n1, -1, np.int32); node2 = np.full(max(n2, 1), -1, np.int32)
        F.check(F.lib().hyorb_search_by_bow_host(self._h, vocab._h, F.ptr(d1), F.ptr(m1), n1, F.ptr(d2), F.ptr(m2), n2, int(levelsup), int(rule),
                                                 thr, ratio, F.ptr(node1), F.ptr(node2), F.ptr(bi), F.ptr(b), F.ptr(s), F.ptr(acc)))
        return bi, b, s, acc, node1, node2[:n2]

    def match_csr_epipolar(self, kps1, desc1, kps2, desc2, cand_off, cand_idx, F12, rule=F.RULE_BOW, thr=None, ratio=1.0, sigma_ref=1.0, size_ref=31.0):
        """Candidate-list scan behind EpipolarConsistencyBoWCriterion (MatchCriteria.cpp:641-676): the form SearchForTriangulation takes
        when the DBoW2 FeatureVectors already exist on the host.  Returns (best_idx, best, second, accepted)."""
        k1 = np.ascontiguousarray(kps1, F.KP_DTYPE); k2 = np.ascontiguousarray(kps2, F.KP_DTYPE)
        d1 = np.ascontiguousarray(desc1, np.uint8); d2 = np.ascontiguousarray(desc2, np.uint8)
        off = np.ascontiguousarray(cand_off, np.int32); idx = np.ascontiguousarray(cand_idx, np.int32)
        Fm = np.ascontiguousarray(F12, np.float32).reshape(9)
        thr = float(self.settings.TH_LOW if thr is None else thr)
        bi, b, s, acc = self._outs(len(d1))
        F.check(F.lib().hyorb_match_csr_epipolar_host(self._h, F.ptr(k1), F.ptr(d1), len(d1), F.ptr(k2), F.ptr(d2), len(d2), F.ptr(off), F.ptr(idx),
                                                      F.ptr(Fm), float(sigma_ref), float(size_ref), int(rule), thr, float(ratio), F.ptr(bi), F.ptr(b),
                                                      F.ptr(s), F.ptr(acc)))
        return bi, b, s, acc

    def SearchForTriangulation(self, vocab, kps1, desc1, kps2, desc2, F12, mask1=None, mask2=None, levelsup=4, sigma_ref=1.0, size_ref=31.0,
                               thr=None, ratio=1.0):
        """FeatureMatcher::SearchForTriangulation (FeatureMatcher.cc:373-402) up to the rotation histogram: BoW-gated scan with
        EpipolarConsistencyBoWCriterion(F12) (MatchCriteria.cpp:641-676) in front of BestMatchBoWCriterion(TH_LOW, 1.0).  mask1 / mask2 =
        the index criteria (unmatched features; stereo features when bOnlyStereo).  Returns (best_idx, best, second, accepted, node1, node2)."""
        k1 = np.ascontiguousarray(kps1, F.KP_DTYPE); k2 = np.ascontiguousarray(kps2, F.KP_DTYPE)
        d1 = np.ascontiguousarray(desc1, np.uint8); d2 = np.ascontiguousarray(desc2, np.uint8)
        m1 = None if mask1 is None else np.ascontiguousarray(mask1, np.uint8)
        m2 = None if mask2 is None else np.ascontiguousarray(mask2, np.uint8)
        Fm = np.ascontiguousarray(F12, np.float32).reshape(9)
        thr = float(self.settings.TH_LOW if thr is None else thr)
        n1, n2 = len(d1), len(d2)
        bi, b, s, acc = self._outs(n1)
        node1 = np.full(n1, -1, np.int32); node2 = np.full(max(n2, 1), -1, np.int32)
        F.check(F.lib().hyorb_search_for_triangulation_host(self._h, vocab._h, F.ptr(k1), F.ptr(d1), F.ptr(m1), n1, F.ptr(k2), F.ptr(d2), F.ptr(m2), n2,
                                                            int(levelsup), F.ptr(Fm), float(sigma_ref), float(size_ref), thr, float(ratio),
                                                            F.ptr(node1), F.ptr(node2), F.ptr(bi), F.ptr(b), F.ptr(s), F.ptr(acc)))
        return bi, b, s, acc, node1, node2[:n2]

    def ComputeDistinctiveDescriptors(self, desc, lm_off):
        """MapPointDBEntry::_computeDistinctiveDescriptor_ (src/core/MapPointDB.cpp:127-171) for many landmarks at once.
        desc: [total, 32] observation descriptors, lm_off: CSR offsets.  Returns (best_idx relative to each list, best_median)."""
        desc = np.ascontiguousarray(desc, np.uint8); lm_off = np.ascontiguousarray(lm_off, np.int32)
        n = len(lm_off) - 1
        bi = np.full(n, -1, np.int32); bm = np.full(n, -1, np.int32)
        F.check(F.lib().hyorb_distinctive_descriptor_host(self._h, F.ptr(desc), F.ptr(lm_off), n, F.ptr(bi), F.ptr(bm)))
        return bi, bm

    def match_bruteforce_device(self, d_q, nq, d_t, nt, rule, thr, ratio, d_best_idx, d_best, d_second, d_accepted):
        F.check(F.lib().hyorb_match_bruteforce_device(self._h, d_q, nq, d_t, nt, int(rule), float(thr), float(ratio),
                                                      d_best_idx, d_best, d_second, d_accepted))
