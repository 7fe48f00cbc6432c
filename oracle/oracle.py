"""ctypes wrapper over oracle/_build/liborb_oracle.so -- the CPU ORACLE (test infrastructure).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product (hyslam_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liborb_oracle.so")

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28


class Params(C.Structure):
    _fields_ = [("nfeatures", C.c_int32), ("scale_factor", C.c_float), ("nlevels", C.c_int32),
                ("cell_px", C.c_int32), ("ini_th", C.c_int32), ("min_th", C.c_int32)]


class StereoParams(C.Structure):
    _fields_ = [("mbf", C.c_float), ("fx", C.c_float), ("n_rows", C.c_int32),
                ("th_high", C.c_float), ("th_low", C.c_float), ("size_ref", C.c_float)]


class Bounds(C.Structure):
    _fields_ = [("min_x", C.c_float), ("max_x", C.c_float), ("min_y", C.c_float), ("max_y", C.c_float)]


class Debug(C.Structure):
    _fields_ = [("pyramid", C.c_void_p), ("blurred", C.c_void_p),
                ("cand_x", C.c_void_p), ("cand_y", C.c_void_p), ("cand_resp", C.c_void_p),
                ("cand_count", C.c_void_p), ("cand_cap", C.c_int32), ("level_count", C.c_void_p)]


WQ_DTYPE = np.dtype([("u", "<f4"), ("v", "<f4"), ("r", "<f4"), ("size_lo", "<f4"), ("size_hi", "<f4"),
                     ("ur", "<f4"), ("ur_radius", "<f4")])


LM_DTYPE = np.dtype([("Pw", "<f4", 3), ("size", "<f4"), ("min_dist", "<f4"), ("max_dist", "<f4"), ("assoc_idx", "<i4")])


class Projection(C.Structure):
    """pose + camera of the frame landmarks are projected into (orc_projection)"""
    _fields_ = [("Rcw", C.c_float * 9), ("tcw", C.c_float * 3), ("Ow", C.c_float * 3), ("K", C.c_float * 9),
                ("mbf", C.c_float), ("stereo", C.c_int32), ("bounds", Bounds)]


def make_projection(Rcw, tcw, Ow, K, mbf, stereo, bounds, cls=None):
    pr = (cls or Projection)()
    pr.Rcw[:] = [float(v) for v in np.asarray(Rcw, np.float32).reshape(9)]
    pr.tcw[:] = [float(v) for v in np.asarray(tcw, np.float32).reshape(3)]
    pr.Ow[:] = [float(v) for v in np.asarray(Ow, np.float32).reshape(3)]
    pr.K[:] = [float(v) for v in np.asarray(K, np.float32).reshape(9)]
    pr.mbf = float(mbf); pr.stereo = int(stereo)
    pr.bounds.min_x, pr.bounds.max_x, pr.bounds.min_y, pr.bounds.max_y = [float(v) for v in bounds]
    return pr


def build(force=False):
    """Compile the oracle with gcc (idempotent)."""
    src = [os.path.join(_HERE, f) for f in ("orb_oracle.c", "quadtree_closed_form.c", "Makefile")]
    src.append(os.path.join(_HERE, "..", "include", "hyorb_brief_pattern.inc"))
    if not force and os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in src):
        return _SO
    env = dict(os.environ)
    env.pop("CC", None)
    subprocess.run(["make", "-C", _HERE, "CC=gcc"], check=True, capture_output=True, env=env)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def default_params(nfeatures=1000, scale_factor=1.2, nlevels=8, cell_px=30, ini_th=20, min_th=4):
    return Params(nfeatures, scale_factor, nlevels, cell_px, ini_th, min_th)


def scale_tables(p):
    n = p.nlevels
    s, i, s2, i2 = (np.zeros(n, np.float32) for _ in range(4))
    q = np.zeros(n, np.int32)
    rc = lib().orc_scale_tables(C.byref(p), _p(s), _p(i), _p(s2), _p(i2), _p(q))
    assert rc == 0, rc
    return s, i, s2, i2, q


def level_sizes(p, W, H):
    _, inv, _, _, _ = scale_tables(p)
    out = []
    for l in range(p.nlevels):
        w, h = C.c_int32(), C.c_int32()
        lib().orc_level_size(W, H, C.c_float(float(inv[l])), C.byref(w), C.byref(h))
        out.append((w.value, h.value))
    return out


def resize_linear(src, dw, dh):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty((dh, dw), np.uint8)
    rc = lib().orc_resize_linear_u8(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dw, dh, dw)
    assert rc == 0, rc
    return dst


def gaussian7(src):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty_like(src)
    rc = lib().orc_gaussian7_u8(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dst.strides[0])
    assert rc == 0, rc
    return dst


def fast9(img, threshold=20, nms=True):
    img = np.ascontiguousarray(img, np.uint8)
    cap = img.size // 2 + 16
    x, y, r = (np.empty(cap, np.float32) for _ in range(3))
    n = lib().orc_fast9(_p(img), img.shape[1], img.shape[0], img.strides[0], threshold, int(nms), _p(x), _p(y), _p(r), cap)
    assert n >= 0, n
    return x[:n].copy(), y[:n].copy(), r[:n].copy()


def detect_level(img, cell_px=30):
    img = np.ascontiguousarray(img, np.uint8)
    cap = img.size // 2 + 16
    x, y, r = (np.empty(cap, np.float32) for _ in range(3))
    n = lib().orc_detect_level(_p(img), img.shape[1], img.shape[0], img.strides[0], cell_px, _p(x), _p(y), _p(r), cap)
    if n < 0:
        raise RuntimeError(f"orc_detect_level rc={n}")
    return x[:n].copy(), y[:n].copy(), r[:n].copy()


def distribute_octtree(x, y, resp, min_x, max_x, min_y, max_y, N, closed_form=False):
    x = np.ascontiguousarray(x, np.float32); y = np.ascontiguousarray(y, np.float32)
    resp = np.ascontiguousarray(resp, np.float32)
    cap = len(x) + 16
    out = np.empty(cap, np.int32)
    fn = lib().orc_quadtree_closed_form if closed_form else lib().orc_distribute_octtree
    n = fn(_p(x), _p(y), _p(resp), len(x), min_x, max_x, min_y, max_y, N, _p(out), cap)
    if n < 0:
        raise RuntimeError(f"distribute rc={n}")
    return out[:n].copy()


def fast_atan2(y, x):
    f = lib().orc_fast_atan2
    f.restype = C.c_float
    return f(C.c_float(y), C.c_float(x))


def ic_angle(img, x, y):
    img = np.ascontiguousarray(img, np.uint8)
    f = lib().orc_ic_angle
    f.restype = C.c_float
    return f(_p(img), img.strides[0], C.c_float(x), C.c_float(y))


def brief(img, x, y, angle):
    img = np.ascontiguousarray(img, np.uint8)
    d = np.empty(32, np.uint8)
    lib().orc_brief(_p(img), img.strides[0], C.c_float(x), C.c_float(y), C.c_float(angle), _p(d))
    return d


def umax():
    u = np.zeros(16, np.int32)
    lib().orc_umax(_p(u))
    return u


def extract(img, p=None, cap=None, debug=False):
    """ORBExtractor::operator() restatement.  Returns (kps[KP_DTYPE], desc[n,32]) (+ debug dict)."""
    p = p or default_params()
    img = np.ascontiguousarray(img, np.uint8)
    H, W = img.shape
    cap = cap or max(4 * p.nfeatures + 1024, 4096)
    kps = np.zeros(cap, KP_DTYPE)
    desc = np.zeros((cap, 32), np.uint8)
    n = C.c_int32()
    dbg = None
    keep = {}
    if debug:
        sizes = level_sizes(p, W, H)
        tot = sum(w * h for w, h in sizes)
        ccap = sizes[0][0] * sizes[0][1] // 2 + 1024
        keep = dict(pyr=np.zeros(tot, np.uint8), blur=np.zeros(tot, np.uint8),
                    cx=np.zeros((p.nlevels, ccap), np.float32), cy=np.zeros((p.nlevels, ccap), np.float32),
                    cr=np.zeros((p.nlevels, ccap), np.float32), cc=np.zeros(p.nlevels, np.int32),
                    lc=np.zeros(p.nlevels, np.int32))
        dbg = Debug(keep["pyr"].ctypes.data, keep["blur"].ctypes.data, keep["cx"].ctypes.data, keep["cy"].ctypes.data,
                    keep["cr"].ctypes.data, keep["cc"].ctypes.data, ccap, keep["lc"].ctypes.data)
    rc = lib().orc_extract(C.byref(p), _p(img), W, H, img.strides[0], _p(kps), _p(desc), cap, C.byref(n),
                           C.byref(dbg) if dbg else None)
    if rc != 0:
        raise RuntimeError(f"orc_extract rc={rc}")
    k, d = kps[:n.value].copy(), desc[:n.value].copy()
    if not debug:
        return k, d
    sizes = level_sizes(p, W, H)
    offs = np.cumsum([0] + [w * h for w, h in sizes])
    info = dict(sizes=sizes,
                pyramid=[keep["pyr"][offs[l]:offs[l + 1]].reshape(sizes[l][1], sizes[l][0]) for l in range(p.nlevels)],
                blurred=[keep["blur"][offs[l]:offs[l + 1]].reshape(sizes[l][1], sizes[l][0]) for l in range(p.nlevels)],
                cand=[(keep["cx"][l, :keep["cc"][l]].copy(), keep["cy"][l, :keep["cc"][l]].copy(),
                       keep["cr"][l, :keep["cc"][l]].copy()) for l in range(p.nlevels)],
                level_count=keep["lc"].copy())
    return k, d, info


def extract_batch(images, p=None, cap=None, nthreads=1):
    p = p or default_params()
    images = np.ascontiguousarray(images, np.uint8)
    B, H, W = images.shape
    cap = cap or max(4 * p.nfeatures + 1024, 4096)
    kps = np.zeros((B, cap), KP_DTYPE)
    desc = np.zeros((B, cap, 32), np.uint8)
    counts = np.zeros(B, np.int32)
    rc = lib().orc_extract_batch(C.byref(p), _p(images), B, W, H, _p(kps), _p(desc), cap, _p(counts), nthreads)
    if rc != 0:
        raise RuntimeError(f"orc_extract_batch rc={rc}")
    return kps, desc, counts


def hamming(a, b):
    a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
    return lib().orc_hamming(_p(a), _p(b))


def epipolar_check(kps1, kps2, i1, i2, F12, sigma_ref=1.0, size_ref=31.0):
    """EpipolarConsistencyBoWCriterion::CheckDistEpipolarLine (MatchCriteria.cpp:659-676) for candidate pairs (i1[c], i2[c])."""
    k1 = np.ascontiguousarray(kps1, KP_DTYPE); k2 = np.ascontiguousarray(kps2, KP_DTYPE)
    i1 = np.ascontiguousarray(i1, np.int32); i2 = np.ascontiguousarray(i2, np.int32)
    Fm = np.ascontiguousarray(F12, np.float32).reshape(9)
    ok = np.zeros(len(i1), np.uint8)
    lib().orc_epipolar_check(_p(k1), _p(k2), _p(i1), _p(i2), len(i1), _p(Fm), C.c_float(sigma_ref), C.c_float(size_ref), _p(ok))
    return ok


def preprocess(img, rgb=True, half_scale=False):
    """ImageProcessing::PreProcessImg (ImageProcessing.cpp:118-138): scale 1.0 / 0.5 then RGB|BGR[A] -> gray.  img: HxW or HxWxC uint8."""
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape[:2]
    cn = 1 if img.ndim == 2 else img.shape[2]
    ow, oh = C.c_int(), C.c_int()
    if lib().orc_preprocess_size(w, h, int(half_scale), C.byref(ow), C.byref(oh)) != 0:
        raise ValueError("preprocess: unsupported size for the half-scale box path")
    out = np.empty((oh.value, ow.value), np.uint8)
    rc = lib().orc_preprocess(_p(img), w, h, img.strides[0], cn, int(rgb), int(half_scale), _p(out), out.strides[0])
    if rc != 0:
        raise RuntimeError(f"orc_preprocess rc={rc}")
    return out


def distinctive_descriptor(desc, lm_off):
    """MapPointDBEntry::_computeDistinctiveDescriptor_ (MapPointDB.cpp:127-171) over CSR landmark lists: (best_idx, best_median)."""
    desc = np.ascontiguousarray(desc, np.uint8); lm_off = np.ascontiguousarray(lm_off, np.int32)
    n = len(lm_off) - 1
    bi = np.empty(n, np.int32); bm = np.empty(n, np.int32)
    rc = lib().orc_distinctive_descriptor(_p(desc), _p(lm_off), n, _p(bi), _p(bm))
    if rc != 0:
        raise RuntimeError(f"orc_distinctive_descriptor rc={rc}")
    return bi, bm


def match_csr(qdesc, tdesc, cand_off=None, cand_idx=None, mode=0, thr=100.0, ratio=0.9):
    qdesc = np.ascontiguousarray(qdesc, np.uint8); tdesc = np.ascontiguousarray(tdesc, np.uint8)
    nq, nt = len(qdesc), len(tdesc)
    if cand_off is not None:
        cand_off = np.ascontiguousarray(cand_off, np.int32); cand_idx = np.ascontiguousarray(cand_idx, np.int32)
    bi = np.empty(nq, np.int32); b = np.empty(nq, np.uint16); s = np.empty(nq, np.uint16); acc = np.empty(nq, np.uint8)
    rc = lib().orc_match_csr(_p(qdesc), nq, _p(tdesc), nt, _p(cand_off), _p(cand_idx), mode,
                             C.c_float(thr), C.c_float(ratio), _p(bi), _p(b), _p(s), _p(acc))
    if rc != 0:
        raise RuntimeError(f"orc_match_csr rc={rc}")
    return bi, b, s, acc


def stereo_match(sp, kl, dl, kr, dr):
    kl = np.ascontiguousarray(kl); kr = np.ascontiguousarray(kr)
    dl = np.ascontiguousarray(dl, np.uint8); dr = np.ascontiguousarray(dr, np.uint8)
    nl = len(kl)
    uR = np.empty(nl, np.float32); depth = np.empty(nl, np.float32)
    br = np.empty(nl, np.int32); bd = np.empty(nl, np.int32)
    rc = lib().orc_stereo_match(C.byref(sp), _p(kl), _p(dl), nl, _p(kr), _p(dr), len(kr), _p(uR), _p(depth), _p(br), _p(bd))
    if rc != 0:
        raise RuntimeError(f"orc_stereo_match rc={rc}")
    return uR, depth, br, bd


def grid_build(kps, bounds):
    kps = np.ascontiguousarray(kps)
    off = np.zeros(64 * 48 + 1, np.int32); idx = np.zeros(max(len(kps), 1), np.int32)
    rc = lib().orc_grid_build(_p(kps), len(kps), C.byref(bounds), _p(off), _p(idx))
    assert rc == 0, rc
    return off, idx[:off[-1]].copy()


def grid_query(kps, bounds, off, idx, x, y, r):
    kps = np.ascontiguousarray(kps)
    idx = np.ascontiguousarray(idx, np.int32) if len(idx) else np.zeros(1, np.int32)
    out = np.empty(len(kps) + 1, np.int32)
    n = lib().orc_grid_query(_p(kps), C.byref(bounds), _p(off), _p(idx), C.c_float(x), C.c_float(y), C.c_float(r),
                             _p(out), len(out))
    assert n >= 0, n
    return out[:n].copy()


def match_window(kps, tdesc, t_uR, t_matched, bounds, off, idx, queries, qdesc, thr=100.0, ratio=0.9):
    kps = np.ascontiguousarray(kps); tdesc = np.ascontiguousarray(tdesc, np.uint8)
    queries = np.ascontiguousarray(queries, WQ_DTYPE); qdesc = np.ascontiguousarray(qdesc, np.uint8)
    t_uR = None if t_uR is None else np.ascontiguousarray(t_uR, np.float32)
    t_matched = None if t_matched is None else np.ascontiguousarray(t_matched, np.uint8)
    idx = np.ascontiguousarray(idx, np.int32) if len(idx) else np.zeros(1, np.int32)
    nq = len(queries)
    bi = np.empty(nq, np.int32); b = np.empty(nq, np.uint16); s = np.empty(nq, np.uint16); acc = np.empty(nq, np.uint8)
    rc = lib().orc_match_window(_p(kps), _p(tdesc), _p(t_uR), _p(t_matched), len(kps), C.byref(bounds), _p(off), _p(idx),
                                _p(queries), _p(qdesc), nq, C.c_float(thr), C.c_float(ratio), _p(bi), _p(b), _p(s), _p(acc))
    if rc != 0:
        raise RuntimeError(f"orc_match_window rc={rc}")
    return bi, b, s, acc


def project_landmarks(pr, lms, t_kps, th, size_ref=31.0, frac_smaller=0.5, frac_larger=1.5):
    """front half of FeatureMatcher::SearchByProjection(Frame&, landmarks, th): (window queries, passed flags)"""
    lms = np.ascontiguousarray(lms, LM_DTYPE); t_kps = np.ascontiguousarray(t_kps)
    n = len(lms)
    q = np.zeros(n, WQ_DTYPE); passed = np.zeros(n, np.uint8)
    rc = lib().orc_project_landmarks(C.byref(pr), _p(lms), n, _p(t_kps), len(t_kps), C.c_float(th), C.c_float(size_ref),
                                     C.c_float(frac_smaller), C.c_float(frac_larger), _p(q), _p(passed))
    if rc != 0:
        raise RuntimeError(f"orc_project_landmarks rc={rc}")
    return q, passed


def projection_rotation(best_idx, accepted, prev_angle, t_kps):
    """RotationConsistencyCriterion over projection matches (MatchCriteria.cpp:363-401): returns the updated accepted flags"""
    bi = np.ascontiguousarray(best_idx, np.int32); acc = np.ascontiguousarray(accepted, np.uint8).copy()
    pa = np.ascontiguousarray(prev_angle, np.float32); t_kps = np.ascontiguousarray(t_kps)
    rc = lib().orc_projection_rotation(_p(bi), _p(acc), len(bi), _p(pa), _p(t_kps), len(t_kps))
    if rc != 0:
        raise RuntimeError(f"orc_projection_rotation rc={rc}")
    return acc


def bow_transform(vocab, desc, levelsup=4):
    """DBoW2 TemplatedVocabulary::transform per feature (published algorithm; see orb_oracle.c): (word_id, node_id, weight).
    vocab: dict(L, child_off, child_idx, node_desc, word_of, weight_of)."""
    desc = np.ascontiguousarray(desc, np.uint8)
    n = len(desc)
    co = np.ascontiguousarray(vocab["child_off"], np.int32); ci = np.ascontiguousarray(vocab["child_idx"], np.int32)
    nd = np.ascontiguousarray(vocab["node_desc"], np.uint8); wo = np.ascontiguousarray(vocab["word_of"], np.int32)
    wt = np.ascontiguousarray(vocab["weight_of"], np.float32)
    w = np.empty(n, np.int32); nid = np.empty(n, np.int32); wgt = np.empty(n, np.float32)
    rc = lib().orc_bow_transform(len(co) - 1, int(vocab["L"]), _p(co), _p(ci), _p(nd), _p(wo), _p(wt), _p(desc), n, int(levelsup), _p(w), _p(nid), _p(wgt))
    if rc != 0:
        raise RuntimeError(f"orc_bow_transform rc={rc}")
    return w, nid, wgt


def rotation_consistency(angle_prev, angle_curr):
    a = np.ascontiguousarray(angle_prev, np.float32); b = np.ascontiguousarray(angle_curr, np.float32)
    keep = np.empty(len(a), np.uint8)
    rc = lib().orc_rotation_consistency(_p(a), _p(b), len(a), _p(keep))
    if rc != 0:
        raise RuntimeError(f"orc_rotation_consistency rc={rc}")
    return keep


# ---- round 2: remaining matcher entry points (Fuse, SearchBySim3, SearchForInitialization) ----
def viewing_angle(Ow, Pw, normal, max_angle=1.047):
    """ViewingAngleCriterionCore (MatchCriteria.cpp:94-110)"""
    Ow = np.ascontiguousarray(Ow, np.float32); Pw = np.ascontiguousarray(Pw, np.float32).reshape(-1, 3)
    normal = np.ascontiguousarray(normal, np.float32).reshape(-1, 3)
    out = np.empty(len(Pw), np.uint8)
    lib().orc_viewing_angle(_p(Ow), _p(Pw), _p(normal), len(Pw), C.c_float(max_angle), _p(out))
    return out


def match_window_ex(kps, tdesc, t_uR, t_matched, bounds, off, idx, queries, qdesc, thr, ratio, rule=0, q_active=None, reproj_thr=-1.0,
                    sigma_ref=1.0, size_ref=31.0):
    kps = np.ascontiguousarray(kps); tdesc = np.ascontiguousarray(tdesc, np.uint8)
    queries = np.ascontiguousarray(queries, WQ_DTYPE); qdesc = np.ascontiguousarray(qdesc, np.uint8)
    t_uR = None if t_uR is None else np.ascontiguousarray(t_uR, np.float32)
    t_matched = None if t_matched is None else np.ascontiguousarray(t_matched, np.uint8)
    q_active = None if q_active is None else np.ascontiguousarray(q_active, np.uint8)
    idx = np.ascontiguousarray(idx, np.int32) if len(idx) else np.zeros(1, np.int32)
    nq = len(queries)
    bi = np.empty(nq, np.int32); b = np.empty(nq, np.uint16); s = np.empty(nq, np.uint16); acc = np.empty(nq, np.uint8)
    rc = lib().orc_match_window_ex(_p(kps), _p(tdesc), _p(t_uR), _p(t_matched), len(kps), C.byref(bounds), _p(off), _p(idx), _p(queries), _p(qdesc),
                                   _p(q_active), nq, C.c_float(thr), C.c_float(ratio), int(rule), C.c_float(reproj_thr), C.c_float(sigma_ref),
                                   C.c_float(size_ref), _p(bi), _p(b), _p(s), _p(acc))
    if rc != 0:
        raise RuntimeError(f"orc_match_window_ex rc={rc}")
    return bi, b, s, acc


def project_sim3(R_a, t_a, sR_ba, t_ba, pr_b, lms, kps_b, th, size_ref=31.0):
    """one direction of SearchBySim3 (FeatureMatcher.cc:783-845): (window queries, passed flags)"""
    f = lambda a: np.ascontiguousarray(a, np.float32).reshape(-1)
    lms = np.ascontiguousarray(lms, LM_DTYPE); kps_b = np.ascontiguousarray(kps_b)
    n = len(lms)
    q = np.zeros(n, WQ_DTYPE); passed = np.zeros(n, np.uint8)
    Ra, ta, sR, tb = f(R_a), f(t_a), f(sR_ba), f(t_ba)
    rc = lib().orc_project_sim3(_p(Ra), _p(ta), _p(sR), _p(tb), C.byref(pr_b), _p(lms), n, _p(kps_b), len(kps_b), C.c_float(th), C.c_float(size_ref),
                                _p(q), _p(passed))
    if rc != 0:
        raise RuntimeError(f"orc_project_sim3 rc={rc}")
    return q, passed


def search_for_initialization(k1, d1, k2, d2, bounds, prev_xy, window, thr=50.0, ratio=0.9):
    """FeatureMatcher::SearchForInitialization (FeatureMatcher.cc:404-462): (n_matches, matches12, updated prev_xy)"""
    k1 = np.ascontiguousarray(k1); k2 = np.ascontiguousarray(k2)
    d1 = np.ascontiguousarray(d1, np.uint8); d2 = np.ascontiguousarray(d2, np.uint8)
    off, idx = grid_build(k2, bounds)
    idx = np.ascontiguousarray(idx, np.int32) if len(idx) else np.zeros(1, np.int32)
    pm = np.ascontiguousarray(prev_xy, np.float32).copy()
    m12 = np.empty(len(k1), np.int32)
    n = lib().orc_search_for_initialization(_p(k1), _p(d1), len(k1), _p(k2), _p(d2), len(k2), C.byref(bounds), _p(off), _p(idx), _p(pm), int(window),
                                            C.c_float(thr), C.c_float(ratio), _p(m12))
    if n < 0:
        raise RuntimeError(f"orc_search_for_initialization rc={n}")
    return n, m12, pm
