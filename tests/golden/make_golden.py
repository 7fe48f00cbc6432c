#!/usr/bin/env python
"""Generates tests/golden/*.npz -- golden vectors for the ORB front-end path.

The reference (bmhopkinson/hyslam) is C++ against OpenCV 3.4 and cannot be built here, and it
has no tests on this path.  This script is an INDEPENDENT second implementation
of the same behaviour used to pin the C oracle (oracle/orb_oracle.c):

  * pixel arithmetic comes from the real OpenCV library (cv2 4.13: cv2.resize INTER_LINEAR,
    cv2.FastFeatureDetector(20, NMS, TYPE_9_16) on each reference cell ROI, cv2.GaussianBlur 7x7
    sigma 2 REFLECT_101, cv2.fastAtan2) -- the same kernels the reference links;
  * the reference's own logic (scale tables, cell lattice, quadtree with std::list semantics and
    the canonical creation-order tie policy, IC_Angle, rotated BRIEF, Hamming scans, stereo
    row-band matcher, 64x48 grid, rotation histogram) is restated in plain Python/numpy from
    src/features/ORBExtractor.cpp, src/features/low_level/ORBFinder.cpp,
    src/features/Stereomatcher.cpp, src/features/MatchCriteria.cpp, src/core/Frame.cc.

It does not import the C oracle.  Run from the repo root:  python tests/golden/make_golden.py
"""
import hashlib
import math
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from hyslam_b200 import synth  # noqa: E402

cv2.setNumThreads(1)
try:
    cv2.ipp.setUseIPP(False)
except Exception:
    pass

f32 = np.float32
KP = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
               ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])


def cv_round(v):
    return int(np.rint(f32(v)))


def load_pattern():
    txt = open(os.path.join(HERE, "..", "..", "include", "hyorb_brief_pattern.inc")).read()
    body = txt.split("*/", 1)[1]
    vals = [int(t) for t in body.replace("\n", "").split(",") if t.strip()]
    assert len(vals) == 1024
    return np.array(vals, np.int32).reshape(512, 2)


PATTERN = load_pattern()


def scale_tables(nfeatures, scale_factor, nlevels):
    sf = float(f32(scale_factor))  # double scaleFactor initialised from the float setting
    scale = [f32(1.0)]
    for _ in range(1, nlevels):
        scale.append(f32(float(scale[-1]) * sf))
    inv = [f32(1.0) / s for s in scale]
    factor = f32(1.0 / sf)
    nd = f32(nfeatures) * (f32(1) - factor) / (f32(1) - f32(math.pow(float(factor), float(nlevels))))
    quota, tot = [], 0
    for _ in range(nlevels - 1):
        q = cv_round(nd)
        quota.append(q)
        tot += q
        nd = f32(nd * factor)
    quota.append(max(nfeatures - tot, 0))
    return scale, inv, quota


def umax_table():
    hp = 15
    umax = [0] * (hp + 1)
    vmax = int(math.floor(hp * math.sqrt(2.0) / 2 + 1))
    vmin = int(math.ceil(hp * math.sqrt(2.0) / 2))
    for v in range(vmax + 1):
        umax[v] = int(np.rint(math.sqrt(hp * hp - v * v)))
    v0 = 0
    for v in range(hp, vmin - 1, -1):
        while umax[v0] == umax[v0 + 1]:
            v0 += 1
        umax[v] = v0
        v0 += 1
    return umax


UMAX = umax_table()
FAST = cv2.FastFeatureDetector_create(20, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)


class Node:
    __slots__ = ("ulx", "uly", "urx", "bry", "keys", "nomore", "seq", "alive")


def divide(n, seqgen):
    hx = int(math.ceil(f32(n.urx - n.ulx) / 2))
    hy = int(math.ceil(f32(n.bry - n.uly) / 2))
    mx, my = n.ulx + hx, n.uly + hy
    ch = []
    boxes = [(n.ulx, n.uly, mx, my), (mx, n.uly, n.urx, my), (n.ulx, my, mx, n.bry), (mx, my, n.urx, n.bry)]
    buckets = [[], [], [], []]
    for k in n.keys:
        x, y = k[0], k[1]
        if x < mx:
            buckets[0 if y < my else 2].append(k)
        else:
            buckets[1 if y < my else 3].append(k)
    for b, box in zip(buckets, boxes):
        c = Node()
        c.ulx, c.uly, c.urx, c.bry = box
        c.keys = b
        c.nomore = len(b) == 1
        c.seq = next(seqgen)
        c.alive = True
        ch.append(c)
    return ch


def distribute(cands, minx, maxx, miny, maxy, N):
    """cands: list of (x, y, resp, order).  std::list emulated with a python list (front = index 0)."""
    if not cands:
        return []
    # C round(): half away from zero
    r = float(f32(maxx - minx) / f32(maxy - miny))
    nini = int(math.floor(r + 0.5))
    hx = f32(maxx - minx) / f32(nini)

    def gen():
        i = 0
        while True:
            yield i
            i += 1
    seqgen = gen()
    roots = []
    for i in range(nini):
        n = Node()
        n.ulx = int(hx * f32(i)); n.urx = int(hx * f32(i + 1)); n.uly = 0; n.bry = maxy - miny
        n.keys = []; n.nomore = False; n.seq = next(seqgen); n.alive = True
        roots.append(n)
    for k in cands:
        roots[int(f32(k[0]) / hx)].keys.append(k)
    lst = []
    for n in roots:
        if len(n.keys) == 1:
            n.nomore = True
            lst.append(n)
        elif len(n.keys) > 1:
            lst.append(n)
    finish = False
    while not finish:
        prev = len(lst)
        vsz = []
        n_expand = 0
        old = lst
        front = []   # pushed to the front: newest first
        kept = []
        for n in old:
            if n.nomore:
                kept.append(n)
                continue
            for c in divide(n, seqgen):
                if c.keys:
                    front.insert(0, c)
                    if len(c.keys) > 1:
                        n_expand += 1
                        vsz.append((len(c.keys), c.seq, c))
        lst = front + kept
        if len(lst) >= N or len(lst) == prev:
            finish = True
        elif len(lst) + 3 * n_expand > N:
            while not finish:
                prev = len(lst)
                pv = sorted(vsz, key=lambda t: (t[0], t[1]))
                vsz = []
                for j in range(len(pv) - 1, -1, -1):
                    node = pv[j][2]
                    for c in divide(node, seqgen):
                        if c.keys:
                            lst.insert(0, c)
                            if len(c.keys) > 1:
                                vsz.append((len(c.keys), c.seq, c))
                    lst.remove(node)
                    if len(lst) >= N:
                        break
                if len(lst) >= N or len(lst) == prev:
                    finish = True
    out = []
    for n in lst:
        best = n.keys[0]
        for k in n.keys[1:]:
            if k[2] > best[2]:
                best = k
        out.append(best)
    return out


def detect_level(img, cell_px):
    rows, cols = img.shape
    minb = 16
    maxbx, maxby = cols - 16, rows - 16
    width, height = f32(maxbx - minb), f32(maxby - minb)
    ncols, nrows = int(width / f32(cell_px)), int(height / f32(cell_px))
    wcell, hcell = int(math.ceil(width / ncols)), int(math.ceil(height / nrows))
    out = []
    for i in range(nrows):
        iniy = minb + i * hcell
        maxy = iniy + hcell + 6
        if iniy >= maxby - 3:
            continue
        maxy = min(maxy, maxby)
        for j in range(ncols):
            inix = minb + j * wcell
            maxx = inix + wcell + 6
            if inix >= maxbx - 6:
                continue
            maxx = min(maxx, maxbx)
            roi = np.ascontiguousarray(img[iniy:maxy, inix:maxx])
            for k in FAST.detect(roi):
                out.append((f32(k.pt[0] + j * wcell), f32(k.pt[1] + i * hcell), f32(k.response), len(out)))
    return out


def ic_angle(img, x, y):
    cx, cy = cv_round(x), cv_round(y)
    m01 = m10 = 0
    for u in range(-15, 16):
        m10 += u * int(img[cy, cx + u])
    for v in range(1, 16):
        d = UMAX[v]
        vs = 0
        for u in range(-d, d + 1):
            p, m = int(img[cy + v, cx + u]), int(img[cy - v, cx + u])
            vs += p - m
            m10 += u * (p + m)
        m01 += v * vs
    return f32(cv2.fastAtan2(float(m01), float(m10)))


def brief(img, x, y, angle):
    ang = f32(angle) * f32(math.pi / 180.0)
    a, b = f32(math.cos(float(ang))), f32(math.sin(float(ang)))
    cx, cy = cv_round(x), cv_round(y)
    px, py = PATTERN[:, 0].astype(f32), PATTERN[:, 1].astype(f32)
    rr = np.rint(px * b + py * a).astype(np.int64)   # separate fp32 mul / add, half-to-even
    cc = np.rint(px * a - py * b).astype(np.int64)
    vals = img[cy + rr, cx + cc].astype(np.int32)
    bits = (vals[0::2] < vals[1::2]).astype(np.uint8)
    return np.packbits(bits, bitorder="little")


def extract(img, nfeatures=1000, scale_factor=1.2, nlevels=8, cell_px=30):
    H, W = img.shape
    scale, inv, quota = scale_tables(nfeatures, scale_factor, nlevels)
    pyr = [img]
    for l in range(1, nlevels):
        w, h = cv_round(f32(W) * inv[l]), cv_round(f32(H) * inv[l])
        pyr.append(cv2.resize(pyr[-1], (w, h), interpolation=cv2.INTER_LINEAR))
    kps, descs, lcount, ccount = [], [], [], []
    for l in range(nlevels):
        im = pyr[l]
        cands = detect_level(im, cell_px)
        ccount.append(len(cands))
        sel = distribute(cands, 16, im.shape[1] - 16, 16, im.shape[0] - 16, quota[l])
        lcount.append(len(sel))
        if not sel:
            continue
        blur = cv2.GaussianBlur(im.copy(), (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
        size = f32(int(f32(31) * scale[l]))
        for (x, y, r, _o) in sel:
            x, y = f32(x + f32(16)), f32(y + f32(16))
            ang = ic_angle(blur, x, y)
            descs.append(brief(blur, x, y, ang))
            if l:
                x, y = f32(x * scale[l]), f32(y * scale[l])
            kps.append((x, y, size, ang, r, l, -1))
    k = np.array(kps, KP) if kps else np.zeros(0, KP)
    d = np.stack(descs) if descs else np.zeros((0, 32), np.uint8)
    return k, d, np.array(lcount, np.int32), np.array(ccount, np.int32), pyr


POP = np.array([bin(i).count("1") for i in range(256)], np.int32)


def hamming_matrix(a, b):
    return POP[a[:, None, :] ^ b[None, :, :]].sum(2)


def scan(dists, order):
    bd = bd2 = float(np.finfo(np.float32).max)
    bi = -1
    for t in order:
        d = float(dists[t])
        if d < bd:
            bd2, bd, bi = bd, d, t
        elif d < bd2:
            bd2 = d
    return bi, bd, bd2


def accept(mode, bd, bd2, thr, ratio):
    bd, bd2, thr, ratio = f32(bd), f32(bd2), f32(thr), f32(ratio)
    with np.errstate(over="ignore"):
        if mode == 0:
            return bool(bd <= thr and not (bd > ratio * bd2))
        if mode == 1:
            return bool(bd < thr and bd < ratio * bd2)
        return bool(bd <= thr and bd < bd2 * ratio)


def stereo(kl, dl, kr, dr, mbf, fx, n_rows, th_high=100.0, th_low=50.0, size_ref=31.0):
    nl = len(kl)
    uR = np.full(nl, -1, f32); depth = np.full(nl, -1, f32)
    thr = f32((f32(th_high) + f32(th_low)) / 2)
    rows = [[] for _ in range(n_rows)]
    for ir in range(len(kr)):
        r = f32(2.0) * kr["size"][ir] / f32(size_ref)
        for y in range(int(math.floor(kr["y"][ir] - r)), int(math.ceil(kr["y"][ir] + r)) + 1):
            rows[y].append(ir)
    mb = f32(mbf) / f32(fx)
    maxd = f32(mbf) / mb
    vd = []
    for il in range(nl):
        ul, vl, lv = kl["x"][il], kl["y"][il], kl["octave"][il]
        cand = rows[int(vl)]
        if not cand:
            continue
        minu, maxu = f32(ul - maxd), ul
        if maxu < 0:
            continue
        best, bi = f32(th_high), 0
        for ir in cand:
            if kr["octave"][ir] < lv - 1 or kr["octave"][ir] > lv + 1:
                continue
            u = kr["x"][ir]
            if minu <= u <= maxu:
                d = f32(POP[dl[il] ^ dr[ir]].sum())
                if d < best:
                    best, bi = d, ir
        if best < thr:
            ur0 = kr["x"][bi]
            disp = f32(ul - ur0)
            if 0 <= disp < maxd:
                if disp <= 0:
                    disp = f32(0.01)
                    ur0 = f32(float(ul) - 0.01)
                depth[il] = f32(mbf) / disp
                uR[il] = ur0
                vd.append((float(best), il))
    if vd:
        vd.sort()
        med = f32(vd[len(vd) // 2][0])
        th = f32(1.5) * f32(1.4) * med
        for d, il in reversed(vd):
            if f32(d) < th:
                break
            uR[il] = -1; depth[il] = -1
    return uR, depth


def grid_build(k, minx, maxx, miny, maxy):
    invw, invh = f32(64) / f32(maxx - minx), f32(48) / f32(maxy - miny)
    cells = [[[] for _ in range(48)] for _ in range(64)]

    def c_round(v):  # round half away from zero
        return int(math.floor(abs(float(v)) + 0.5) * (1 if v >= 0 else -1))
    for i in range(len(k)):
        px, py = c_round((k["x"][i] - f32(minx)) * invw), c_round((k["y"][i] - f32(miny)) * invh)
        if 0 <= px < 64 and 0 <= py < 48:
            cells[px][py].append(i)
    return cells, invw, invh


def grid_query(k, cells, invw, invh, minx, miny, x, y, r):
    x, y, r = f32(x), f32(y), f32(r)
    x0 = max(0, int(math.floor((x - f32(minx) - r) * invw)))
    x1 = min(63, int(math.ceil((x - f32(minx) + r) * invw)))
    y0 = max(0, int(math.floor((y - f32(miny) - r) * invh)))
    y1 = min(47, int(math.ceil((y - f32(miny) + r) * invh)))
    if x0 >= 64 or x1 < 0 or y0 >= 48 or y1 < 0:
        return []
    out = []
    for ix in range(x0, x1 + 1):
        for iy in range(y0, y1 + 1):
            for j in cells[ix][iy]:
                if abs(f32(k["x"][j] - x)) < r and abs(f32(k["y"][j] - y)) < r:
                    out.append(j)
    return out


def rotation_consistency(ap, ac):
    L = 30
    factor = f32(1.0) / f32(L)
    bins = []
    for p, c in zip(ap, ac):
        rot = f32(p - c)
        if rot < 0:
            rot = f32(rot + f32(360.0))
        v = float(rot * factor)
        b = int(math.floor(v + 0.5))
        if b == L:
            b = 0
        bins.append(b)
    hist = np.bincount(bins, minlength=L)
    m1 = m2 = m3 = 0
    i1 = i2 = i3 = -1
    for i in range(L):
        s = int(hist[i])
        if s > m1:
            m3, m2, m1, i3, i2, i1 = m2, m1, s, i2, i1, i
        elif s > m2:
            m3, m2, i3, i2 = m2, s, i2, i
        elif s > m3:
            m3, i3 = s, i
    if m2 < f32(0.1) * f32(m1):
        i2 = i3 = -1
    elif m3 < f32(0.1) * f32(m1):
        i3 = -1
    return np.array([b in (i1, i2, i3) for b in bins], np.uint8)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    out = {}
    # --- extraction cases -------------------------------------------------------------------
    cases = [
        ("ext_noise_320x240", "noise", 240, 320, 1, dict(nfeatures=500, scale_factor=1.2, nlevels=8, cell_px=30)),
        ("ext_blocks_376x280", "blocks", 280, 376, 2, dict(nfeatures=300, scale_factor=1.2, nlevels=6, cell_px=30)),
        ("ext_blocks_640x360", "blocks", 360, 640, 5, dict(nfeatures=1500, scale_factor=1.2, nlevels=8, cell_px=30)),
        ("ext_c1_noise_752x480", "noise", 480, 752, 0, dict(nfeatures=1000, scale_factor=1.2, nlevels=8, cell_px=30)),
    ]
    for name, kind, h, w, seed, prm in cases:
        img = synth.noise_image(h, w, seed) if kind == "noise" else synth.blocks_image(h, w, seed)
        k, d, lc, cc, pyr = extract(img, **prm)
        print(name, "kps", len(k), "per-level", lc.tolist(), "cands", cc.tolist())
        np.savez_compressed(os.path.join(HERE, name + ".npz"), kind=kind, h=h, w=w, seed=seed,
                            image_sha=sha(img), kps=k, desc=d, level_count=lc, cand_count=cc,
                            pyr_sha=np.array([sha(p) for p in pyr]), **{k_: v for k_, v in prm.items()})
    # --- stereo case ------------------------------------------------------------------------
    h, w, seed = 240, 480, 7
    L, R = synth.stereo_pair(h, w, seed)
    prm = dict(nfeatures=600, scale_factor=1.2, nlevels=8, cell_px=30)
    kl, dl, _, _, _ = extract(L, **prm)
    kr, dr, _, _, _ = extract(R, **prm)
    mbf, fx = 386.1448, 718.856
    uR, depth = stereo(kl, dl, kr, dr, mbf, fx, h)
    print("stereo matches", int((uR >= 0).sum()), "of", len(kl))
    np.savez_compressed(os.path.join(HERE, "stereo_noise_480x240.npz"), h=h, w=w, seed=seed, mbf=mbf, fx=fx,
                        left_sha=sha(L), right_sha=sha(R), kl=kl, dl=dl, kr=kr, dr=dr, uR=uR, depth=depth, **prm)
    # --- descriptor matching cases ------------------------------------------------------------
    a = synth.random_descriptors(300, 11)
    b, perm = synth.perturbed_descriptors(a, 12)
    D = hamming_matrix(a, b)
    res = {}
    for mode, thr, ratio in [(0, 100.0, 0.9), (1, 50.0, 1.0), (1, 50.0, 0.6), (2, 50.0, 0.9)]:
        bi, bd, bs, ac = [], [], [], []
        for q in range(len(a)):
            i, d1, d2 = scan(D[q], range(len(b)))
            bi.append(i); bd.append(int(d1)); bs.append(int(d2) if d2 < 1e30 else 65535)
            ac.append(accept(mode, d1, d2, thr, ratio))
        res[f"m{mode}_{thr}_{ratio}"] = np.stack([np.array(bi), np.array(bd), np.array(bs), np.array(ac, np.int64)])
    # CSR candidate lists (BoW-node style): random subsets incl. empty and singleton lists
    rng = np.random.default_rng(13)
    off, idx = [0], []
    for q in range(len(a)):
        n = int(rng.choice([0, 1, 2, 5, 40]))
        idx += rng.choice(len(b), size=n, replace=False).tolist()
        off.append(len(idx))
    bi, bd, bs, ac = [], [], [], []
    for q in range(len(a)):
        i, d1, d2 = scan(D[q], idx[off[q]:off[q + 1]])
        bi.append(i); bd.append(int(d1) if i >= 0 else 65535); bs.append(int(d2) if d2 < 1e30 else 65535)
        ac.append(i >= 0 and accept(1, d1, d2, 50.0, 0.6))
    np.savez_compressed(os.path.join(HERE, "match_300.npz"), a=a, b=b, csr_off=np.array(off, np.int32),
                        csr_idx=np.array(idx, np.int32),
                        csr_res=np.stack([np.array(bi), np.array(bd), np.array(bs), np.array(ac, np.int64)]), **res)
    # --- grid / window / rotation ---------------------------------------------------------------
    k = kl
    minx, maxx, miny, maxy = 0.0, float(w), 0.0, float(h)
    cells, invw, invh = grid_build(k, minx, maxx, miny, maxy)
    flat_off, flat_idx = [0], []
    for ix in range(64):
        for iy in range(48):
            flat_idx += cells[ix][iy]
            flat_off.append(len(flat_idx))
    rngq = np.random.default_rng(17)
    queries, qres = [], []
    for _ in range(60):
        x, y, r = float(rngq.uniform(-20, w + 20)), float(rngq.uniform(-20, h + 20)), float(rngq.uniform(1, 60))
        queries.append((x, y, r))
        qres.append(grid_query(k, cells, invw, invh, minx, miny, x, y, r))
    qoff = np.cumsum([0] + [len(q) for q in qres]).astype(np.int32)
    ap = rngq.uniform(0, 360, 400).astype(f32); ac_ = (ap + rngq.normal(20, 25, 400)).astype(f32) % f32(360)
    keep = rotation_consistency(ap, ac_)
    np.savez_compressed(os.path.join(HERE, "grid_rotation.npz"), kps=k, bounds=np.array([minx, maxx, miny, maxy], f32),
                        cell_off=np.array(flat_off, np.int32), cell_idx=np.array(flat_idx, np.int32),
                        queries=np.array(queries, f32), q_off=qoff,
                        q_idx=np.array([j for q in qres for j in q], np.int32),
                        angle_prev=ap, angle_curr=ac_, keep=keep)
    print("done")


if __name__ == "__main__":
    main()
