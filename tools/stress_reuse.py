"""Workspace-reuse stress: long-lived extractor handles see a random sequence of shapes, batch sizes and entry points
(single frame, frame batch, stereo batch), alone and from three host threads at once (one handle each).  Every image's
keypoints + descriptors must equal what a fresh handle returned for that image.
  python tools/stress_reuse.py [ops]"""
import hashlib, os, sys, threading, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import hyslam_b200 as hb
from hyslam_b200 import synth

ops = int(sys.argv[1]) if len(sys.argv) > 1 else 200
NF = 1200
shapes = [(480, 752), (376, 1241), (240, 320), (600, 800), (1080, 1920), (333, 517)]
NIMG = 8
imgs = {s: np.stack([synth.noise_image(s[0], s[1], 7 * k + 1) if k % 3 else synth.blocks_image(s[0], s[1], k) for k in range(NIMG)]) for s in shapes}
cams = {s: hb.StereoCamera(mbf=bench.CAM['mbf'], fx=bench.CAM['fx'], mnMaxY=float(s[0])) for s in shapes}


def digest(kps, desc):
    return hashlib.sha256(np.ascontiguousarray(kps).tobytes() + np.ascontiguousarray(desc).tobytes()).hexdigest()


ref = {}
for s in shapes:
    ex = hb.ORBExtractor(hb.FeatureExtractorSettings(nFeatures=NF))
    for k in range(NIMG):
        ref[(s, k)] = digest(*ex(imgs[s][k], None))
    ex.close()

bad = []


def worker(seed, n):
    rng = np.random.default_rng(seed)
    ex = hb.ORBExtractor(hb.FeatureExtractorSettings(nFeatures=NF))
    for it in range(n):
        s = shapes[int(rng.integers(len(shapes)))]
        op = int(rng.integers(3))
        if op == 0:
            k = int(rng.integers(NIMG))
            got = {k: digest(*ex(imgs[s][k], None))}
        elif op == 1:
            k0 = int(rng.integers(NIMG - 1)); nb = int(rng.integers(1, NIMG - k0 + 1))
            kps, desc, cnt = ex.extract_batch(imgs[s][k0:k0 + nb])
            got = {k0 + i: digest(kps[i, :cnt[i]], desc[i, :cnt[i]]) for i in range(nb)}
        else:
            k0 = 2 * int(rng.integers(NIMG // 2 - 1)); npair = int(rng.integers(1, (NIMG - k0) // 2 + 1))
            out = ex.process_stereo_batch(imgs[s][k0:k0 + 2 * npair], cams[s])
            kps, desc, cnt = out[0], out[1], out[2]
            got = {k0 + i: digest(kps[i, :cnt[i]], desc[i, :cnt[i]]) for i in range(2 * npair)}
        for k, d in got.items():
            if d != ref[(s, k)]:
                bad.append((seed, it, s, op, k))
    ex.close()


t0 = time.time()
worker(1, ops)
print(f"single handle: {ops} ops, {len(bad)} mismatches, {time.time() - t0:.1f}s", flush=True)
nb = len(bad)
ths = [threading.Thread(target=worker, args=(10 + i, ops)) for i in range(3)]
for t in ths: t.start()
for t in ths: t.join()
print(f"three threads: {3 * ops} ops, {len(bad) - nb} mismatches, {time.time() - t0:.1f}s", flush=True)
for b in bad[:10]:
    print("MISMATCH", b)
sys.exit(1 if bad else 0)
