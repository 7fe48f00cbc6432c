"""Handle churn stress: create an extractor, extract a few frames, destroy it -- hundreds of times over changing shapes --
so that the allocator hands the same device addresses (workspace, tensor maps) to handles of different geometry.
Every (shape, settings, seed) result must be identical each time it comes round again.
  python tools/stress_handles.py [iterations]"""
import hashlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hyslam_b200 as hb
from hyslam_b200 import synth

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 400
shapes = [(480, 752, 1000), (376, 1241, 2000), (240, 320, 500), (600, 800, 1500), (160, 211, 300), (1080, 1920, 4000), (480, 640, 1000), (333, 517, 700)]
imgs = {s: [synth.noise_image(s[0], s[1], k) for k in range(2)] for s in shapes}
seen, bad = {}, 0
rng = np.random.default_rng(5)
t0 = time.time()
for it in range(iters):
    s = shapes[int(rng.integers(len(shapes)))]
    nl = int(rng.integers(3, 6 if s[0] < 200 else 9))
    try:
        ex = hb.ORBExtractor(hb.FeatureExtractorSettings(nFeatures=s[2], nLevels=nl))
        for k in range(2):
            kps, desc = ex(imgs[s][k], None)
            d = hashlib.sha256(np.ascontiguousarray(desc).tobytes() + np.ascontiguousarray(kps).tobytes()).hexdigest()
            key = (s, nl, k)
            if seen.setdefault(key, d) != d:
                bad += 1
                print("MISMATCH", key, flush=True)
        ex.close()
    except Exception as e:          # noqa: BLE001 -- report and stop: the context is gone after a device fault
        print("ERROR at iteration", it, s, nl, e, flush=True)
        sys.exit(2)
print(f"{iters} handle cycles, {len(seen)} distinct configs, {bad} mismatches, {time.time() - t0:.1f}s")
sys.exit(1 if bad else 0)
