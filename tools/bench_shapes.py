#!/usr/bin/env python
"""Informational timings of the other BASELINE.json shapes (not the headline metric): C1 752x480 mono 1000 features,
C3 3840x2160 8000 features, through the host-buffer batch entry point.  python tools/bench_shapes.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hyslam_b200 as hb
from hyslam_b200 import synth

for name, (h, w), nf, B in [("C1 752x480 / 1000 features", (480, 752), 1000, 64), ("C3 3840x2160 / 8000 features", (2160, 3840), 8000, 8),
                            ("C2 1241x376 / 2000 features, blocks images", (376, 1241), 2000, 64)]:
    kind = synth.blocks_image if "blocks" in name else synth.noise_image
    imgs = np.stack([kind(h, w, i) for i in range(min(B, 4))] * (B // min(B, 4)))
    ex = hb.ORBExtractor(hb.FeatureExtractorSettings(nFeatures=nf))
    cap = 6 * nf
    ex.extract_batch(imgs, capacity=cap)
    t0 = time.perf_counter(); reps = 5
    for _ in range(reps):
        kps, desc, counts = ex.extract_batch(imgs, capacity=cap)
    dt = (time.perf_counter() - t0) / reps
    one = imgs[0]
    ex(one, None, capacity=cap)
    t0 = time.perf_counter()
    for _ in range(20):
        ex(one, None, capacity=cap)
    d1 = (time.perf_counter() - t0) / 20
    print(f"{name}: batch of {B}: {B / dt:9.0f} frames/s ({1e3 * dt / B:.3f} ms/frame, {counts.mean():.0f} keypoints/frame); single synchronous call {1e3 * d1:.2f} ms")
    ex.close()
