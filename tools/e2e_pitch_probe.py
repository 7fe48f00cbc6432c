#!/usr/bin/env python
"""End-to-end rate of hyorb_process_stereo_batch_host with DENSE host frames (1241-byte rows: uploaded flat, repacked on the device by
k_repack) against PITCHED host frames (row pitch 1248 = a multiple of 16: uploaded flat and read in place through TMA).  Two host
threads, one handle each, 4 lanes, pinned buffers, 128 pairs per call.   python tools/e2e_pitch_probe.py"""
import os, sys, threading, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as Bn
import hyslam_b200 as hb
from hyslam_b200 import _ffi as F

P, cap = 128, 2560
W, H = Bn.W, Bn.H
dense = Bn.make_pairs(P)
WP = (W + 15) & ~15
pitched_store = torch.zeros((2 * P, H, WP), dtype=torch.uint8).pin_memory()
pitched_store.numpy()[:, :, :W] = dense
inputs = {"dense": torch.from_numpy(dense).pin_memory().numpy(), "pitched": pitched_store.numpy()[:, :, :W]}
cam = hb.StereoCamera(**Bn.CAM)


def outs():
    o = (np.empty((2 * P, cap), F.KP_DTYPE), np.empty((2 * P, cap, 32), np.uint8), np.zeros(2 * P, np.int32), np.empty((P, cap), np.float32), np.empty((P, cap), np.float32))
    pin = [torch.from_numpy(x.view(np.uint8).reshape(-1)).pin_memory() for x in o]
    return tuple(po.numpy().view(x.dtype).reshape(x.shape) for po, x in zip(pin, o)), pin


res = {}
for name, img in inputs.items():
    hs = [hb.ORBExtractor(hb.FeatureExtractorSettings(nFeatures=Bn.NFEAT)) for _ in range(2)]
    os_ = [outs() for _ in range(2)]
    for h_, (o, _) in zip(hs, os_):
        h_.set_pipelining(host_lanes=4)
        h_.process_stereo_batch(img, cam, capacity=cap, out=o)
    steps = 20

    def work(i):
        for _ in range(steps):
            hs[i].process_stereo_batch(img, cam, capacity=cap, out=os_[i][0])
    ths = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for t in ths: t.start()
    for t in ths: t.join()
    torch.cuda.synchronize()
    res[name] = 2 * steps * P / (time.perf_counter() - t0)
    counts = os_[0][0][2].copy()
    for h_ in hs: h_.close()
    res[name + "_kps"] = int(counts.sum())
print(res)
