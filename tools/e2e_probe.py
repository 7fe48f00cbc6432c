"""End-to-end probe: pairs/s of hyorb_process_stereo_batch_host for host-lane settings, one or more host threads (one
handle each).   python tools/e2e_probe.py [pairs]"""
import os, sys, threading, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import hyslam_b200 as hb
from hyslam_b200 import synth, _ffi as F

P = int(sys.argv[1]) if len(sys.argv) > 1 else 128
H, W, NF = 376, 1241, 2000
B = 2 * P
imgs = [torch.from_numpy(np.stack([synth.noise_image(H, W, 100 * r + i) for i in range(B)])).pin_memory() for r in range(4)]
cam = hb.StereoCamera(**bench.CAM)
cap = 2560

def make():
    ex = hb.ORBExtractor(hb.FeatureExtractorSettings(nFeatures=NF))
    o = (np.empty((B, cap), F.KP_DTYPE), np.empty((B, cap, 32), np.uint8), np.zeros(B, np.int32), np.empty((P, cap), np.float32), np.empty((P, cap), np.float32))
    pin = [torch.from_numpy(x.view(np.uint8).reshape(-1)).pin_memory() for x in o]
    return ex, tuple(po.numpy().view(x.dtype).reshape(x.shape) for po, x in zip(pin, o)), pin

handles = [make() for _ in range(3)]

def run(lanes, grade, workers, n=16):
    for ex, outs, _ in handles[:workers]:
        ex.set_pipelining(host_lanes=lanes)
        for _ in range(2): ex.process_stereo_batch(imgs[0].numpy(), cam, capacity=cap, out=outs)
    start = threading.Barrier(workers + 1)
    def work(wi):
        ex, outs, _ = handles[wi]
        start.wait()
        for i in range(n): ex.process_stereo_batch(imgs[(i * workers + wi) & 3].numpy(), cam, capacity=cap, out=outs)
    ths = [threading.Thread(target=work, args=(wi,)) for wi in range(workers)]
    for t in ths: t.start()
    start.wait(); t0 = time.perf_counter()
    for t in ths: t.join()
    dt = time.perf_counter() - t0
    print(f"workers {workers} lanes {lanes:2d} grade {grade or '-':9s} {P * n * workers / dt:9.0f} pairs/s  {dt / n * 1e3:.3f} ms/round", flush=True)

for rep in range(2):
    for workers in (1, 2):
        for lanes in (2, 4, 8, 12):
            run(lanes, None, workers)
