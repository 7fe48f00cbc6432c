"""Per-frame latency of the single-image entry point (the drop-in's use inside hySLAM): wall time per call and per-stage CUDA-event times."""
import os, sys, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hyslam_b200 as hb
from hyslam_b200 import synth
for (h,w,nf) in [(480,752,1000),(376,1241,2000)]:
    img=synth.noise_image(h,w,1)
    ex=hb.ORBExtractor(hb.FeatureExtractorSettings(nFeatures=nf))
    for _ in range(5): ex(img,None)
    ex.set_profiling(True); ex.stage_times(reset=True)
    t0=time.perf_counter()
    for _ in range(50): ex(img,None)
    dt=(time.perf_counter()-t0)/50
    st,calls=ex.stage_times(reset=True)
    print(h,w,"wall ms",round(dt*1e3,3),"stages ms",{k:round(v/calls,3) for k,v in st.items()},"sum",round(sum(st.values())/calls,3))

# one stereo pair: two single-frame extractions + host-side Stereomatcher vs the fused call
import bench
cam = hb.StereoCamera(**bench.CAM)
L, R = synth.stereo_pair(376, 1241, 5)
ex1 = hb.ORBExtractor(hb.FeatureExtractorSettings(nFeatures=2000)); ex2 = hb.ORBExtractor(hb.FeatureExtractorSettings(nFeatures=2000))
pair = np.stack([L, R])
def separate():
    kl, dl = ex1(L, None); kr, dr = ex2(R, None)
    sm = hb.Stereomatcher((kl, dl, kr, dr), cam)          # the reference constructs one per frame (ImageProcessing.cpp:100)
    sm.computeStereoMatches()
    out = sm.getData()
    sm.close()
    return out
def fused():
    return ex1.process_stereo_batch(pair, cam)
for name, fn in (("separate (2 x extract + stereo matcher)", separate), ("fused (one call)", fused)):
    try:
        for _ in range(5): fn()
        t0 = time.perf_counter()
        for _ in range(50): fn()
        print("stereo pair,", name, "ms per pair", round((time.perf_counter() - t0) / 50 * 1e3, 3))
    except Exception as e:          # noqa: BLE001
        print("stereo pair,", name, "not measured:", e)
