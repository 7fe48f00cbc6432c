#!/usr/bin/env python
"""bench.py -- ORB stereo front end throughput on B200 (BASELINE.json metric: ORB frames/s, extract + match).

Workload (config.workload): C2 -- KITTI-shaped 1241x376 stereo pairs, 2000 features per image, 8 levels, scale 1.2,
extract(left) + extract(right) + ComputeStereoMatches.  One "frame" = one stereo pair.  One "step" = one batch of
`--pairs` pairs through hyorb_process_stereo_batch_* (ImageProcessing::ProcessStereoImage, ImageProcessing.cpp:69-116).

  value : frames/s with the batch already resident in HBM (device pointers in, device pointers out), CUDA events on
          the launching stream, max over ranks.
  e2e   : the same through the host-buffer C-ABI call (pinned host images in, host keypoints/descriptors/uR/depth
          out), H2D and D2H inside the timed region.
  roofline : dominant kernel (FAST detection) algorithmic bytes / its live CUDA-event duration vs MEASURED_PEAKS hbm_gbs.
  cpu_baseline : the reference's OWN code (oracle/_ref: hySLAM's ORBExtractor / ORBFinder / Stereomatcher translation units
          compiled unmodified over oracle/cvshim's SIMD OpenCV primitives) on a bounded sample of the same workload, all host
          threads; next to it the reference-shaped run (2 threads per pair, ImageProcessing.cpp:82-84), the scalar C port and
          the cost of the same pixel work through cv2's own primitives.

`--impl reference` times that same reference code on the box's host cores with the same metric/config (kind "reference";
it falls back to the scalar oracle port, kind "port", only if oracle/_ref did not travel).
Multi-GPU: one process per GPU (torchrun), frames sharded by rank, no data-path collective (weak scaling).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

# hardware queues for the lanes' streams: must be in the environment before the CUDA context exists (hyslam_b200/csrc/api.cu)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, NFEAT = 1241, 376, 2000
CAM = dict(mbf=386.1448, fx=718.856, mnMaxY=376.0)
METRIC = "orb_stereo_frames_per_sec"
UNIT = "frames/s"


def workload_name(pairs):
    return f"C2: {W}x{H} stereo pairs, {NFEAT} features/image, 8 levels, scale 1.2, extract L+R + ComputeStereoMatches; {pairs} pairs per step"


def make_pairs(n_pairs, seed0=1000, distinct=8, first=0):
    """[2*n_pairs, H, W] uint8, (left, right) interleaved.  `distinct` pairs come from the seeded generators
    (hyslam_b200/synth.py, G-noise); the rest are horizontal rolls of them (distinct addresses and content
    positions, same statistics) to keep host-side generation short."""
    from hyslam_b200 import synth
    base = [synth.stereo_pair(H, W, seed0 + i) for i in range(distinct if first else min(distinct, n_pairs))]
    out = np.empty((2 * n_pairs, H, W), np.uint8)
    for q in range(n_pairs):
        p = first + q              # `first`: pairs [first, first + n_pairs) of the same endless sequence
        L, R = base[p % len(base)]
        s = 37 * (p // len(base))
        out[2 * q] = np.roll(L, s, axis=1)
        out[2 * q + 1] = np.roll(R, s, axis=1)
    return out


def level_bytes(ex):
    """pyramid level pixels of one WxH image as the library lays them out (ORBExtractor.cpp:569)"""
    out = []
    for l in range(ex.GetLevels()):
        lw, lh = ex.level_size(W, H, l)
        out.append(lw * lh)
    return out


def bind_near_gpu(local):
    """Pin this rank's host threads (and therefore the first-touch placement of its pinned staging buffers) to the NUMA node
    of its GPU: with 8 ranks on one box every rank otherwise pulls its uploads from whichever socket the allocator picked."""
    info = {"numa_node": None, "cpus": len(os.sched_getaffinity(0))}
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        bus = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return info
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            info = {"numa_node": node, "cpus": len(cpus)}
    except Exception:
        pass
    return info


def other_configs(hb, F, torch, dev, local, stream, P):
    """Line items for the BASELINE.json configs that are not the headline (rank 0 only, a few steps each):
    C1 / C3 through the host-buffer batch call, and the C2 step on SPARSE-corner frames (G-blocks, about 1 % FAST corners --
    a camera-like density; the headline's G-noise frames have 24 %), device-resident, with its serialised stage times."""
    from hyslam_b200 import synth
    out = {}
    for name, (h, w), nf, nb in (("C1", (480, 752), 1000, 64), ("C3", (2160, 3840), 8000, 8)):
        try:
            imgs = np.stack([synth.noise_image(h, w, i) for i in range(4)] * (nb // 4))
            pin = torch.from_numpy(imgs).pin_memory().numpy()
            ex = hb.ORBExtractor(hb.FeatureExtractorSettings(nFeatures=nf), device=local)
            capn = ex.keypoint_bound(w, h) + 64
            ex.extract_batch(pin, capacity=capn)
            t0 = time.perf_counter(); reps = 5
            for _ in range(reps):
                _, _, counts = ex.extract_batch(pin, capacity=capn)
            dt = (time.perf_counter() - t0) / reps
            one = pin[0]
            ex(one, None, capacity=capn)
            t0 = time.perf_counter()
            for _ in range(20):
                ex(one, None, capacity=capn)
            d1 = (time.perf_counter() - t0) / 20
            out[name] = {"workload": f"{w}x{h} mono, {nf} features, 8 levels, scale 1.2; hyorb_extract_batch_host, {nb} frames per call, copies included",
                         "frames_per_s": nb / dt, "keypoints_per_frame": float(counts.mean()), "single_frame_call_ms": 1e3 * d1}
            ex.close()
        except Exception as e:
            out[name] = {"error": str(e)[:200]}
    try:
        B = 2 * P
        base = [synth.stereo_pair(H, W, 7000 + i, kind="blocks") for i in range(8)]
        WP = (W + 15) & ~15
        t = torch.zeros((B, H, WP), dtype=torch.uint8, device=dev)
        host = np.empty((B, H, W), np.uint8)
        for p_ in range(P):
            L, R = base[p_ % 8]
            host[2 * p_] = np.roll(L, 37 * (p_ // 8), axis=1); host[2 * p_ + 1] = np.roll(R, 37 * (p_ // 8), axis=1)
        t[:, :, :W] = torch.from_numpy(host).to(dev)
        cap = 2560
        ex = hb.ORBExtractor(hb.FeatureExtractorSettings(nFeatures=NFEAT), device=local, stream=stream.cuda_stream)
        sp = ex.stereo_params(hb.StereoCamera(**CAM))
        d_kps = torch.empty((B, cap, 7), dtype=torch.float32, device=dev); d_desc = torch.empty((B, cap, 32), dtype=torch.uint8, device=dev)
        d_counts = torch.zeros(B, dtype=torch.int32, device=dev)
        d_uR = torch.empty((P, cap), dtype=torch.float32, device=dev); d_depth = torch.empty((P, cap), dtype=torch.float32, device=dev)
        step = lambda: ex.process_stereo_batch_device(sp, t.data_ptr(), P, W, H, WP, WP * H, d_kps.data_ptr(), d_desc.data_ptr(), cap,
                                                      d_counts.data_ptr(), d_uR.data_ptr(), d_depth.data_ptr())
        for _ in range(3):
            step()
        ex.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(10):
            step()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        ex.set_pipelining(device_lanes=1, side_blur=0)
        step(); ex.sync()
        ex.set_profiling(True); ex.stage_times(reset=True)
        for _ in range(5):
            step()
        st_ms, calls = ex.stage_times(reset=True)
        out["C2_sparse_corners"] = {"workload": f"the C2 step on G-blocks frames (flat canvas + rectangles + +-2 noise), {P} pairs per step, device-resident",
                                    "frames_per_s": P / (ms * 1e-3), "ms_per_step": ms, "keypoints_per_frame": float(d_counts.sum().item()) / P,
                                    "stage_ms_per_step": {k: v / max(calls, 1) for k, v in st_ms.items()}}
        ex.close()
    except Exception as e:
        out["C2_sparse_corners"] = {"error": str(e)[:200]}
    return out


def output_digest(ex, cam, cap, rank, world, dist, n_pairs=256, chunk=64):
    """SHA-256 over the outputs of a FIXED set of `n_pairs` stereo pairs (the same frames whatever the number of ranks): rank r
    processes its contiguous share through the host-buffer ABI call, every pair is hashed on its own (counts, keypoints,
    descriptors, uR, depth -- only the entries that exist), rank 0 hashes the per-pair digests in frame order.  Equal digests
    at N = 1, 2, 4, 8 mean frame k produced the same bytes on every partition."""
    import hashlib
    lo, hi = n_pairs * rank // world, n_pairs * (rank + 1) // world
    mine = []
    for a in range(lo, hi, chunk):
        b = min(a + chunk, hi)
        imgs = make_pairs(b - a, seed0=5000, distinct=8, first=a)
        k, d, c, uR, dep = ex.process_stereo_batch(imgs, cam, capacity=cap)
        for p in range(b - a):
            nl, nr = int(c[2 * p]), int(c[2 * p + 1])
            hsh = hashlib.sha256()
            hsh.update(np.int32([nl, nr]).tobytes())
            hsh.update(np.ascontiguousarray(k[2 * p][:nl]).tobytes()); hsh.update(np.ascontiguousarray(k[2 * p + 1][:nr]).tobytes())
            hsh.update(np.ascontiguousarray(d[2 * p][:nl]).tobytes()); hsh.update(np.ascontiguousarray(d[2 * p + 1][:nr]).tobytes())
            hsh.update(np.ascontiguousarray(uR[p][:nl]).tobytes()); hsh.update(np.ascontiguousarray(dep[p][:nl]).tobytes())
            mine.append(hsh.digest())
    parts = [mine]
    if dist is not None:
        parts = [None] * world
        dist.all_gather_object(parts, mine)
    if rank != 0:
        return None
    top = hashlib.sha256()
    for part in parts:
        for dg in part:
            top.update(dg)
    return {"sha256": top.hexdigest(), "pairs": n_pairs, "frames": "make_pairs(256, seed0=5000): identical frames for every N, sharded contiguously by rank"}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ts, l in self.lines:
            if t0 is not None and not (t0 <= ts <= t1 + 0.2):
                continue
            f = [x.strip() for x in l.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples in the timed region"], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def cpu_port_run(images, n_threads):
    """reference-shaped CPU run of the same step on host cores: the oracle's ORBExtractor restatement over all images
    (pthread pool, one extractor per thread as ImageProcessing.cpp:82-84 does for L/R) + its Stereomatcher restatement."""
    from oracle import oracle as O
    p = O.default_params(NFEAT)
    kps, desc, counts = O.extract_batch(images, p, nthreads=n_threads)
    sp = O.StereoParams(CAM["mbf"], CAM["fx"], int(CAM["mnMaxY"]), 100.0, 50.0, 31.0)
    nk = 0
    for i in range(0, len(images), 2):
        nl, nr = counts[i], counts[i + 1]
        O.stereo_match(sp, kps[i, :nl], desc[i, :nl], kps[i + 1, :nr], desc[i + 1, :nr])
        nk += int(nl + nr)
    return nk


def ref_available():
    try:
        from oracle import ref as R
        if not R.available():
            return False
        R.lib()
        return True
    except Exception:
        return False


def cpu_ref_run(images, n_workers, threads_per_pair=1):
    """The reference's own code (oracle/_ref) over the pairs of `images` ([2n, H, W], L/R interleaved): `n_workers` host threads,
    each driving whole pairs through ORBExtractor x 2 + Stereomatcher (ref_process_stereo_pairs releases the GIL).
    threads_per_pair = 2 reproduces ImageProcessing.cpp:82-84 (left extractor on a transient std::thread)."""
    from oracle import oracle as O
    from oracle import ref as R
    p = O.default_params(NFEAT)
    sp = O.StereoParams(CAM["mbf"], CAM["fx"], int(CAM["mnMaxY"]), 100.0, 50.0, 31.0)
    L, Rt = np.ascontiguousarray(images[0::2]), np.ascontiguousarray(images[1::2])
    n = len(L)
    n_workers = max(1, min(n_workers, n))
    bounds = [n * i // n_workers for i in range(n_workers + 1)]
    out = [0] * n_workers

    def work(i):
        a, b = bounds[i], bounds[i + 1]
        if b > a:
            out[i] = R.process_stereo_pairs(p, sp, L[a:b], Rt[a:b], threads_per_pair)[0]
    ths = [threading.Thread(target=work, args=(i,)) for i in range(n_workers)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    return sum(out)


def cv2_primitives_ms(image):
    """cross-check of the baseline's pixel kernels: the same resize chain + per-level FAST + GaussianBlur through cv2's own
    (IPP / AVX) primitives, one thread, versus oracle/cvshim's -- so that the reader can see the shim is not a slow stand-in"""
    try:
        import cv2
        from oracle import oracle as O
        from oracle import ref as R
    except Exception:
        return None
    cv2.setNumThreads(1)
    p = O.default_params(NFEAT)
    sizes = O.level_sizes(p, image.shape[1], image.shape[0])[1:]
    fd = cv2.FastFeatureDetector_create(20, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)

    def run_cv2():
        cur, lv = image, [image]
        for (w, h) in sizes:
            cur = cv2.resize(cur, (w, h), interpolation=cv2.INTER_LINEAR)
            lv.append(cur)
        for l in lv:
            fd.detect(l)
            cv2.GaussianBlur(l, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)

    def run_shim():
        cur, lv = image, [image]
        for (w, h) in sizes:
            cur = R.shim_resize(cur, w, h)
            lv.append(cur)
        for l in lv:
            R.shim_fast(l)
            R.shim_blur(l)
    res = {}
    for name, fn in (("cv2", run_cv2), ("cvshim", run_shim)):
        fn()
        t0 = time.perf_counter()
        for _ in range(3):
            fn()
        res[name + "_ms_per_image"] = (time.perf_counter() - t0) / 3 * 1e3
    res["what"] = "resize chain + FAST-9/16 NMS + GaussianBlur 7x7 over the 8 levels of one frame, 1 thread"
    return res


def c4_cpu_arms(c4_data, cores):
    """C4 on the host cores, bounded samples of the same 8000 x 8000 problem, all threads (ctypes releases the GIL): the reference's own
    distance path (FeatureDescriptor::distance -> ORBDistance::distance, one cv::Mat clone per call: FeatureDescriptor.h:31) and a lean
    scan (contiguous descriptors, hardware popcount); both with the best / second-best bookkeeping, checked against the GPU's answer."""
    from oracle import ref as R
    da, db, gpu_idx, gpu_best = c4_data
    out = {}
    for name, lean, per_thread, reps in (("reference_distance_path", False, 400, 1), ("lean_popcount", True, 500, 20)):
        nqs = min(len(da), cores * per_thread)
        bounds = [nqs * i // cores for i in range(cores + 1)]
        res = [None] * cores

        def work(i):
            a, b = bounds[i], bounds[i + 1]
            for _ in range(reps):
                if b > a:
                    res[i] = R.bf_scan(da[a:b], db, lean)
        ths = [threading.Thread(target=work, args=(i,)) for i in range(cores)]
        t0 = time.perf_counter()
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        dt = time.perf_counter() - t0
        idx = np.concatenate([r[0] for r in res if r is not None]); best = np.concatenate([r[1] for r in res if r is not None])
        out[name] = {"distance_evals_per_s": reps * nqs * len(db) / dt, "cores": cores,
                     "sample": f"{nqs} of the 8000 queries x 8000 targets" + (f", {reps} repetitions" if reps > 1 else "") + f" ({dt:.2f} s)",
                     "agrees_with_gpu": bool(np.array_equal(idx, gpu_idx[:nqs]) and np.array_equal(best, gpu_best[:nqs].astype(np.int32)))}
    return out


def cpu_baseline_block(images, cores, seconds):
    """cpu_baseline object: all-cores throughput of the reference's code (or the port when _ref is absent) on `images`,
    repeated for about `seconds`; plus the labelled side figures."""
    n_pairs = len(images) // 2
    use_ref = ref_available()
    run = (lambda: cpu_ref_run(images, cores, 1)) if use_ref else (lambda: cpu_port_run(images, cores))
    run()
    t0 = time.perf_counter()
    reps = 0
    while True:
        run()
        reps += 1
        if time.perf_counter() - t0 > seconds or reps >= 200:
            break
    dt = time.perf_counter() - t0
    cpu = {"value": n_pairs * reps / dt, "unit": UNIT, "cores": cores, "kind": "reference" if use_ref else "port",
           "sample": (f"{n_pairs} pairs x {reps} repetitions of the same workload ({dt:.1f} s); "
                      + ("hySLAM's own ORBExtractor / ORBFinder / Stereomatcher sources compiled unmodified (oracle/_ref) over oracle/cvshim's "
                         f"SIMD OpenCV primitives, {cores} host threads, one pair stream per thread" if use_ref else
                         f"scalar C port of the reference (oracle/orb_oracle.c) on all {cores} host threads"))}
    if use_ref:
        # reference-shaped threading: ONE pair stream, left extractor on a transient thread (ImageProcessing.cpp:82-84)
        k = min(n_pairs, 4)
        t0 = time.perf_counter()
        cpu_ref_run(images[: 2 * k], 1, 2)
        cpu["reference_shaped"] = {"value": k / (time.perf_counter() - t0), "unit": UNIT, "cores": 2,
                                   "sample": f"{k} pairs, one pair at a time, 2 threads per pair as ImageProcessing.cpp:82-84 runs them"}
        k = min(n_pairs, max(1, cores // 2))
        t0 = time.perf_counter()
        cpu_port_run(images[: 2 * k], cores)
        cpu["scalar_port"] = {"value": k / (time.perf_counter() - t0), "unit": UNIT, "cores": cores,
                              "sample": f"{k} pairs, scalar C restatement (oracle/orb_oracle.c: the parity checker, not tuned), all host threads"}
        cpu["pixel_primitives"] = cv2_primitives_ms(images[0])
    return cpu


def run_reference(args):
    """--impl reference: the reference's own CPU code on host cores (oracle/_ref; oracle port if it did not travel)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle as O
    O.build()
    cores = len(os.sched_getaffinity(0))
    use_ref = ref_available()
    # bounded sample: one pair per host thread per step with the reference's code (about 70 ms per pair and thread)
    sample_pairs = max(1, min(args.pairs, cores if use_ref else max(1, cores // 2)))
    images = make_pairs(sample_pairs)
    run = (lambda: cpu_ref_run(images, cores, 1)) if use_ref else (lambda: cpu_port_run(images, cores))
    for _ in range(args.warmup):
        run()
    t0 = time.perf_counter()
    nk = 0
    for _ in range(args.steps):
        nk += run()
    dt = time.perf_counter() - t0
    value = sample_pairs * args.steps / dt
    kind = "reference" if use_ref else "port"
    what = ("hySLAM's own ORBExtractor / ORBFinder / Stereomatcher sources compiled unmodified (oracle/_ref) over oracle/cvshim's SIMD OpenCV "
            "primitives" if use_ref else "scalar C port of the reference (oracle/orb_oracle.c)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic", "config": {"workload": workload_name(args.pairs), "sample": f"{sample_pairs} pairs per step"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{sample_pairs} pairs x {args.steps} steps of the same workload; {what}; all {cores} host threads, one pair stream per thread"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "kpts_per_sec": nk / dt,
    }
    if use_ref:
        k = min(sample_pairs, 4)
        t0 = time.perf_counter()
        cpu_ref_run(images[: 2 * k], 1, 2)
        line["cpu_baseline"]["reference_shaped"] = {"value": k / (time.perf_counter() - t0), "unit": UNIT, "cores": 2,
                                                    "sample": f"{k} pairs, one at a time, 2 threads per pair (ImageProcessing.cpp:82-84)"}
        line["cpu_baseline"]["pixel_primitives"] = cv2_primitives_ms(images[0])
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=128, help="stereo pairs per step (per GPU)")
    ap.add_argument("--rotate", type=int, default=4, help="distinct device-resident batches cycled so inputs exceed L2")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-workers", type=int, default=2, help="host threads (one extractor handle each) of the end-to-end leg")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import hyslam_b200 as hb
    from hyslam_b200 import _ffi as F

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (libhyorb has no CPU fallback)")
    torch.cuda.set_device(local)
    full_affinity = os.sched_getaffinity(0)
    affinity = bind_near_gpu(local)          # before any pinned allocation
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    P = args.pairs
    B = 2 * P
    cap = 2560                       # >= keypoints per image (about 2010 for 2000 requested features)
    stream = torch.cuda.Stream(device=dev)     # an explicit stream: the handle enqueues on it and the events are recorded on it
    torch.cuda.set_stream(stream)
    ex = hb.ORBExtractor(hb.FeatureExtractorSettings(nFeatures=NFEAT), device=local, stream=stream.cuda_stream)
    cam = hb.StereoCamera(**CAM)
    sp = ex.stereo_params(cam)

    # ---- inputs: `rotate` distinct batches resident in HBM (rotate * B * W*H bytes > 126 MB L2), rank-specific seeds
    host_batches = [make_pairs(P, seed0=1000 + 100 * rank + 10 * r) for r in range(min(args.rotate, 3))]
    pinned = [torch.from_numpy(hb_).pin_memory() for hb_ in host_batches]
    # device-resident frames are pitched 2-D images (row pitch rounded up to 16 bytes, as cudaMallocPitch / GpuMat give):
    # the library then reads them in place through TMA; a dense odd-pitch batch would first be repacked on the device
    WP = (W + 15) & ~15
    dev_batches = []
    for r in range(args.rotate):
        t = pinned[r % len(pinned)].to(dev, non_blocking=True)
        if r >= len(pinned):
            t = torch.roll(t, shifts=11 * r, dims=2)
        tp = torch.zeros((B, H, WP), dtype=torch.uint8, device=dev)
        tp[:, :, :W] = t
        dev_batches.append(tp)
    in_bytes = B * W * H
    d_kps = torch.empty((B, cap, 7), dtype=torch.float32, device=dev)
    d_desc = torch.empty((B, cap, 32), dtype=torch.uint8, device=dev)
    d_counts = torch.zeros(B, dtype=torch.int32, device=dev)
    d_uR = torch.empty((P, cap), dtype=torch.float32, device=dev)
    d_depth = torch.empty((P, cap), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()

    def step_device(i):
        t = dev_batches[i % len(dev_batches)]
        ex.process_stereo_batch_device(sp, t.data_ptr(), P, W, H, WP, WP * H, d_kps.data_ptr(), d_desc.data_ptr(), cap,
                                       d_counts.data_ptr(), d_uR.data_ptr(), d_depth.data_ptr())

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput (value): the library's default pipelining (sub-batch lanes on separate streams)
    for i in range(args.warmup):
        step_device(i)
    ex.sync()
    launches0 = ex.launch_count()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    e0.record(stream)
    for i in range(args.steps):
        step_device(args.warmup + i)
    e1.record(stream)
    barrier()
    t_wall1 = time.perf_counter()
    ex.sync()                                         # surfaces device-side status (capacity etc.)
    clocks = sampler.stop(t_wall0, t_wall1)
    ms = e0.elapsed_time(e1)
    launches = ex.launch_count() - launches0
    counts = d_counts.cpu().numpy()
    matched = int(((d_uR.cpu().numpy() >= 0) & (np.arange(cap)[None, :] < counts[0::2][:, None])).sum())
    kp_per_step = int(counts.sum())
    tmax = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_max = float(tmax.item())
    value = world * P * args.steps / (ms_max / 1e3)

    # ---- the same measurement over a timed region of at least ~0.6 s (K steps of 3 ms are a 60 ms window)
    sus_steps = max(args.steps, int(0.6 / max(ms_max / args.steps / 1e3, 1e-6)) + 1)
    barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record(stream)
    for i in range(sus_steps):
        step_device(i)
    s1.record(stream)
    barrier()
    ex.sync()
    tsus = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tsus, op=dist.ReduceOp.MAX)
    sustained = {"steps": sus_steps, "seconds": float(tsus.item()) / 1e3, "value": world * P * sus_steps / (float(tsus.item()) / 1e3), "unit": UNIT}

    # ---- per-kernel durations for the roofline: the same steps once more with the kernels serialised (one lane, no side
    # stream), CUDA events on the launching stream around every stage; under the default pipelining kernels of different
    # lanes overlap and an event pair would also time the neighbours
    ex.set_pipelining(device_lanes=1, side_blur=0)
    step_device(0)
    ex.sync()
    ex.set_profiling(True)
    ex.stage_times(reset=True)
    ksteps = max(3, min(args.steps, 10))
    for i in range(ksteps):
        step_device(args.warmup + i)
    stage_ms, stage_calls = ex.stage_times(reset=True)
    ex.set_profiling(False)
    ex.set_pipelining(device_lanes=int(os.environ.get("HYORB_LANES", "3")), side_blur=int(os.environ.get("HYORB_SIDE_BLUR", "2")),
                      host_lanes=int(os.environ.get("HYORB_HOST_LANES", "-1")))

    # ---- end to end through the host-buffer ABI call (pinned host in, host out).  Each call is synchronous (uploads, kernels
    # and downloads of one batch, pipelined over lanes inside the call).  Two numbers: one handle called back to back, and the
    # way a throughput user drives it -- two host threads, one extractor handle (own stream + workspace) each, like the
    # reference runs its left and right extractor objects on two threads (ImageProcessing.cpp:82-84) -- so that one
    # call's PCIe transfers overlap the other's kernels.  ctypes releases the GIL for the duration of a call.
    def make_outs():
        o = (np.empty((B, cap), F.KP_DTYPE), np.empty((B, cap, 32), np.uint8), np.zeros(B, np.int32),
             np.empty((P, cap), np.float32), np.empty((P, cap), np.float32))
        pin = [torch.from_numpy(x.view(np.uint8).reshape(-1)).pin_memory() for x in o]
        return tuple(po.numpy().view(x.dtype).reshape(x.shape) for po, x in zip(pin, o)), pin
    outs, _keep0 = make_outs()
    h_in = pinned[0].numpy()
    e2e_steps = max(4, min(args.steps, 20))
    for _ in range(2):
        ex.process_stereo_batch(h_in, cam, capacity=cap, out=outs)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        ex.process_stereo_batch(pinned[i % len(pinned)].numpy(), cam, capacity=cap, out=outs)
    torch.cuda.synchronize()
    dt1 = time.perf_counter() - t0

    n_workers = max(1, args.e2e_workers)
    # several handles overlap each other's copies and kernels already: each then runs best with fewer, larger lanes (measured 2 handles
    # x 4 / 6 / 8 / 12 lanes: 46.7k / 46.3k / 45.8k / 45.3k pairs/s; one handle alone: 33.2k / 34.7k / 35.8k / 35.8k)
    mt_lanes = int(os.environ.get("HYORB_HOST_LANES", "4" if n_workers > 1 else "-1"))
    ex.set_pipelining(host_lanes=mt_lanes)
    workers = [(ex, outs)]
    keep = []
    for _ in range(n_workers - 1):
        o2, k2 = make_outs()
        keep.append(k2)
        workers.append((hb.ORBExtractor(hb.FeatureExtractorSettings(nFeatures=NFEAT), device=local), o2))
        workers[-1][0].set_pipelining(host_lanes=mt_lanes)
    for exw, ow in workers[1:]:
        exw.process_stereo_batch(h_in, cam, capacity=cap, out=ow)          # allocate its workspace outside the timed region
    errs = []

    def run_threads(batches):
        """every worker thread calls its handle e2e_steps times on `batches`; wall time from the common start to the last join"""
        start = threading.Barrier(n_workers + 1)

        def work(wi):
            exw, ow = workers[wi]
            try:
                torch.cuda.set_device(local)
                start.wait()
                for i in range(wi, n_workers * e2e_steps, n_workers):
                    exw.process_stereo_batch(batches[i % len(batches)], cam, capacity=cap, out=ow)
            except Exception as e:                      # surfaced after the join
                errs.append(e)
        ths = [threading.Thread(target=work, args=(wi,)) for wi in range(n_workers)]
        for t in ths:
            t.start()
        barrier()
        start.wait()
        t0 = time.perf_counter()
        for t in ths:
            t.join()
        torch.cuda.synchronize()
        return time.perf_counter() - t0
    dt2 = run_threads([pb.numpy() for pb in pinned])
    # the same frames in host buffers whose row pitch is a multiple of 16 bytes (1248): uploaded flat like the dense ones, but read in
    # place through TMA -- the device-side repack (k_repack) of dense 1241-byte rows drops out.  Reported next to the headline, which
    # stays on dense rows (what a cv::Mat of this width is).
    WPh = (W + 15) & ~15
    pitched_host = []
    for pb in pinned:
        ph = torch.zeros((B, H, WPh), dtype=torch.uint8).pin_memory()
        ph.numpy()[:, :, :W] = pb.numpy()
        pitched_host.append(ph)
    for exw, ow in workers:
        exw.process_stereo_batch(pitched_host[0].numpy()[:, :, :W], cam, capacity=cap, out=ow)
    dt3 = run_threads([ph.numpy()[:, :, :W] for ph in pitched_host])
    if errs:
        raise errs[0]
    for exw, _ in workers[1:]:
        exw.close()
    tm = torch.tensor([dt1, dt2, dt3], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    e2e_single = world * P * e2e_steps / float(tm[0].item())
    e2e_value = world * P * n_workers * e2e_steps / float(tm[1].item())
    e2e_pitched = world * P * n_workers * e2e_steps / float(tm[2].item())
    # bytes the call moves per step: the images up; per image min(capacity, the extractor's keypoint bound) entries down (2-D copies)
    rows = min(cap, ex.keypoint_bound(W, H))
    d2h = B * 4 + B * rows * (F.KP_DTYPE.itemsize + 32) + 2 * P * rows * 4
    # what each rank's PCIe link carried during its own end-to-end leg
    link = torch.tensor([in_bytes * n_workers * e2e_steps / dt2 / 1e9, d2h * n_workers * e2e_steps / dt2 / 1e9], dtype=torch.float64, device=dev)
    links = [link]
    if dist is not None:
        links = [torch.zeros_like(link) for _ in range(world)]
        dist.all_gather(links, link)
    per_rank_gbs = [{"h2d": round(float(l[0]), 2), "d2h": round(float(l[1]), 2)} for l in links]
    digest = output_digest(ex, cam, cap, rank, world, dist)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    # ---- auxiliary line item (config C4, not the headline metric): keyframe-pair brute-force Hamming matching, 8000 x 8000
    # descriptors, BoW acceptance rule (SearchForTriangulation's scan), device-resident
    c4 = None
    c4_data = None
    try:
        from hyslam_b200 import synth
        nq = nt = 8000
        da = synth.random_descriptors(nq, 7)
        db = np.roll(da, 17, axis=0) ^ synth.random_descriptors(nt, 8) & 0x11
        tq, tt = torch.from_numpy(da).to(dev), torch.from_numpy(db).to(dev)
        obi = torch.empty(nq, dtype=torch.int32, device=dev); ob = torch.empty(nq, dtype=torch.int16, device=dev)
        osec = torch.empty(nq, dtype=torch.int16, device=dev); oacc = torch.empty(nq, dtype=torch.uint8, device=dev)
        fm = hb.FeatureMatcher(device=local, stream=stream.cuda_stream)
        run = lambda: fm.match_bruteforce_device(tq.data_ptr(), nq, tt.data_ptr(), nt, F.RULE_BOW, 50.0, 0.6, obi.data_ptr(), ob.data_ptr(),
                                                 osec.data_ptr(), oacc.data_ptr())
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        m0.record(stream)
        for _ in range(reps):
            run()
        m1.record(stream)
        torch.cuda.synchronize()
        mms = m0.elapsed_time(m1) / reps
        c4 = {"workload": "C4: 8000 x 8000 descriptors, brute force, best/second-best + ratio test", "ms_per_pair_of_keyframes": mms,
              "distance_evals_per_s": nq * nt / (mms * 1e-3), "popc32_per_s": 8 * nq * nt / (mms * 1e-3), "accepted": int(oacc.sum().item())}
        c4_data = (da, db, obi.cpu().numpy(), ob.cpu().numpy())
        try:      # the POPC32 issue rate measured on a B200 of this pool by tools/popc_bench.cu (profiles/r2_popc_peak.json)
            pk = json.load(open(os.path.join(ROOT, "profiles", "r2_popc_peak.json")))
            c4["popc32_peak_measured_per_s"] = pk["popc_plus_iadd"]["popc_per_s"]
            c4["frac_of_popc_peak_plain_form"] = c4["popc32_per_s"] / pk["popc_plus_iadd"]["popc_per_s"]
            c4["note"] = ("popc32_per_s counts the plain form's 8 POPC32 per 256-bit distance; the kernel issues 4 (carry-save compression of the 8 XOR words, "
                          "hamming256() in common.cuh), so the plain-form fraction may exceed what the POPC pipe alone could do")
        except Exception:
            pass
    except Exception as e:                      # never let the auxiliary item break the headline line
        c4 = {"error": str(e)[:200]}

    others = other_configs(hb, F, torch, dev, local, stream, P)

    # ---- roofline of the dominant kernel
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    lv = level_bytes(ex)
    sumP = sum(lv)
    fused = os.environ.get("HYORB_FUSED_LEVELS", "1") != "0"
    alg = {   # algorithmic bytes per image
        # level.cu (default): every level read ONCE, blurred copy and next level written; pyramid.cu + blur.cu: each level read twice
        "pyramid": (2 * sumP + sum(lv[1:])) if fused else (sum(lv[:-1]) + sum(lv[1:])),
        "fast": sumP,
        "blur": 0 if fused else 2 * sumP,                # materialised blur: read level + write blurred level
        "describe": sumP + 60 * (kp_per_step / B),       # read blurred windows (<= one pass) + 28 B keypoint + 32 B descriptor
        "quadtree": 0, "stereo": 0,
    }
    stages = {}
    if fused:
        stage_ms = {k: v for k, v in stage_ms.items() if k != "blur"}       # the blur is part of the "pyramid" stage (k_level)
    for k, v in stage_ms.items():
        per_launch_ms = v / max(stage_calls, 1)
        stages[k] = {"ms_per_step": per_launch_ms, "share": v / max(sum(stage_ms.values()), 1e-9),
                     "algorithmic_GBps": (alg[k] * B / (per_launch_ms * 1e-3) / 1e9) if alg.get(k) and per_launch_ms > 0 else None}
    dom = max(stage_ms, key=lambda k: stage_ms[k])
    dom_ms = stage_ms[dom] / max(stage_calls, 1)
    achieved = alg[dom] * B / (dom_ms * 1e-3) / 1e9 if alg.get(dom) else 0.0
    kname = {"fast": "k_fast", "pyramid": "k_level" if fused else "k_resize", "blur": "k_blur", "describe": "k_describe", "quadtree": "k_quadtree", "stereo": "k_stereo"}[dom]
    # DRAM traffic of that kernel from the committed `ncu --set full` capture of this same command (profiles/traffic.json:
    # bytes per image, dram__bytes_read.sum + dram__bytes_write.sum), scaled to the images of one launch
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if kname in tj:
            traffic = float(tj[kname]["dram_bytes_per_image"]) * B
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": kname,
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg[dom] * B, "launch_ms": dom_ms,
                "timing": f"CUDA events on the launching stream, {ksteps} steps with the kernels serialised (1 lane, no side stream)",
                "note": "integer-logic kernel: ncu shows its shared-memory pipe 73 % and ALU pipe 57 % busy at 59 % issue, DRAM 2 % (profiles/r2i_k_fast_full.txt); reported against the HBM roofline as the contract asks; DESIGN.md section 4"}

    # ---- CPU baseline: the reference's own code on all host threads, bounded sample (kind "reference"; "port" without _ref)
    cpu = None
    if not args.no_cpu and world == 1:        # the CPU arm is timed at N = 1 only (rank 0); multi-GPU lines carry cpu_baseline = null
        os.sched_setaffinity(0, full_affinity)        # the GPU legs ran on the GPU's NUMA node; the CPU arm gets every host thread
        cores = len(os.sched_getaffinity(0))
        sample_pairs = max(1, min(P, cores))
        cpu = cpu_baseline_block(host_batches[0][: 2 * sample_pairs], cores, args.cpu_seconds)
        if c4_data is not None and isinstance(c4, dict) and "error" not in c4 and ref_available():
            c4["cpu_baseline"] = c4_cpu_arms(c4_data, cores)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic",
        "config": {"workload": workload_name(P), "frame": "one stereo pair = 2 images (images/s = 2 x value)", "pairs_per_step_per_gpu": P, "partition": "frames sharded by rank, no collective on the data path",
                   "l2": f"{args.rotate} distinct device-resident input batches cycled ({args.rotate * in_bytes / 1e6:.0f} MB of inputs > 126 MB L2); "
                         f"per-step intermediates ({B} pyramids + blurred copies) also exceed L2"},
        "kpts_per_sec": value * kp_per_step / P, "keypoints_per_frame": kp_per_step / P, "stereo_matches_per_frame": matched / P,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": d2h, "steps": n_workers * e2e_steps,
                "api": f"hyorb_process_stereo_batch_host (pinned host buffers in and out), {n_workers} host threads with one extractor handle each"
                       + (f", {mt_lanes} lanes per call (hyorb_extractor_set_pipelining)" if mt_lanes > 0 else ""),
                "single_handle_value": e2e_single,
                "pitched_host_rows_value": e2e_pitched, "pitched_host_rows": f"same call, host frames with a {(W + 15) & ~15}-byte row pitch (read in place through TMA, no device-side repack)",
                "per_rank_link_GBps": per_rank_gbs, "host_affinity": affinity},
        "sustained": sustained, "digest": digest,
        "gpu_launches": int(launches), "roofline": roofline, "stages": stages, "c4_match": c4, "other_configs": others, "cpu_baseline": cpu, "clocks": clocks,
    }
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
