#!/usr/bin/env python
"""BASELINE config C5: an offline synthetic stereo sequence reprocessed across G GPUs, frames sharded by rank, no collective
on the data path; the only exchange is the final gather of per-frame digests (and, optionally, of the results).

  python tools/run_sequence.py --frames 10000                       (1 GPU)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/run_sequence.py --frames 10000

The sequence cycles a pool of --pool distinct seeded pairs (frame k = pool[k % pool] rolled by 7*(k // pool) columns), so
every frame's content is a pure function of its index; rank 0 prints one JSON line with the whole-sequence digest, which
must not depend on G (frame k's keypoints / descriptors / uR / depth are bit-identical for any sharding)."""
import argparse, hashlib, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

W, H, NFEAT = 1241, 376, 2000


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=10000)
    ap.add_argument("--pool", type=int, default=16)
    ap.add_argument("--batch", type=int, default=128, help="pairs per call")
    args = ap.parse_args()
    import torch
    import hyslam_b200 as hb
    from hyslam_b200 import shard, synth
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pool = [synth.stereo_pair(H, W, 5000 + i) for i in range(args.pool)]
    lo, hi = shard.frame_range(args.frames, rank, world)
    ex = hb.ORBExtractor(hb.FeatureExtractorSettings(nFeatures=NFEAT), device=local)
    cam = hb.StereoCamera(mbf=386.1448, fx=718.856, mnMaxY=float(H))
    cap = 2560
    buf = torch.empty((2 * args.batch, H, W), dtype=torch.uint8).pin_memory().numpy()
    digests = np.zeros((hi - lo, 32), np.uint8)
    nk = 0
    ex.process_stereo_batch(np.stack([pool[0][0], pool[0][1]] * args.batch), cam, capacity=cap)      # allocate the workspace outside the timed region
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for b0 in range(lo, hi, args.batch):
        n = min(args.batch, hi - b0)
        for j in range(n):
            k = b0 + j
            L, R = pool[k % args.pool]
            s = 7 * (k // args.pool)
            buf[2 * j] = np.roll(L, s, axis=1); buf[2 * j + 1] = np.roll(R, s, axis=1)
        kps, desc, counts, uR, depth = ex.process_stereo_batch(buf[: 2 * n], cam, capacity=cap)
        for j in range(n):
            nl, nr = counts[2 * j], counts[2 * j + 1]
            h = hashlib.sha256()
            for a in (kps[2 * j, :nl], desc[2 * j, :nl], kps[2 * j + 1, :nr], desc[2 * j + 1, :nr], uR[j, :nl], depth[j, :nl]):
                h.update(np.ascontiguousarray(a).tobytes())
            digests[b0 - lo + j] = np.frombuffer(h.digest(), np.uint8)
            nk += int(nl + nr)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    merged = shard.gather_results({"digest": digests, "stats": np.array([[nk, dt]], np.float64)}, dist)
    if rank == 0:
        total = hashlib.sha256(np.ascontiguousarray(merged["digest"]).tobytes()).hexdigest()
        st = merged["stats"]
        print(json.dumps({"config": "C5", "frames": args.frames, "n_gpus": world, "sequence_sha256": total, "keypoints": int(st[:, 0].sum()),
                          "wall_s_max_over_ranks": float(st[:, 1].max()), "frames_per_s_incl_host_frame_synthesis_and_hashing": args.frames / float(st[:, 1].max())}))
    if dist is not None:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
