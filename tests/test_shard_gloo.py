"""world_size-2 gloo test of the frame sharding + variable-length result gather (the N>1 host logic), on CPU.
The per-frame results come from the CPU oracle here: what is under test is the partition / gather plumbing, which is
identical on GPUs (there the extractor fills the same arrays)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _frames(n):
    from hyslam_b200 import synth
    return np.stack([synth.noise_image(120, 160, 100 + i) if i % 3 else synth.blocks_image(120, 160, 100 + i) for i in range(n)])


def _extract_block(frames):
    from oracle import oracle as O
    p = O.default_params(300, 1.2, 4, 30)
    return O.extract_batch(frames, p, cap=1200)


def _worker(rank, world, port, n_frames, out):
    import torch.distributed as dist
    from hyslam_b200 import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard.frame_range(n_frames, rank, world)
    kps, desc, counts = _extract_block(_frames(n_frames)[lo:hi])
    local = shard.pack_results(counts, kps, desc)
    merged = shard.gather_results(local, dist)
    if rank == 0:
        np.savez(out, **merged)
    dist.barrier()
    dist.destroy_process_group()


def test_frame_range_partitions_exactly():
    from hyslam_b200 import shard
    for n in (0, 1, 7, 10000):
        for g in (1, 2, 4, 8):
            blocks = [shard.frame_range(n, r, g) for r in range(g)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(g - 1))
            assert max(b[1] - b[0] for b in blocks) - min(b[1] - b[0] for b in blocks) <= 1


def test_two_rank_gloo_run_equals_single_process(tmp_path):
    import torch.multiprocessing as mp
    from hyslam_b200 import shard
    n_frames = 5
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / "merged.npz")
    mp.spawn(_worker, args=(2, port, n_frames, out), nprocs=2, join=True)
    got = np.load(out)
    kps, desc, counts = _extract_block(_frames(n_frames))
    want = shard.pack_results(counts, kps, desc)
    assert np.array_equal(got["counts"], want["counts"])
    assert np.array_equal(got["kps"], want["kps"]) and np.array_equal(got["desc"], want["desc"])
    assert want["counts"].sum() > 500
