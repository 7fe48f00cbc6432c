"""Frame sharding for offline reprocessing runs (BASELINE config C5): frames are independent, so rank r of G processes
takes the contiguous block [r*F/G, (r+1)*F/G) and no collective touches the data path.  The only exchange is the
final variable-length gather of results (counts -> offsets -> payload), done with torch.distributed (NCCL on GPUs,
gloo in the CPU tests).  Output for a frame is identical for any G (tests/test_shard_gloo.py)."""
import numpy as np


def frame_range(n_frames, rank, world):
    """Contiguous block of frame indices owned by `rank`."""
    lo = (n_frames * rank) // world
    hi = (n_frames * (rank + 1)) // world
    return lo, hi


def pack_results(counts, kps, desc, extra=None):
    """Flatten per-frame variable-length results ([B,cap] arrays + counts) into contiguous payloads."""
    counts = np.asarray(counts, np.int64)
    ks = [kps[i, :c] for i, c in enumerate(counts)]
    ds = [desc[i, :c] for i, c in enumerate(counts)]
    out = {"counts": counts, "kps": np.concatenate(ks) if ks else kps[:0, 0], "desc": np.concatenate(ds) if ds else desc[:0, 0]}
    if extra:
        for name, arr in extra.items():
            out[name] = np.concatenate([arr[i, :c] for i, c in enumerate(counts)]) if len(counts) else arr[:0, 0]
    return out


def gather_results(local, dist=None, dst=0):
    """Variable-length gather of pack_results() dicts to rank `dst`, concatenated in rank (== frame) order.
    Uses tensor collectives only: all_gather of the sizes, then a padded gather of each payload as bytes."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    merged = {}
    for name in sorted(local):
        a = np.ascontiguousarray(local[name])
        raw = torch.from_numpy(a.view(np.uint8).reshape(-1).copy()).to(dev)
        n = torch.tensor([raw.numel()], dtype=torch.int64, device=dev)
        sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(sizes, n)
        sizes = [int(s.item()) for s in sizes]
        mx = max(max(sizes), 1)
        pad = torch.zeros(mx, dtype=torch.uint8, device=dev)
        pad[: raw.numel()] = raw
        bufs = [torch.zeros(mx, dtype=torch.uint8, device=dev) for _ in range(world)]
        dist.all_gather(bufs, pad)
        if rank == dst:
            parts = [bufs[r][: sizes[r]].cpu().numpy() for r in range(world)]
            flat = np.concatenate(parts) if parts else np.zeros(0, np.uint8)
            item = a.dtype.itemsize * int(np.prod(a.shape[1:], dtype=np.int64)) if a.ndim > 1 else a.dtype.itemsize
            merged[name] = flat.view(a.dtype).reshape((-1,) + a.shape[1:]) if item else flat
    return merged if rank == dst else None
