#!/usr/bin/env python
"""Does write-combined pinned host memory (cudaHostAllocWriteCombined: no CPU-cache snooping on the device's reads) raise the upload
rate of the bench's image batches, alone and with all ranks copying at once?   python tools/pcie_wc.py   or under torchrun.
Rank 0 prints one JSON line: per-GPU H2D GB/s from default pinned memory and from write-combined pinned memory (119.5 MB copies, every
rank inside the same barrier-bracketed window, D2H of 33 MB running at the same time as in the bench's end-to-end leg)."""
import ctypes, json, os, time
import numpy as np
import torch
from cuda.bindings import runtime as cudart

UP, DOWN, REPS = 119453696, 33162240, 20


def host_alloc(nbytes, flags):
    err, ptr = cudart.cudaHostAlloc(nbytes, flags)
    assert int(err) == 0, err
    buf = (ctypes.c_uint8 * nbytes).from_address(int(ptr))
    return np.frombuffer(buf, dtype=np.uint8), int(ptr)


def main():
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    d = torch.empty(UP, dtype=torch.uint8, device="cuda")
    y = torch.empty(DOWN, dtype=torch.uint8, device="cuda"); hy = torch.empty(DOWN, dtype=torch.uint8).pin_memory()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    res = []
    for flags in (cudart.cudaHostAllocDefault, cudart.cudaHostAllocWriteCombined):
        arr, ptr = host_alloc(UP, flags)
        arr[:] = 7                      # touch
        def copy():
            err, = cudart.cudaMemcpyAsync(d.data_ptr(), ptr, UP, cudart.cudaMemcpyKind.cudaMemcpyHostToDevice, s1.cuda_stream)
            assert int(err) == 0, err
            with torch.cuda.stream(s2): hy.copy_(y, non_blocking=True)
        for _ in range(3): copy()
        torch.cuda.synchronize()
        if dist is not None: dist.barrier()
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(REPS): copy()
        torch.cuda.synchronize()
        res.append(REPS * UP / (time.perf_counter() - t) / 1e9)
        if dist is not None: dist.barrier()
        cudart.cudaFreeHost(ptr)
    t = torch.tensor(res, dtype=torch.float64, device="cuda")
    allr = [t]
    if dist is not None:
        allr = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allr, t)
    if rank == 0:
        rows = [[round(float(v), 2) for v in r] for r in allr]
        print(json.dumps({"n_gpus": world, "h2d_GBps_default_pinned": [r[0] for r in rows], "h2d_GBps_write_combined": [r[1] for r in rows],
                          "sum_default": round(sum(r[0] for r in rows), 1), "sum_write_combined": round(sum(r[1] for r in rows), 1)}))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
