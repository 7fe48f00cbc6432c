#!/usr/bin/env python
"""What the box's host fabric delivers when N ranks copy at once: pinned-host <-> device bandwidth per GPU with the bench's
end-to-end transfer sizes (119.5 MB of images up, 33.2 MB of results down per step), alone (N = 1) or concurrently under torchrun:

  python tools/pcie_bw.py
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_bw.py

Rank 0 prints one JSON line: per-GPU GB/s for H2D only, D2H only and both directions at once (all ranks inside the same
barrier-bracketed window), plus the stereo pairs/s those rates could carry at most (the ceiling of bench.py's e2e value)."""
import json, os, time
import torch

UP, DOWN = 119453696, 33162240          # bytes per 128-pair step: bench.py e2e h2d_bytes_per_step / d2h_bytes_per_step
REPS = 20


def main():
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    x = torch.empty(UP, dtype=torch.uint8).pin_memory(); d = torch.empty(UP, dtype=torch.uint8, device="cuda")
    y = torch.empty(DOWN, dtype=torch.uint8, device="cuda"); hy = torch.empty(DOWN, dtype=torch.uint8).pin_memory()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(up, down):
        for _ in range(3):
            if up: d.copy_(x, non_blocking=True)
            if down: hy.copy_(y, non_blocking=True)
        barrier()
        t = time.perf_counter()
        for _ in range(REPS):
            if up:
                with torch.cuda.stream(s1): d.copy_(x, non_blocking=True)
            if down:
                with torch.cuda.stream(s2): hy.copy_(y, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
        barrier()
        return dt

    res = []
    for up, down in ((1, 0), (0, 1), (1, 1)):
        dt = timed(up, down)
        res += [REPS * UP / dt / 1e9 if up else 0.0, REPS * DOWN / dt / 1e9 if down else 0.0, REPS * 128 / dt if (up and down) else 0.0]
    t = torch.tensor(res, dtype=torch.float64, device="cuda")
    allr = [t]
    if dist is not None:
        allr = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allr, t)
    if rank == 0:
        rows = [[round(float(v), 2) for v in r] for r in allr]
        line = {"n_gpus": world, "bytes_up_per_step": UP, "bytes_down_per_step": DOWN,
                "per_gpu": [{"h2d_only_GBps": r[0], "d2h_only_GBps": r[4], "bidir_h2d_GBps": r[6], "bidir_d2h_GBps": r[7], "bidir_pairs_per_s_ceiling": r[8]} for r in rows],
                "sum_bidir_h2d_GBps": round(sum(r[6] for r in rows), 1), "sum_pairs_per_s_ceiling": round(sum(r[8] for r in rows), 0),
                "what": "torch pinned-host <-> device copies, every rank inside the same barrier-bracketed window, 20 repetitions"}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
