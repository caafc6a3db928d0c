"""tools/pcie_probe.py -- what the host link(s) of this box sustain, to place bench.py's e2e numbers.

    python tools/pcie_probe.py                                   one GPU
    python -m torch.distributed.run --nproc-per-node N ... tools/pcie_probe.py     N GPUs AT ONCE (one process each)

Every rank copies pinned host memory <-> its own GPU (50 MB H2D pieces, 50 MB D2H pieces on a second stream), alone and
all ranks at the same time, with plain cudaMemcpyAsync through torch.  The sum over the ranks is the ceiling the
host memory path (the VM's memory controllers, PCIe root complexes, IOMMU) puts on any end-to-end pipeline, whatever the
kernels do.  --bind: each rank also pins itself to a distinct block of cores first (NUMA first-touch of the buffers)."""
import json
import os
import sys
import time

import torch


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    bind = "--bind" in sys.argv
    if bind:
        ncpu = os.cpu_count() or 1
        per = max(1, ncpu // max(world, 1))
        os.sched_setaffinity(0, set(range(local * per, min(ncpu, (local + 1) * per))))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist = dist_mod
    n = 4096 * 4096 * 3
    h = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(4)]
    for t in h:
        t.fill_(rank + 1)  # first touch on this rank's cores
    d = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(4)]
    ho = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(4)]
    do = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(4)]
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(up, down, reps=24):
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        for i in range(reps):
            if up:
                with torch.cuda.stream(s1):
                    d[i % 4].copy_(h[i % 4], non_blocking=True)
            if down:
                with torch.cuda.stream(s2):
                    ho[i % 4].copy_(do[i % 4], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return reps * n / dt / 1e9

    out = {"ranks": world, "bind_cores": bind, "host_threads": os.cpu_count()}
    run(True, True, 4)
    for label, up, down in (("h2d_only", True, False), ("d2h_only", False, True), ("both", True, True)):
        per_rank = run(up, down)
        out[label + "_gbs_per_rank"] = round(per_rank, 2)
        out[label + "_gbs_all_ranks" + ("_each_direction" if label == "both" else "")] = round(per_rank * world, 2)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
