#!/usr/bin/env python
"""Host-to-device ceiling of the box with NO compute: what bounds bench.py's end-to-end number at N GPUs.

    python benchmarks/h2d_probe.py                         # one GPU: one pinned 26.6 MB blob over 1..8 copy streams
    torchrun --nproc-per-node 8 benchmarks/h2d_probe.py    # all ranks copy concurrently; then 1, 2, 4 of them alone

Every rank uploads the C3 tick's ragged batch size (26.6 MB, what BatchedTracker.step_host_packed moves per tick) from
pinned host memory to its own GPU, 40 times back to back, all participating ranks started by a barrier.  Rank 0 prints one
JSON line per group size: per-GPU and aggregate GB/s (max time over the participating ranks).  Variants: the pinned
buffer allocated after binding the rank to its GPU's CPUs / NUMA node (NVML affinity), 2 MB-aligned sub-blobs, two copy
streams per rank."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

N_BYTES = 26_600_000


def copy_ms(h, d, n_streams, iters=40):
    streams = [torch.cuda.Stream() for _ in range(n_streams)]
    step = (((N_BYTES + n_streams - 1) // n_streams) + (1 << 21) - 1) >> 21 << 21        # 2 MB-aligned sub-blobs
    cur = torch.cuda.current_stream()

    def go():
        ev = cur.record_event()
        for i, st in enumerate(streams):
            lo, hi = i * step, min(N_BYTES, (i + 1) * step)
            if lo >= hi:
                continue
            st.wait_event(ev)
            with torch.cuda.stream(st):
                d[lo:hi].copy_(h[lo:hi], non_blocking=True)
            cur.wait_event(st.record_event())

    for _ in range(3):
        go()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        go()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    affinity = None
    if "--affinity" in sys.argv:
        try:
            import pynvml
            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
            affinity = len(os.sched_getaffinity(0))
        except Exception as ex:
            affinity = "unavailable: %r" % (ex,)
    h = torch.randint(0, 255, (N_BYTES,), dtype=torch.uint8).pin_memory()        # allocated after the affinity call
    d = torch.empty(N_BYTES, dtype=torch.uint8, device="cuda")
    groups = [g for g in (1, 2, 4, 8) if g <= world] if world > 1 else [1]
    for n_streams in ((1, 2) if world > 1 else (1, 2, 3, 4, 8)):
        for g in groups:
            if world > 1:
                dist.barrier()
            ms = copy_ms(h, d, n_streams) if rank < g else 0.0
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if rank == 0:
                worst = float(t[0])
                print(json.dumps({"ranks_copying": g, "of": world, "copy_streams_per_rank": n_streams, "bytes": N_BYTES,
                                  "ms_max": worst, "GBps_per_gpu": N_BYTES / worst / 1e6,
                                  "GBps_aggregate": g * N_BYTES / worst / 1e6, "cpu_affinity": affinity}), flush=True)
            time.sleep(0.05)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
