#!/usr/bin/env python
"""C5-shaped benchmark (BASELINE.json configs[4], one GPU's share): mixed YOLOv5 (25200x85 head) and
SSD-MobileNet (1917 anchors x 91 classes) streams through the whole device pipeline -- head decode + box
filter + NMS + gather + tracker tick + count-line + count reduction.

    python benchmarks/bench_pipeline.py [--yolo 1024] [--ssd 1024] [--steps 10] [--warmup 3]

Heads are resident in HBM (YOLO: 8.57 MB per stream per tick, far larger than L2).  The same heads are
replayed every tick (objects stand still, features = per-object identity + noise), so tracks persist and
galleries fill; `--preroll` ticks run before timing.  Prints one JSON line.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from deepdish_b200.batched import BatchedTracker  # noqa: E402
from deepdish_b200.pipeline import DetectTrackPipeline, YoloFrontEnd, SsdFrontEnd  # noqa: E402
from oracle_free_anchors import ssd_anchors  # noqa: E402

LABELS = ["person", "bicycle", "car", "bus"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--yolo", type=int, default=1024)
    ap.add_argument("--ssd", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--preroll", type=int, default=40)
    ap.add_argument("--objects", type=int, default=40)
    args = ap.parse_args()
    # under torchrun: one rank per GPU, every rank runs its own share of streams (weak scaling) and the [C,4] counters
    # are all-reduced over NCCL after every tick -- C5's only collective
    import torch.distributed as dist
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    gen = torch.Generator(device=dev).manual_seed(rank)
    SY, SS, D = args.yolo, args.ssd, 64
    S = SY + SS
    NA, NC = 25200, 80
    coco = ["person", "bicycle", "car", "motorbike", "aeroplane", "bus"] + ["c%02d" % i for i in range(6, 80)]
    ssd_names = ["???"] + ["c%02d" % i for i in range(1, 91)]
    ssd_names[1], ssd_names[2], ssd_names[3], ssd_names[6] = "person", "bicycle", "car", "bus"
    # YOLO heads: background rows + `objects` confident rows per frame with unique scores
    head = torch.empty((SY, NA, 5 + NC), device=dev)
    for lo in range(0, SY, 64):
        n = min(64, SY - lo)
        h = head[lo:lo + n]
        h[..., 0:2] = 0.1 + 0.8 * torch.rand((n, NA, 2), device=dev, generator=gen)
        h[..., 2:4] = 0.02 + 0.08 * torch.rand((n, NA, 2), device=dev, generator=gen)
        h[..., 4] = 0.2 * torch.rand((n, NA), device=dev, generator=gen)
        h[..., 5:] = 0.5 * torch.rand((n, NA, NC), device=dev, generator=gen)
        rows = torch.rand((n, NA), device=dev, generator=gen).argsort(dim=1)[:, :args.objects]
        fi = torch.arange(n, device=dev)[:, None].expand_as(rows)
        score = 0.5 + 0.49 * (torch.rand((n, args.objects), device=dev, generator=gen).argsort(dim=1).float() + 0.5) / args.objects
        h[fi, rows, 4] = score
        h[fi, rows, 5:] = 0.01
        cls = torch.tensor([0, 1, 2, 5], device=dev)[torch.randint(0, 4, (n, args.objects), device=dev, generator=gen)]
        h[fi, rows, 5 + cls] = 1.0
    rb = 0.3 * torch.randn((SS, 1917, 4), device=dev, generator=gen)
    sc = 0.3 * torch.rand((SS, 1917, 91), device=dev, generator=gen) ** 4         # background: below the 0.5 threshold
    hot = torch.rand((SS, 1917), device=dev, generator=gen).argsort(dim=1)[:, :12]
    fi = torch.arange(SS, device=dev)[:, None].expand_as(hot)
    sc[fi, hot, 1 + torch.tensor([0, 1, 2, 5], device=dev)[torch.randint(0, 4, (SS, 12), device=dev, generator=gen)]] = \
        0.55 + 0.44 * torch.rand((SS, 12), device=dev, generator=gen)
    anchors = torch.from_numpy(ssd_anchors()).to(dev)
    ident = torch.randn((S, D, 128), device=dev, generator=gen)
    ident = ident / ident.norm(dim=-1, keepdim=True)

    bt = BatchedTracker(S, LABELS, max_tracks=128, max_dets=D, budget=100, max_age=60, n_chunks=2, device=dev)
    pipe = DetectTrackPipeline(bt, [
        YoloFrontEnd(0, SY, coco, LABELS, LABELS, ncap=1024),
        SsdFrontEnd(SY, S, ssd_names, LABELS, LABELS, anchors)])

    def feats():
        f = ident + 0.02 * torch.randn((S, D, 128), device=dev, generator=gen)
        return f / f.norm(dim=-1, keepdim=True)

    heads = [head, (rb, sc)]
    fl = [feats() for _ in range(4)]
    for t in range(args.preroll + args.warmup):
        pipe.step(heads, fl[t % 4], join=False)
    bt.join()
    pipe.check()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    for t in range(args.steps):
        pipe.detect(heads)
    ev[1].record()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    ev[2].record()
    for t in range(args.steps):
        pipe.step(heads, fl[t % 4], join=False)
        bt.all_reduce_counts(reduced=True, async_op=True)
    bt.join()
    bt.wait_counts()
    ev[3].record()
    torch.cuda.synchronize()
    pipe.check()
    global_counts = bt.all_reduce_counts(reduced=True).cpu().tolist()
    det_ms = ev[0].elapsed_time(ev[1]) / args.steps
    all_ms = ev[2].elapsed_time(ev[3]) / args.steps
    if world > 1:
        tm = torch.tensor([all_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        all_ms = float(tm[0])
    if rank != 0:
        dist.destroy_process_group()
        return
    yolo_bytes = SY * NA * (5 + NC) * 4 + SS * 1917 * 95 * 4
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    print(json.dumps({
        "workload": "C5 share of one GPU: %d YOLOv5 (25200x85 f32) + %d SSD-MobileNet (1917x91) streams, decode + box "
                    "filter + NMS + gather + tracker tick + count-line + count reduce" % (SY, SS),
        "n_gpus": world, "stream_frames_per_s": world * S / all_ms * 1e3, "ms_per_tick": all_ms, "detect_only_ms": det_ms,
        "detect_head_GBps": yolo_bytes / det_ms / 1e6, "detect_frac_of_measured_hbm": yolo_bytes / det_ms / 1e6 / peak,
        "dets_per_stream": float(pipe.det_count.float().mean()), "tracks_per_stream": float(bt.v["n_tracks"].float().mean()),
        "gallery_rows_per_stream": float(bt.gallery_vectors().float().mean()),
        "counts_all_ranks": global_counts}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
