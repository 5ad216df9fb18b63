#!/usr/bin/env python
"""Encoder-input micro-benchmark: extract_image_patch for every detection of a batch of frames,
and the reference's arithmetic DummyImageEncoder on top.

    python benchmarks/bench_patches.py [--frames 1024] [--dets 48] [--iters 20]

Boxes follow the synthetic scene of SURVEY.md 8d (w in [20,40], h in [40,100], 640x480 frames).  Algorithmic
bytes per launch = crop pixels read once (sw*sh*3 per box) + patch bytes written (ph*pw*3 per box); the frames
(0.94 GB at 1024 frames) and the patches (1.2 GB) are both larger than L2.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepdish_b200 import ops  # noqa: E402
from bench_detect import peak, timed  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=1024)
    ap.add_argument("--dets", type=int, default=48)
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    gen = torch.Generator(device="cuda").manual_seed(0)
    B, D, H, W = args.frames, args.dets, 480, 640
    frames = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device="cuda", generator=gen)
    w = torch.randint(20, 41, (B, D), device="cuda", generator=gen)
    h = torch.randint(40, 101, (B, D), device="cuda", generator=gen)
    x = (torch.rand((B, D), device="cuda", generator=gen) * (W - 40)).long()
    y = (torch.rand((B, D), device="cuda", generator=gen) * (H - 100)).long()
    boxes = torch.stack([x, y, w, h], -1).double().contiguous()
    counts = torch.full((B,), D, dtype=torch.int32, device="cuda")
    pk = peak()
    for shape in ((128, 64), (16, 8)):
        buf = {}

        def run(i):
            buf["o"] = ops.extract_patches(frames, boxes, counts, shape, out=buf.get("o"))

        ms = timed(run, args.iters)
        crop = float(((h * 0.5 * shape[1] / shape[0] * 2).long().clamp(max=W) * h).sum()) * 3
        nbytes = crop + B * D * shape[0] * shape[1] * 3
        rec = {"kernel": "k_extract_patches %dx%d" % shape, "frames": B, "boxes": B * D, "ms": ms,
               "patches_per_s": B * D / ms * 1e3, "frames_per_s": B / ms * 1e3, "algorithmic_GBps": nbytes / ms / 1e6,
               "frac_of_measured_hbm": nbytes / ms / 1e6 / pk, "valid": int(buf["o"][1].sum())}
        print(json.dumps(rec))
        if shape == (16, 8):
            p = buf["o"][0]
            fo = {}

            def enc(i):
                fo["f"] = ops.dummy_encode(p, out=fo.get("f"))

            ms2 = timed(enc, args.iters)
            print(json.dumps({"kernel": "k_dummy_encode", "patches": B * D, "ms": ms2,
                              "algorithmic_GBps": B * D * (384 + 512) / ms2 / 1e6}))

            def both(i):
                buf["o"] = ops.extract_patches(frames, boxes, counts, shape, out=buf.get("o"))
                fo["f"] = ops.dummy_encode(buf["o"][0], out=fo.get("f"))

            ms3 = timed(both, args.iters)
            print(json.dumps({"kernel": "dummy box encoder (extract 16x8 + encode)", "frames": B, "ms": ms3,
                              "frames_per_s": B / ms3 * 1e3}))


if __name__ == "__main__":
    main()
