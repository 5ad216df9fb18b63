#!/usr/bin/env python
"""Kernel timeline of the C3 tick with several stream chunks (no nsys in the image): CUDA events recorded between the
kernels of every chunk (dd_tracker_tick_profiled) over a few steady-state ticks, printed relative to the first one.

    python benchmarks/timeline.py [--chunks 2] [--ticks 4] [bench.py's A/B knobs]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench as B  # noqa: E402
from deepdish_b200.batched import BatchedTracker  # noqa: E402
from deepdish_b200.scene import Scene  # noqa: E402

NAMES = ["prep", "gate", "gallery", "match", "apply", "countline", "reduce"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunks", type=int, default=2)
    ap.add_argument("--ticks", type=int, default=4)
    ap.add_argument("--cosine-ctas", type=int, default=0)
    ap.add_argument("--gallery-stages", type=int, default=0)
    ap.add_argument("--gallery-impl", default="default")
    a = ap.parse_args()
    S = B.S_PER_GPU
    bt = BatchedTracker(S, B.LABELS, max_tracks=B.TMAX, max_dets=B.DMAX, budget=B.BUDGET, max_age=B.MAX_AGE,
                        n_chunks=a.chunks, cosine_ctas_per_sm=a.cosine_ctas, gallery_stages=a.gallery_stages,
                        gallery_impl=a.gallery_impl)
    scene = Scene(S, B.N_OBJECTS, B.DMAX, n_labels=len(B.LABELS), seed=1, device="cuda")
    for _ in range(B.PREROLL):
        bt.step(scene.step())
    warm = [scene.step() for _ in range(6)]
    frames = [scene.step() for _ in range(a.ticks)]
    for b in warm:
        bt.step(b, join=False, reduce=True)
    evs = [[bt.new_events(8) for _ in range(a.chunks)] for _ in range(a.ticks)]
    bt._timeline = [e for tick in evs for e in tick]
    for b in frames:
        bt.step(b, join=False, reduce=True)
    bt.join()
    torch.cuda.synchronize()
    base = evs[0][0][0]
    for t in range(a.ticks):
        for c in range(a.chunks):
            ts = [bt.elapsed_ms(base, evs[t][c][i]) for i in range(8)]
            print("tick %d chunk %d: start %7.3f | " % (t, c, ts[0]) +
                  "  ".join("%s %.3f-%.3f" % (NAMES[i], ts[i], ts[i + 1]) for i in range(7)))
    last = max(bt.elapsed_ms(base, evs[-1][c][7]) for c in range(a.chunks))
    first = min(bt.elapsed_ms(base, evs[0][c][0]) for c in range(a.chunks))
    print("ms per tick over the window: %.3f" % ((last - first) / a.ticks))
    bt.check()


if __name__ == "__main__":
    main()
