#!/usr/bin/env python
"""Kernel timeline of the C3 tick as the product runs it (native engine, one captured graph per chunk and tick; no
nsys in the image): with config.timeline every tick kernel stamps the start of its first CTA and the end of its last
one (%globaltimer) into the blob; printed per tick and chunk relative to the first stamp, plus the time each kernel
pair spends overlapped.

    python benchmarks/timeline.py [--chunks 2] [--ticks 6] [bench.py's A/B knobs]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench as B  # noqa: E402
from deepdish_b200.batched import BatchedTracker  # noqa: E402
from deepdish_b200.scene import Scene  # noqa: E402

NAMES = ["prep", "gate", "gallery", "match", "apply", "countline"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunks", type=int, default=2)
    ap.add_argument("--ticks", type=int, default=6)
    ap.add_argument("--cosine-ctas", type=int, default=0)
    ap.add_argument("--gallery-stages", type=int, default=0)
    ap.add_argument("--gallery-waves", type=int, default=0)
    ap.add_argument("--gallery-impl", default="default")
    ap.add_argument("--no-turns", action="store_true")
    ap.add_argument("--host", action="store_true", help="end-to-end ticks from ragged pinned host batches (step_host_packed)")
    a = ap.parse_args()
    S = B.S_PER_GPU
    bt = BatchedTracker(S, B.LABELS, max_tracks=B.TMAX, max_dets=B.DMAX, budget=B.BUDGET, max_age=B.MAX_AGE,
                        n_chunks=a.chunks, cosine_ctas_per_sm=a.cosine_ctas, gallery_stages=a.gallery_stages,
                        gallery_impl=a.gallery_impl, gallery_waves=a.gallery_waves, timeline=1,
                        gallery_turns=not a.no_turns)
    scene = Scene(S, B.N_OBJECTS, B.DMAX, n_labels=len(B.LABELS), seed=1, device="cuda")
    for _ in range(B.PREROLL):
        bt.step(scene.step())
    frames = [scene.step() for _ in range(8 + a.ticks)]
    torch.cuda.synchronize()
    for c in bt.chunks:
        c.v["timeline"][:, :, 0] = torch.iinfo(torch.int64).max
        c.v["timeline"][:, :, 1] = 0
        c.v["timeline"][:, 7, 0] = 0          # slot 7: latest k_match CTA start (max), longest single k_match CTA (max)
    torch.cuda.synchronize()
    t0 = bt.engine_stats()[0]
    if a.host:
        ids_host = torch.empty((S, B.DMAX), dtype=torch.int32).pin_memory()
        packed = [bt.pack_host(b) for b in frames]
        torch.cuda.synchronize()
        for hb in packed:
            bt.step_host_packed(hb, ids_host)
    else:
        for b in frames:
            bt.step(b, join=False, reduce=True)
    bt.join()
    torch.cuda.synchronize()
    tl = [c.v["timeline"].cpu().numpy() for c in bt.chunks]
    ticks = [(t0 + 8 + k) & 63 for k in range(a.ticks)]
    base = min(int(tl[c][ticks[0], k, 0]) for c in range(a.chunks) for k in range(6))
    iv = {}
    for t in ticks:
        for c in range(a.chunks):
            row = []
            for k in range(6):
                s, e = (int(tl[c][t, k, 0]) - base) / 1e6, (int(tl[c][t, k, 1]) - base) / 1e6
                iv[(t, c, k)] = (s, e)
                row.append("%s %.3f-%.3f (%.3f)" % (NAMES[k], s, e, e - s))
            print("tick %2d chunk %d: " % (t, c) + "  ".join(row))
    first = min(iv[(ticks[0], c, 0)][0] for c in range(a.chunks))
    last = min(iv[(ticks[-1], c, 0)][0] for c in range(a.chunks))
    print("ms per tick over the window: %.4f" % ((last - first) / (a.ticks - 1)))
    # mean duration of every kernel, and how much of each match interval another chunk's gallery stream covers
    for k in range(6):
        d = [iv[(t, c, k)][1] - iv[(t, c, k)][0] for t in ticks for c in range(a.chunks)]
        print("%-10s mean %.4f ms  (min %.4f max %.4f)" % (NAMES[k], sum(d) / len(d), min(d), max(d)))
    late, longest = [], []
    for t in ticks:
        for c in range(a.chunks):
            late.append((int(tl[c][t, 7, 0]) - base) / 1e6 - iv[(t, c, 3)][0])
            longest.append(int(tl[c][t, 7, 1]) / 1e6)
    print("k_match: last CTA starts %.4f ms after the first (mean; max %.4f); longest single CTA %.4f ms (mean; max %.4f)"
          % (sum(late) / len(late), max(late), sum(longest) / len(longest), max(longest)))
    cov = []
    for t in ticks:
        for c in range(a.chunks):
            ms, me = iv[(t, c, 3)]
            got = 0.0
            for (t2, c2, k2), (s2, e2) in iv.items():
                if k2 == 2 and c2 != c:
                    got += max(0.0, min(me, e2) - max(ms, s2))
            cov.append(got / max(1e-9, me - ms))
    print("share of the match intervals covered by another chunk's gallery stream: %.2f" % (sum(cov) / len(cov)))
    bt.check()


if __name__ == "__main__":
    main()
