#!/usr/bin/env python
"""Kernel timeline of the chunked tick (no nsys in the image): every chunk's update is launched through
dd_tracker_update_profiled, whose CUDA events (recorded on the chunk's stream between the kernels) give each
kernel's completion time relative to a common origin.  Prints, per tick and chunk, the end time of
prep / gate / cosine / match / apply in microseconds.

    python benchmarks/timeline.py [--chunks 4] [--ticks 3]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepdish_b200 import _lib  # noqa: E402
from deepdish_b200.batched import BatchedTracker  # noqa: E402
from deepdish_b200.scene import Scene  # noqa: E402
import bench as B  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunks", type=int, default=4)
    ap.add_argument("--ticks", type=int, default=3)
    ap.add_argument("--chain", type=int, default=0)
    ap.add_argument("--cosine-ctas", type=int, default=None)
    ap.add_argument("--cs", type=int, default=None, help="A/B knob: 1 = streaming (evict-first) gallery loads")
    ap.add_argument("--match-cta", type=int, default=None, help="A/B knob: 1 = four warps per stream in k_match, 0 = one")
    ap.add_argument("--pdl", type=int, default=None, help="A/B knob: programmatic dependent launch on / off")
    ap.add_argument("--prio", type=int, default=None)
    ap.add_argument("--gate-impl", type=int, default=None)
    args = ap.parse_args()
    for key, val in ((0, args.gate_impl), (1, args.cosine_ctas), (2, args.prio), (3, args.cs), (4, args.pdl), (5, args.match_cta)):
        if val is not None:
            _lib.check(_lib.lib().dd_tuning_set(key, val), "dd_tuning_set")
    dev = torch.device("cuda", 0)
    S = B.S_PER_GPU
    bt = BatchedTracker(S, B.LABELS, max_tracks=B.TMAX, max_dets=B.DMAX, budget=B.BUDGET, max_age=B.MAX_AGE,
                        device=dev, n_chunks=args.chunks)
    scene = Scene(S, B.N_OBJECTS, B.DMAX, n_labels=len(B.LABELS), seed=1234, device=dev)
    for _ in range(B.PREROLL):
        bt.step(scene.step())
    frames = [scene.step() for _ in range(args.ticks + 3)]
    for b in frames[:3]:
        bt.step(b, join=False)
    bt.join()
    torch.cuda.synchronize()
    evs = [[bt.new_events(6) for _ in bt.chunks] for _ in range(args.ticks)]
    lib = bt.lib
    o = torch.cuda.Event(enable_timing=True)
    o.record()
    bt._fork()
    for k, b in enumerate(frames[3:]):
        for i, c in enumerate(bt.chunks):
            p = bt._ptrs(c, b.tlwh, b.conf, b.label, b.feat, b.count, bt.det_track_id)
            sp = bt._sp(c)
            _lib.check(lib.dd_tracker_predict(c.state, c.cfgp, sp), "predict")
            _lib.check(lib.dd_tracker_update_profiled(c.state, c.cfgp, *p, sp, evs[k][i],
                                                      c.turn_wait if args.chain else None,
                                                      c.turn_done if args.chain else None), "update")
            _lib.check(lib.dd_tracker_countline(c.state, c.cfgp, bt._line_ptr(c), bt.line_per_stream, sp), "cl")
    bt.join()
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    torch.cuda.synchronize()
    print("total ms for %d ticks: %.3f" % (args.ticks, o.elapsed_time(e)))
    base = evs[0][0][0]
    names = ["start", "prep", "gate", "cosine", "match", "apply"]
    for k in range(args.ticks):
        for i in range(len(bt.chunks)):
            ts = [bt.elapsed_ms(base, evs[k][i][j]) * 1e3 for j in range(6)]
            print("tick %d chunk %d  " % (k, i) + "  ".join("%s %7.1f" % (n, t) for n, t in zip(names, ts)) +
                  "   | dur " + " ".join("%6.1f" % (ts[j + 1] - ts[j]) for j in range(5)))


if __name__ == "__main__":
    main()
