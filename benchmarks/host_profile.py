#!/usr/bin/env python
"""Host-side cost of enqueueing one C3 tick (diagnostic): cProfile over the resident-input loop of bench.py,
with and without the pool polls, and the wall / device time of the same loop."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench as B  # noqa: E402
from deepdish_b200.batched import BatchedTracker  # noqa: E402
from deepdish_b200.scene import Scene  # noqa: E402


def main():
    chunks = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    K = 40
    S = B.S_PER_GPU
    bt = BatchedTracker(S, B.LABELS, max_tracks=B.TMAX, max_dets=B.DMAX, budget=B.BUDGET, max_age=B.MAX_AGE, n_chunks=chunks)
    scene = Scene(S, B.N_OBJECTS, B.DMAX, n_labels=len(B.LABELS), seed=1, device="cuda")
    for _ in range(B.PREROLL):
        bt.step(scene.step())
    frames = [scene.step() for _ in range(K)]
    torch.cuda.synchronize()
    for poll in (True, False):
        bt._poll_pool = poll
        for rep in range(2):
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            pr = cProfile.Profile()
            t0 = time.perf_counter()
            s.record()
            if rep:
                pr.enable()
            for b in frames:
                bt.step(b, join=False, reduce=True)
            if rep:
                pr.disable()
            enq = time.perf_counter() - t0
            bt.join()
            e.record()
            torch.cuda.synchronize()
            print("poll=%s rep=%d: host enqueue %.3f ms/tick, device %.3f ms/tick" % (poll, rep, 1e3 * enq / K, s.elapsed_time(e) / K), flush=True)
            if rep:
                pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
    bt.check()


if __name__ == "__main__":
    main()
