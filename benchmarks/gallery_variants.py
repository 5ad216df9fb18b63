#!/usr/bin/env python
"""Where does the gallery stream's time go?  Elimination on the real workload: build with -DDD_GS_VARIANTS, bring the C3
tracker to steady state, then REPLAY the gallery kernel on the last tick's work list with parts switched off
(dd_gallery_replay; the variants' wrong costs are never read) and time each variant with CUDA events.

    DD_NVCC_EXTRA=-DDD_GS_VARIANTS python -m deepdish_b200.build --force && python benchmarks/gallery_variants.py"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench as B  # noqa: E402
from deepdish_b200 import _lib  # noqa: E402
from deepdish_b200.batched import BatchedTracker  # noqa: E402
from deepdish_b200.scene import Scene  # noqa: E402

NAMES = {0: "product kernel", 1: "checker evaluates nothing", 2: "no candidate listing", 3: "1 + 2",
         4: "no fragment loads / mma", 7: "1 + 2 + 4", 8: "one query row per job", 15: "1 + 2 + 4 + 8",
         16: "ragged last page copied whole", 31: "everything off"}


def main():
    S = B.S_PER_GPU
    bt = BatchedTracker(S, B.LABELS, max_tracks=B.TMAX, max_dets=B.DMAX, budget=B.BUDGET, max_age=B.MAX_AGE, n_chunks=1)
    lib = bt.lib
    if not hasattr(lib, "dd_gallery_replay"):
        raise SystemExit("build with DD_NVCC_EXTRA=-DDD_GS_VARIANTS first")
    lib.dd_gallery_replay.argtypes = [ctypes.c_void_p, ctypes.POINTER(_lib.TrackerConfig), ctypes.c_int, ctypes.c_void_p]
    scene = Scene(S, B.N_OBJECTS, B.DMAX, n_labels=len(B.LABELS), seed=1234, device="cuda")
    for _ in range(B.PREROLL + 8):
        bt.step(scene.step())
    torch.cuda.synchronize()
    c = bt.chunks[0]
    n = int(c.v["work_ctl"][0])
    rec = c.v["work_rec"][:n]
    rows, cands = int(rec[:, 2].sum()), int(rec[:, 3].sum())
    must = 256.0 * rows + 260.0 * cands + 64.0 * n
    sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    prof = (ctypes.c_uint64 * 16)()
    lib.dd_gallery_prof.argtypes = [ctypes.POINTER(ctypes.c_uint64), ctypes.c_void_p]
    for skip in (0, 1, 2, 3, 4, 7, 8, 15, 16, 31, 0):
        for _ in range(3):
            _lib.check(lib.dd_gallery_replay(c.state, c.cfgp, skip, sp), "dd_gallery_replay")
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        ms = 0.0
        for _ in range(10):
            # L2 holds 126 MB of the 1.2 GB a launch streams: consecutive launches do not help each other
            s.record()
            _lib.check(lib.dd_gallery_replay(c.state, c.cfgp, skip, sp), "dd_gallery_replay")
            e.record()
            torch.cuda.synchronize()
            ms += s.elapsed_time(e) / 10
        lib.dd_gallery_prof(prof, sp)                      # reset
        _lib.check(lib.dd_gallery_replay(c.state, c.cfgp, skip, sp), "dd_gallery_replay")
        lib.dd_gallery_prof(prof, sp)
        p = [int(x) for x in prof]
        frac = lambda a, b: round(a / max(1, b), 3)
        roles = {"producer": {"wait_empty": frac(p[1], p[0]), "wait_hfree": frac(p[2], p[0])},
                 "mma": {"wait_full": frac(p[5], p[4]), "wait_hfull": frac(p[6], p[4]), "wait_mfree": frac(p[7], p[4])},
                 "checker": {"wait_mfull": frac(p[9], p[8])}}
        print(json.dumps({"skip": skip, "what": NAMES[skip], "ms": round(ms, 4), "work_items": n, "gallery_rows": rows,
                          "must_move_GB": round(must / 1e9, 4), "GBps_of_must_move": round(must / ms / 1e6, 1),
                          "share_of_role_time_in_waits": roles}), flush=True)


if __name__ == "__main__":
    main()
