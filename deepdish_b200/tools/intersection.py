"""``tools.intersection`` mirror (reference tools/intersection.py:4-30): segment crossing tests on the
CUDA device function dd_segments_intersect (the one the count-line kernel uses)."""
import numpy as np
import torch

from .. import ops


def intersection(p, pr, q, qs):
    """tools/intersection.py:4-24: do segments p->pr and q->qs intersect?"""
    seg = np.concatenate([np.asarray(x, dtype=np.float64).reshape(2) for x in (p, pr, q, qs)])[None]
    return bool(ops.segments_intersect(ops._dev(seg, torch.float64))[0].item())


def any_intersection(p1, q1, pts):
    """tools/intersection.py:26-30: any consecutive pair of the polyline crosses p1->q1 (one launch)."""
    pts = np.asarray(pts, dtype=np.float64).reshape(-1, 2)
    if len(pts) < 2:
        return False
    p1, q1 = np.asarray(p1, dtype=np.float64).reshape(2), np.asarray(q1, dtype=np.float64).reshape(2)
    seg = np.concatenate([np.tile(np.r_[p1, q1], (len(pts) - 1, 1)), pts[:-1], pts[1:]], axis=1)
    return bool(ops.segments_intersect(ops._dev(seg, torch.float64)).any().item())
