"""``deep_sort.preprocessing`` mirror (reference deep_sort/preprocessing.py); dd_nms kernel."""
import numpy as np
import torch

from .. import ops


def non_max_suppression(boxes, max_bbox_overlap, scores=None):
    """preprocessing.py:6-73: indices of the surviving boxes in pick order (descending score; without
    scores the reference ranks by bottom edge y2, which is passed to the kernel as the score)."""
    if len(boxes) == 0:
        return []
    b = np.asarray(boxes).astype(float).reshape(-1, 4)
    s = np.asarray(scores, dtype=np.float32) if scores is not None else (b[:, 1] + b[:, 3]).astype(np.float32)
    n = len(b)
    keep, nkeep = ops.nms(ops._dev(b[None], torch.float64), ops._dev(s[None], torch.float32),
                          torch.tensor([n], dtype=torch.int32, device="cuda"), max_bbox_overlap)
    return keep[0, :int(nkeep[0])].cpu().numpy().astype(int).tolist()
