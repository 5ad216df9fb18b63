"""``deep_sort.iou_matching`` mirror (reference deep_sort/iou_matching.py); dd_iou_cost kernel."""
import numpy as np
import torch

from .. import ops
from . import linear_assignment  # noqa: F401  (INFTY_COST lives there, like the reference)


def iou(bbox, candidates):
    """iou_matching.py:7-39: IoU of one tlwh box against candidate tlwh boxes [N,4] -> [N]."""
    cand = np.asarray(candidates, dtype=np.float64).reshape(-1, 4)
    if len(cand) == 0:
        return np.zeros(0)
    cost = ops.iou_cost(ops._dev(np.asarray(bbox, dtype=np.float64).reshape(1, 4), torch.float64),
                        torch.zeros(1, dtype=torch.int32, device="cuda"), ops._dev(cand, torch.float64))
    return 1. - cost[0].cpu().numpy()


def iou_cost(tracks, detections, track_indices=None, detection_indices=None):
    """iou_matching.py:42-81: 1 - IoU cost matrix, rows of tracks with time_since_update > 1 = INFTY."""
    if track_indices is None:
        track_indices = np.arange(len(tracks))
    if detection_indices is None:
        detection_indices = np.arange(len(detections))
    if len(track_indices) == 0 or len(detection_indices) == 0:
        return np.zeros((len(track_indices), len(detection_indices)))
    trk = np.stack([tracks[i].to_tlwh() for i in track_indices])
    tsu = np.array([tracks[i].time_since_update for i in track_indices], dtype=np.int32)
    det = np.stack([detections[i].tlwh for i in detection_indices])
    return ops.iou_cost(ops._dev(trk, torch.float64), ops._dev(tsu, torch.int32),
                        ops._dev(det, torch.float64)).cpu().numpy()
