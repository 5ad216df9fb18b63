"""``deep_sort.linear_assignment`` mirror (reference deep_sort/linear_assignment.py).

The assignment is solved by the CUDA restatement of scipy's solver (dd_lsap, scipy-exact
tie-breaking) and the gate by dd_kalman_gating_distance; the list bookkeeping around them follows the
reference so that the ORDER of the returned lists (which fixes new track ids) is identical."""
import numpy as np
import torch

from .. import ops
from . import kalman_filter

INFTY_COST = 1e+5


def min_cost_matching(distance_metric, max_distance, tracks, detections, track_indices=None,
                      detection_indices=None):
    """linear_assignment.py:11-75 -> (matches, unmatched_tracks, unmatched_detections)."""
    if track_indices is None:
        track_indices = np.arange(len(tracks))
    if detection_indices is None:
        detection_indices = np.arange(len(detections))
    if len(detection_indices) == 0 or len(track_indices) == 0:
        return [], track_indices, detection_indices
    cost = np.array(distance_metric(tracks, detections, track_indices, detection_indices), dtype=np.float64)
    cost[cost > max_distance] = max_distance + 1e-5
    rows, cols = ops.linear_sum_assignment(cost)
    col_of_row = dict(zip(rows.tolist(), cols.tolist()))
    taken = set(cols.tolist())
    matches, unmatched_tracks, unmatched_detections = [], [], []
    for c, d in enumerate(detection_indices):
        if c not in taken:
            unmatched_detections.append(d)
    for r, t in enumerate(track_indices):
        if r not in col_of_row:
            unmatched_tracks.append(t)
    for r, c in zip(rows.tolist(), cols.tolist()):
        t, d = track_indices[r], detection_indices[c]
        if cost[r, c] > max_distance:
            unmatched_tracks.append(t)
            unmatched_detections.append(d)
        else:
            matches.append((t, d))
    return matches, unmatched_tracks, unmatched_detections


def matching_cascade(distance_metric, max_distance, cascade_depth, tracks, detections,
                     track_indices=None, detection_indices=None):
    """linear_assignment.py:78-141."""
    if track_indices is None:
        track_indices = list(range(len(tracks)))
    if detection_indices is None:
        detection_indices = list(range(len(detections)))
    unmatched_detections = detection_indices
    matches = []
    for level in range(cascade_depth):
        if len(unmatched_detections) == 0:
            break
        level_tracks = [k for k in track_indices if tracks[k].time_since_update == 1 + level]
        if len(level_tracks) == 0:
            continue
        m, _, unmatched_detections = min_cost_matching(distance_metric, max_distance, tracks, detections,
                                                       level_tracks, unmatched_detections)
        matches += m
    unmatched_tracks = list(set(track_indices) - set(k for k, _ in matches))
    return matches, unmatched_tracks, unmatched_detections


def gate_cost_matrix(kf, cost_matrix, tracks, detections, track_indices, detection_indices,
                     gated_cost=INFTY_COST, only_position=False):
    """linear_assignment.py:144-190: one batched gating launch for all rows."""
    if len(track_indices) == 0 or len(detection_indices) == 0:
        return cost_matrix
    thr = kalman_filter.chi2inv95[2 if only_position else 4]
    meas = np.asarray([detections[i].to_xyah() for i in detection_indices], dtype=np.float64)
    mean = np.stack([tracks[i].mean for i in track_indices])
    cov = np.stack([tracks[i].covariance for i in track_indices])
    d2 = ops.kalman_gating_distance(ops._dev(mean, torch.float64), ops._dev(cov, torch.float64),
                                    ops._dev(meas, torch.float64), only_position).cpu().numpy()
    cost_matrix[d2 > thr] = gated_cost
    return cost_matrix
