"""``deep_sort.nn_matching`` mirror (reference deep_sort/nn_matching.py).  ``distance`` runs the CUDA
nearest-neighbour kernel (dd_nn_distance); the gallery bookkeeping of ``partial_fit`` is host-side like
the reference (the batched tracker keeps its galleries on the device instead)."""
import numpy as np
import torch

from .. import ops


def _as_f32(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float32).reshape(-1, 128 if np.size(x) else 128))


def _nn(x, y, metric):
    x, y = np.asarray(x, dtype=np.float32), np.asarray(y, dtype=np.float32)
    if x.ndim != 2 or y.ndim != 2 or x.shape[1] != 128 or y.shape[1] != 128:
        raise ValueError("deepdish_b200 distance kernels expect [n,128] float32 features")
    off = torch.tensor([0, len(x)], dtype=torch.int32, device="cuda")
    out = ops.nn_distance(ops._dev(x, torch.float32), off, ops._dev(y, torch.float32), metric)
    return out[0].cpu().numpy()


def _nn_euclidean_distance(x, y):
    """nn_matching.py:57-75 -- smallest squared Euclidean distance from each row of y to the samples x."""
    return _nn(x, y, "euclidean")


def _nn_cosine_distance(x, y):
    """nn_matching.py:78-96 -- smallest cosine distance from each row of y to the samples x."""
    return _nn(x, y, "cosine")


class NearestNeighborDistanceMetric(object):
    """nn_matching.py:99-177: per-target sample galleries with an optional budget."""

    def __init__(self, metric, matching_threshold, budget=None):
        if metric not in ("euclidean", "cosine"):
            raise ValueError("Invalid metric; must be either 'euclidean' or 'cosine'")
        self._metric_name = metric
        self._metric = _nn_euclidean_distance if metric == "euclidean" else _nn_cosine_distance
        self.matching_threshold = matching_threshold
        self.budget = budget
        self.samples = {}

    def partial_fit(self, features, targets, active_targets):
        """nn_matching.py:137-154."""
        for feature, target in zip(features, targets):
            self.samples.setdefault(target, []).append(feature)
            if self.budget is not None:
                self.samples[target] = self.samples[target][-self.budget:]
        self.samples = {k: self.samples[k] for k in active_targets}

    def distance(self, features, targets):
        """nn_matching.py:156-177 -> f64 [len(targets), len(features)], one kernel launch."""
        features = np.asarray(features, dtype=np.float32)
        if len(targets) == 0 or len(features) == 0:
            return np.zeros((len(targets), len(features)))
        gal = [np.asarray(self.samples[t], dtype=np.float32).reshape(-1, 128) for t in targets]
        off = np.r_[0, np.cumsum([len(g) for g in gal])].astype(np.int32)
        out = ops.nn_distance(ops._dev(np.concatenate(gal), torch.float32), ops._dev(off, torch.int32),
                              ops._dev(features.reshape(-1, 128), torch.float32), self._metric_name)
        return out.cpu().numpy()
