"""``deep_sort.kalman_filter`` mirror (reference deep_sort/kalman_filter.py).  Same signatures, numpy in
/ numpy out; every method is one launch of the f64 CUDA kernels in csrc/dd_kalman.cuh."""
import numpy as np
import torch

from .. import ops

# 0.95 quantiles of the chi-square distribution, kalman_filter.py:11-20
chi2inv95 = {1: 3.8415, 2: 5.9915, 3: 7.8147, 4: 9.4877, 5: 11.070, 6: 12.592, 7: 14.067,
             8: 15.507, 9: 16.919}


def _d(a, shape):
    return ops._dev(np.asarray(a, dtype=np.float64).reshape(shape), torch.float64)


class KalmanFilter(object):
    """8-state constant-velocity filter over (x, y, a, h) boxes (kalman_filter.py:23-53)."""

    def __init__(self):
        self._std_weight_position = 1. / 20
        self._std_weight_velocity = 1. / 160

    def initiate(self, measurement):
        """kalman_filter.py:55-86 -> (mean[8], covariance[8,8])."""
        mean, cov = ops.kalman_initiate(_d(measurement, (1, 4)))
        return mean[0].cpu().numpy(), cov[0].cpu().numpy()

    def predict(self, mean, covariance):
        """kalman_filter.py:88-123."""
        m, c = ops.kalman_predict_(_d(mean, (1, 8)), _d(covariance, (1, 8, 8)))
        return m[0].cpu().numpy(), c[0].cpu().numpy()

    def project(self, mean, covariance):
        """kalman_filter.py:125-152 -> (mean[4], covariance[4,4])."""
        m, c = ops.kalman_project(_d(mean, (1, 8)), _d(covariance, (1, 8, 8)))
        return m[0].cpu().numpy(), c[0].cpu().numpy()

    def update(self, mean, covariance, measurement):
        """kalman_filter.py:154-186."""
        m, c = ops.kalman_update_(_d(mean, (1, 8)), _d(covariance, (1, 8, 8)), _d(measurement, (1, 4)))
        return m[0].cpu().numpy(), c[0].cpu().numpy()

    def gating_distance(self, mean, covariance, measurements, only_position=False):
        """kalman_filter.py:188-229 -> squared Mahalanobis distance per measurement [N]."""
        meas = np.asarray(measurements, dtype=np.float64).reshape(-1, 4)
        if len(meas) == 0:
            return np.zeros(0)
        out = ops.kalman_gating_distance(_d(mean, (1, 8)), _d(covariance, (1, 8, 8)), _d(meas, (-1, 4)),
                                         only_position)
        return out[0].cpu().numpy()
