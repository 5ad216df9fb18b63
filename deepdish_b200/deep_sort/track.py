"""``deep_sort.track`` mirror (reference deep_sort/track.py).  A Track is a host snapshot of one slot of
the device-resident tracker state (materialised by Tracker after every predict / update); its methods
that do arithmetic (predict / update) call the CUDA Kalman kernels."""
import numpy as np


class TrackState:
    """track.py:5-17."""
    Tentative = 1
    Confirmed = 2
    Deleted = 3


class Track:
    def __init__(self, mean, covariance, track_id, n_init, max_age, detection=None):
        self.mean = mean
        self.covariance = covariance
        self.track_id = track_id
        self.hits = 1
        self.age = 1
        self.time_since_update = 0
        self.state = TrackState.Tentative
        self.features = []
        self.labels = []
        self.dist = {}
        self.detections = []
        if detection is not None:                      # track.py:75-80
            self.features.append(detection.feature)
            self.labels.append(detection.label)
            self.dist[detection.label] = [detection.confidence]
            self.detections.append(detection)
        self._n_init = n_init
        self._max_age = max_age
        # label statistics mirrored from the device state: {label: (count, sum of confidences)}
        self._label_stats = None
        # detections applied on the host since the last device call (Track.update called from outside the
        # tracker, e.g. FrameRecords.process_tracking): Tracker._flush writes them back to the device state
        self._pending = []

    def to_tlwh(self):
        """track.py:84-97."""
        ret = self.mean[:4].copy()
        ret[2] *= ret[3]
        ret[:2] -= ret[2:] / 2
        return ret

    def to_tlbr(self):
        """track.py:99-111."""
        ret = self.to_tlwh()
        ret[2:] = ret[:2] + ret[2:]
        return ret

    def predict(self, kf):
        """track.py:113-125."""
        self.mean, self.covariance = kf.predict(self.mean, self.covariance)
        self.age += 1
        self.time_since_update += 1

    def update(self, kf, detection):
        """track.py:127-152."""
        self.mean, self.covariance = kf.update(self.mean, self.covariance, detection.to_xyah())
        self.features.append(detection.feature)
        self.hits += 1
        self.time_since_update = 0
        if self.state == TrackState.Tentative and self.hits >= self._n_init:
            self.state = TrackState.Confirmed
        self.labels.append(detection.label)
        self.dist.setdefault(detection.label, []).append(detection.confidence)
        self.detections.append(detection)
        if self._label_stats is not None:
            c, s = self._label_stats.get(detection.label, (0, 0.0))
            self._label_stats[detection.label] = (c + 1, s + detection.confidence)
        self._pending.append(detection)

    def _stats(self):
        if self._label_stats is not None:
            return [(l, c, s / c) for l, (c, s) in self._label_stats.items() if c > 0]
        return [(l, len(v), float(np.average(v))) for l, v in self.dist.items()]

    def get_label(self, return_confidence=False):
        """track.py:154-188: Dirichlet-expected vote, reverse (value, name) order, motorbike rule."""
        stats = self._stats()
        if not stats:
            return (None, 0) if return_confidence else None
        alphas = np.array([a for _, _, a in stats])
        cnt = np.array([c for _, c, _ in stats])
        ranked = sorted(zip((alphas + cnt) / (cnt.sum() + alphas.sum()), [l for l, _, _ in stats]), reverse=True)
        pick = ranked[0][1]
        if len(ranked) > 1 and ranked[0][1] == 'motorbike' and ranked[1][1] == 'bicycle':
            pick = 'motorbike' if ranked[0][0] > ranked[1][0] * 4 else 'bicycle'
        if return_confidence:
            return pick, dict((l, a) for l, _, a in stats)[pick]
        return pick

    def mark_missed(self):
        """track.py:190-196."""
        if self.state == TrackState.Tentative:
            self.state = TrackState.Deleted
        elif self.time_since_update > self._max_age:
            self.state = TrackState.Deleted

    def is_tentative(self):
        return self.state == TrackState.Tentative

    def is_confirmed(self):
        return self.state == TrackState.Confirmed

    def is_deleted(self):
        return self.state == TrackState.Deleted
