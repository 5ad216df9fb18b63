"""``deep_sort.tracker`` mirror (reference deep_sort/tracker.py): ``Tracker`` is the S = 1 view of
``deepdish_b200.batched.BatchedTracker``.  ``predict()`` / ``update(detections)`` launch the same CUDA
kernels as the batched path; ``tracks`` / ``deleted_tracks`` are host snapshots (lists of Track) rebuilt
from the device state after each call, in the reference's list order."""
import numpy as np
import torch

from .. import _lib
from ..batched import BatchedTracker
from . import kalman_filter
from .track import Track, TrackState  # noqa: F401  (TrackState re-exported like the reference module)


class Tracker:
    MAX_TRACKS = 256          # slots (live + deleted-this-tick + new); overflow raises, never truncates
    MAX_DETS = 256

    def __init__(self, metric, max_iou_distance=0.7, max_age=30, n_init=3):
        self.metric = metric
        self.max_iou_distance = max_iou_distance
        self.max_age = max_age
        self.n_init = n_init
        self.kf = kalman_filter.KalmanFilter()
        self._tracks = []
        self._device_ids = []                  # track ids of the device's live list, in its order
        self._late_features = []               # (track_id, unit feature) to enter the gallery after the next matching
        self.deleted_tracks = []
        self._labels = []                      # label vocabulary in first-seen order
        self._bt = None
        if getattr(metric, "_metric_name", "cosine") != "cosine":
            # the device tick computes the cosine metric only (deepdish.py:516 never builds another one); the
            # euclidean metric stays available through NearestNeighborDistanceMetric.distance
            raise NotImplementedError("Tracker runs the 'cosine' metric on the device; got %r" % metric._metric_name)

    # -- host edits written back to the device ------------------------------------------------
    @property
    def tracks(self):
        return self._tracks

    @tracks.setter
    def tracks(self, value):
        """``tracker.tracks = [...]`` (deepdish.py:1047): tracks missing from the new list are dropped from the
        device's live list at the next device call."""
        self._tracks = list(value)

    def _flush(self):
        """Write host-side edits back before the device runs again: tracks dropped from ``tracks`` leave the live
        list (their slots become free), and tracks updated through ``Track.update(kf, detection)`` outside the tracker
        (framerecords.py:157-160) get their Kalman state, hits, time_since_update, state and label votes stored.
        Their features join the gallery after the NEXT matching, exactly when the reference's partial_fit would
        move them from ``track.features`` to ``metric.samples`` (tracker.py:84-93)."""
        if self._bt is None:
            return
        v = self._bt.chunks[0].v
        keep = {t.track_id for t in self._tracks}
        if any(tid not in keep for tid in self._device_ids):
            n = int(v["n_tracks"][0])
            order = v["order"][0, :n].cpu().numpy()
            ids = v["track_id"][0].cpu().numpy()
            live = [int(s) for s in order if int(ids[s]) in keep]
            for s in order:
                if int(ids[s]) not in keep:
                    v["state"][0, int(s)] = 0                                   # DD_STATE_FREE
            if live:
                v["order"][0, :len(live)] = torch.as_tensor(live, dtype=torch.int32, device=v["order"].device)
            v["n_tracks"][0] = len(live)
            self._device_ids = [int(ids[s]) for s in live]
        for t in self._tracks:
            if not t._pending:
                continue
            s = t._slot
            dev = v["mean"].device
            v["mean"][0, s] = torch.as_tensor(np.asarray(t.mean, np.float64), device=dev)
            v["cov"][0, s] = torch.as_tensor(np.asarray(t.covariance, np.float64), device=dev)
            v["hits"][0, s], v["tsu"][0, s], v["state"][0, s] = int(t.hits), int(t.time_since_update), int(t.state)
            for d in t._pending:
                c = self._label_id(d.label)
                v["lab_cnt"][0, s, c] += 1
                v["lab_sum"][0, s, c] += float(d.confidence)
                f = np.asarray(d.feature, np.float32)
                self._late_features.append((t.track_id, f / np.float32(np.sqrt(np.sum(f * f, dtype=np.float32)))))
            t._pending = []

    def _apply_late_features(self, matched_ids):
        """Gallery surgery after a matching: a late feature goes in BEFORE the feature the update just appended for
        the same track (the order partial_fit would produce), or at the end if the track was not matched."""
        if not self._late_features:
            return
        slot_of = {t.track_id: t._slot for t in self._tracks}
        for tid, f in self._late_features:
            if tid in slot_of:
                self._bt.gallery_insert(0, slot_of[tid], f, before_newest=tid in matched_ids)
        self._late_features = []

    # -- device state -----------------------------------------------------------------------
    def _ensure(self):
        if self._bt is None:
            names = ["\x00unused%03d" % i for i in range(_lib.DD_MAX_LABELS)]
            self._bt = BatchedTracker(1, names, max_tracks=self.MAX_TRACKS, max_dets=self.MAX_DETS,
                                      budget=self.metric.budget, max_cosine_distance=self.metric.matching_threshold,
                                      max_iou_distance=self.max_iou_distance, max_age=self.max_age,
                                      n_init=self.n_init, pool_pages=2048, seg_pages=2048, page_cap=64)
        return self._bt

    @property
    def _next_id(self):
        return 1 if self._bt is None else int(self._bt.v["next_id"][0])

    def _label_id(self, label):
        if label not in self._labels:
            if len(self._labels) >= _lib.DD_MAX_LABELS:
                raise RuntimeError("more than %d distinct labels" % _lib.DD_MAX_LABELS)
            self._labels.append(label)
            cfg = self._bt.cfg                  # ranks follow ascending label-name order (track.py:170)
            order = sorted(range(len(self._labels)), key=lambda i: self._labels[i])
            for r, i in enumerate(order):
                cfg.label_rank[i] = r
            cfg.label_motorbike = self._labels.index("motorbike") if "motorbike" in self._labels else -1
            cfg.label_bicycle = self._labels.index("bicycle") if "bicycle" in self._labels else -1
        return self._labels.index(label)

    def _snapshot(self):
        bt = self._bt
        v = bt.host_view(["n_tracks", "n_deleted", "order", "deleted", "mean", "cov", "track_id", "hits", "age",
                          "tsu", "state", "lab_cnt", "lab_sum", "gal_len", "gal_pos"])
        old = {t.track_id: t for t in self._tracks + self.deleted_tracks}

        def make(slot):
            tid = int(v["track_id"][0, slot])
            t = old.get(tid) or Track(None, None, tid, self.n_init, self.max_age)
            t.mean = v["mean"][0, slot].copy()
            t.covariance = v["cov"][0, slot].copy()
            t.hits, t.age = int(v["hits"][0, slot]), int(v["age"][0, slot])
            t.time_since_update, t.state = int(v["tsu"][0, slot]), int(v["state"][0, slot])
            t._slot = int(slot)
            t._label_stats = {self._labels[c]: (int(v["lab_cnt"][0, slot, c]), float(v["lab_sum"][0, slot, c]))
                              for c in range(len(self._labels)) if v["lab_cnt"][0, slot, c] > 0}
            return t

        self._tracks = [make(s) for s in v["order"][0, :int(v["n_tracks"][0])]]
        self._device_ids = [t.track_id for t in self._tracks]
        self.deleted_tracks = [make(s) for s in v["deleted"][0, :int(v["n_deleted"][0])]]
        self.metric.samples = _LazySamples(self)

    # -- reference API ----------------------------------------------------------------------
    def predict(self):
        """tracker.py:51-57."""
        if self._bt is None:
            return
        self._flush()
        self._bt.predict()
        self._snapshot()

    def update(self, detections):
        """tracker.py:59-93."""
        bt = self._ensure()
        self._flush()
        D = bt.cfg.max_dets
        n = len(detections)
        if n > D:
            raise RuntimeError("detection capacity exceeded (%d > %d)" % (n, D))
        tlwh = np.zeros((1, D, 4)); conf = np.zeros((1, D), np.float32)
        lab = np.zeros((1, D), np.int32); feat = np.zeros((1, D, 128), np.float32)
        for i, d in enumerate(detections):
            tlwh[0, i], conf[0, i], lab[0, i] = d.tlwh, d.confidence, self._label_id(d.label)
            feat[0, i] = d.feature
        ids = bt.update(torch.from_numpy(tlwh).cuda(), torch.from_numpy(conf).cuda(), torch.from_numpy(lab).cuda(),
                        torch.from_numpy(feat).cuda(), torch.tensor([n], dtype=torch.int32, device="cuda"))
        self.last_detection_track_ids = ids[0, :n].cpu().numpy()
        if bt._poll_pool:
            bt.maintain(wait=True)              # unbounded galleries (nn_budget=None): grow pool / page table
        bt.check()
        self._snapshot()
        self._apply_late_features({int(i) for i in self.last_detection_track_ids})
        by_id = {t.track_id: t for t in self.tracks + self.deleted_tracks}
        for i, d in enumerate(detections):              # host-side caches kept like track.py:75-80,147-151
            t = by_id.get(int(self.last_detection_track_ids[i]))
            if t is not None:
                t.labels.append(d.label)
                t.dist.setdefault(d.label, []).append(d.confidence)
                t.detections.append(d)
                t.features = [] if t.is_confirmed() else t.features + [d.feature]


class _LazySamples(dict):
    """``metric.samples``: {track_id: [unit-normalised gallery vectors, oldest first]} read from the
    device ring on first access (nn_matching.py:132-154 keeps only confirmed targets)."""

    def __init__(self, trk):
        super().__init__()
        self._trk, self._loaded = trk, False

    def _load(self):
        if self._loaded:
            return
        self._loaded = True
        trk = self._trk
        for t in trk.tracks:
            if t.is_confirmed():
                dict.__setitem__(self, t.track_id, list(trk._bt.gallery(0, t._slot).cpu().numpy()))

    def __getitem__(self, k):
        self._load(); return dict.__getitem__(self, k)

    def __iter__(self):
        self._load(); return dict.__iter__(self)

    def __len__(self):
        self._load(); return dict.__len__(self)

    def keys(self):
        self._load(); return dict.keys(self)

    def items(self):
        self._load(); return dict.items(self)

    def __contains__(self, k):
        self._load(); return dict.__contains__(self, k)
