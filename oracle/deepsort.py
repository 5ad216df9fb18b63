"""CPU restatement of the reference's DeepSORT update (numpy / scipy).  TEST INFRASTRUCTURE.

Every function cites the reference ``file:line`` it follows (paths relative to /root/reference).
The numerics deliberately go through the same numpy / scipy routines as the reference (``np.dot``,
``np.linalg.multi_dot``, ``scipy.linalg.cho_factor`` ...) so that this restatement is bit-identical
to the reference on the same inputs; ``tests/test_golden.py`` checks it against the fixtures ``oracle/make_golden.py``
generated from the unmodified reference in the build container (``tests/golden/``, which also travel to the GPU box).

Pinned: against the unmodified reference run in this container (fixtures in tests/golden/).
"""
import numpy as np
import scipy.linalg
from scipy.optimize import linear_sum_assignment

INFTY_COST = 1e5            # deep_sort/linear_assignment.py:8
CHI2INV95_4DOF = 9.4877     # deep_sort/kalman_filter.py:11-20 (4 dof)
CHI2INV95 = {1: 3.8415, 2: 5.9915, 3: 7.8147, 4: 9.4877, 5: 11.070, 6: 12.592, 7: 14.067,
             8: 15.507, 9: 16.919}

TENTATIVE, CONFIRMED, DELETED = 1, 2, 3   # deep_sort/track.py:15-17

_STD_POS = 1.0 / 20     # deep_sort/kalman_filter.py:52
_STD_VEL = 1.0 / 160    # deep_sort/kalman_filter.py:53

_F = np.eye(8)
for _i in range(4):
    _F[_i, 4 + _i] = 1.0    # deep_sort/kalman_filter.py:44-46 (dt = 1)
_H = np.eye(4, 8)           # deep_sort/kalman_filter.py:47


# ----------------------------------------------------------------------------- detection boxes
def tlwh_to_xyah(tlwh):
    """deep_sort/detection.py:43-50 -- centre x, centre y, aspect w/h, height."""
    out = np.array(tlwh, dtype=float)
    out[:2] += out[2:] / 2
    out[2] /= out[3]
    return out


def tlwh_to_tlbr(tlwh):
    """deep_sort/detection.py:35-41."""
    out = np.array(tlwh, dtype=float)
    out[2:] += out[:2]
    return out


class Det:
    """deep_sort/detection.py:29-33 -- tlwh f64[4], label, confidence float, feature f32[128]."""
    __slots__ = ("tlwh", "label", "confidence", "feature")

    def __init__(self, tlwh, label, confidence, feature):
        self.tlwh = np.asarray(tlwh, dtype=float)
        self.label = label
        self.confidence = float(confidence)
        self.feature = np.asarray(feature, dtype=np.float32)

    def to_xyah(self):
        return tlwh_to_xyah(self.tlwh)

    def to_tlbr(self):
        return tlwh_to_tlbr(self.tlwh)


# ----------------------------------------------------------------------------- Kalman filter
def kf_initiate(z):
    """deep_sort/kalman_filter.py:55-86."""
    z = np.asarray(z, dtype=float)
    mean = np.r_[z, np.zeros_like(z)]
    h = z[3]
    std = [2 * _STD_POS * h, 2 * _STD_POS * h, 1e-2, 2 * _STD_POS * h,
           10 * _STD_VEL * h, 10 * _STD_VEL * h, 1e-5, 10 * _STD_VEL * h]
    return mean, np.diag(np.square(std))


def kf_predict(mean, cov):
    """deep_sort/kalman_filter.py:88-123 -- Q uses h = mean[3] BEFORE propagation."""
    h = mean[3]
    std = np.r_[[_STD_POS * h, _STD_POS * h, 1e-2, _STD_POS * h],
                [_STD_VEL * h, _STD_VEL * h, 1e-5, _STD_VEL * h]]
    q = np.diag(np.square(std))
    new_mean = np.dot(_F, mean)
    new_cov = np.linalg.multi_dot((_F, cov, _F.T)) + q
    return new_mean, new_cov


def kf_project(mean, cov):
    """deep_sort/kalman_filter.py:125-152."""
    h = mean[3]
    std = [_STD_POS * h, _STD_POS * h, 1e-1, _STD_POS * h]
    r = np.diag(np.square(std))
    pm = np.dot(_H, mean)
    pc = np.linalg.multi_dot((_H, cov, _H.T))
    return pm, pc + r


def kf_update(mean, cov, z):
    """deep_sort/kalman_filter.py:154-186 -- Cholesky solve for the gain."""
    pm, pc = kf_project(mean, cov)
    cf, lower = scipy.linalg.cho_factor(pc, lower=True, check_finite=False)
    gain = scipy.linalg.cho_solve((cf, lower), np.dot(cov, _H.T).T, check_finite=False).T
    innov = z - pm
    new_mean = mean + np.dot(innov, gain.T)
    new_cov = cov - np.linalg.multi_dot((gain, pc, gain.T))
    return new_mean, new_cov


def kf_gating_distance(mean, cov, measurements, only_position=False):
    """deep_sort/kalman_filter.py:188-229 -- squared Mahalanobis distance per measurement."""
    pm, pc = kf_project(mean, cov)
    measurements = np.asarray(measurements, dtype=float)
    if only_position:
        pm, pc = pm[:2], pc[:2, :2]
        measurements = measurements[:, :2]
    chol = np.linalg.cholesky(pc)
    d = measurements - pm
    z = scipy.linalg.solve_triangular(chol, d.T, lower=True, check_finite=False, overwrite_b=True)
    return np.sum(z * z, axis=0)


# ----------------------------------------------------------------------------- appearance metric
def cosine_distance(a, b):
    """deep_sort/nn_matching.py:31-54 -- rows normalised in f32, 1 - a.b^T."""
    a = np.asarray(a) / np.linalg.norm(a, axis=1, keepdims=True)
    b = np.asarray(b) / np.linalg.norm(b, axis=1, keepdims=True)
    return 1. - np.dot(a, b.T)


def nn_cosine_distance(gallery, queries):
    """deep_sort/nn_matching.py:78-96 -- min over the gallery axis."""
    return cosine_distance(gallery, queries).min(axis=0)


def pdist_sq(a, b):
    """deep_sort/nn_matching.py:5-28."""
    a, b = np.asarray(a), np.asarray(b)
    if len(a) == 0 or len(b) == 0:
        return np.zeros((len(a), len(b)))
    a2, b2 = np.square(a).sum(axis=1), np.square(b).sum(axis=1)
    r2 = -2. * np.dot(a, b.T) + a2[:, None] + b2[None, :]
    return np.clip(r2, 0., float(np.inf))


def nn_euclidean_distance(gallery, queries):
    """deep_sort/nn_matching.py:57-75."""
    return np.maximum(0.0, pdist_sq(gallery, queries).min(axis=0))


class Metric:
    """deep_sort/nn_matching.py:99-177 -- per-target galleries with an optional budget."""

    def __init__(self, metric, matching_threshold, budget=None):
        if metric == "euclidean":
            self._fn = nn_euclidean_distance
        elif metric == "cosine":
            self._fn = nn_cosine_distance
        else:
            raise ValueError("Invalid metric; must be either 'euclidean' or 'cosine'")
        self.matching_threshold = matching_threshold
        self.budget = budget
        self.samples = {}

    def partial_fit(self, features, targets, active_targets):
        """nn_matching.py:137-154."""
        for f, t in zip(features, targets):
            lst = self.samples.setdefault(t, [])
            lst.append(f)
            if self.budget is not None:
                self.samples[t] = lst[-self.budget:]
        self.samples = {k: self.samples[k] for k in active_targets}

    def distance(self, features, targets):
        """nn_matching.py:156-177 -- [len(targets), len(features)] f64 filled from f32 rows."""
        out = np.zeros((len(targets), len(features)))
        for i, t in enumerate(targets):
            out[i, :] = self._fn(self.samples[t], features)
        return out


# ----------------------------------------------------------------------------- IoU
def iou(bbox, candidates):
    """deep_sort/iou_matching.py:7-39 -- tlwh boxes, no +1 convention."""
    b_tl, b_br = bbox[:2], bbox[:2] + bbox[2:]
    c_tl, c_br = candidates[:, :2], candidates[:, :2] + candidates[:, 2:]
    tl = np.c_[np.maximum(b_tl[0], c_tl[:, 0])[:, None], np.maximum(b_tl[1], c_tl[:, 1])[:, None]]
    br = np.c_[np.minimum(b_br[0], c_br[:, 0])[:, None], np.minimum(b_br[1], c_br[:, 1])[:, None]]
    wh = np.maximum(0., br - tl)
    inter = wh.prod(axis=1)
    return inter / (bbox[2:].prod() + candidates[:, 2:].prod(axis=1) - inter)


def track_tlwh(mean):
    """deep_sort/track.py:84-97."""
    r = mean[:4].copy()
    r[2] *= r[3]
    r[:2] -= r[2:] / 2
    return r


def track_tlbr(mean):
    """deep_sort/track.py:99-111."""
    r = track_tlwh(mean)
    r[2:] = r[:2] + r[2:]
    return r


def iou_cost(tracks, dets, track_indices, det_indices):
    """deep_sort/iou_matching.py:42-81 -- 1 - IoU; whole row INFTY when time_since_update > 1."""
    out = np.zeros((len(track_indices), len(det_indices)))
    for r, ti in enumerate(track_indices):
        if tracks[ti].time_since_update > 1:
            out[r, :] = INFTY_COST
            continue
        cand = np.asarray([dets[i].tlwh for i in det_indices])
        out[r, :] = 1. - iou(track_tlwh(tracks[ti].mean), cand)
    return out


# ----------------------------------------------------------------------------- assignment
def lsap_port(cost):
    """Pure-Python restatement of scipy 1.18.1 ``linear_sum_assignment`` (rectangular_lsap.cpp,
    modified Jonker-Volgenant / Crouse shortest augmenting path) INCLUDING its tie-breaking.

    Third-party dependency of deep_sort/linear_assignment.py:58 (absent from /root/reference);
    pinned version = the scipy installed in this image (1.18.1).  Rules that decide ties:
      * more rows than columns -> solve the transpose, report pairs sorted by original row;
      * rows are inserted in order; for each row the not-yet-scanned column list starts in
        REVERSE order (remaining[it] = nc-1-it) and a scanned column is removed by swap-with-last;
      * scanning in list order, the running minimum is replaced when strictly lower, or when
        equal and the column is unassigned.
    Returns (row_ind, col_ind) like scipy.  ``tests/test_hostemu.py::test_lsap_device_code_and_oracle_port_match_scipy``
    checks it against scipy on tie-heavy matrices.
    """
    cost = np.asarray(cost, dtype=float)
    nr, nc = cost.shape
    if nr == 0 or nc == 0:
        return np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)
    transpose = nc < nr
    if transpose:
        cost = cost.T.copy()
        nr, nc = nc, nr
    inf = float("inf")
    c = cost.tolist()
    u = [0.0] * nr
    v = [0.0] * nc
    col4row = [-1] * nr
    row4col = [-1] * nc
    path = [-1] * nc
    for cur in range(nr):
        spc = [inf] * nc
        sr = [False] * nr
        sc = [False] * nc
        remaining = [nc - 1 - it for it in range(nc)]
        nrem = nc
        min_val = 0.0
        i = cur
        sink = -1
        while sink == -1:
            index = -1
            lowest = inf
            sr[i] = True
            ci = c[i]
            ui = u[i]
            for it in range(nrem):
                j = remaining[it]
                r = min_val + ci[j] - ui - v[j]
                if r < spc[j]:
                    path[j] = i
                    spc[j] = r
                if spc[j] < lowest or (spc[j] == lowest and row4col[j] == -1):
                    lowest = spc[j]
                    index = it
            min_val = lowest
            if min_val == inf:
                raise ValueError("cost matrix is infeasible")
            j = remaining[index]
            if row4col[j] == -1:
                sink = j
            else:
                i = row4col[j]
            sc[j] = True
            nrem -= 1
            remaining[index] = remaining[nrem]
        u[cur] += min_val
        for i2 in range(nr):
            if sr[i2] and i2 != cur:
                u[i2] += min_val - spc[col4row[i2]]
        for j2 in range(nc):
            if sc[j2]:
                v[j2] -= min_val - spc[j2]
        j = sink
        while True:
            i2 = path[j]
            row4col[j] = i2
            col4row[i2], j = j, col4row[i2]
            if i2 == cur:
                break
    if transpose:
        order = np.argsort(np.asarray(col4row), kind="stable")
        return np.asarray(col4row, dtype=np.int64)[order], order.astype(np.int64)
    return np.arange(nr, dtype=np.int64), np.asarray(col4row, dtype=np.int64)


def cpython_set_difference_order(n, matched):
    """Iteration order of ``list(set(range(n)) - set(matched))`` on CPython 3.12 (setobject.c),
    restated without using ``set`` -- the order deep_sort/linear_assignment.py:140 hands to
    deep_sort/tracker.py:120-122.  ``matched`` = distinct ints in [0, n).

    * ``(n >> 2) > len(matched)`` : set_copy_and_difference -> ascending survivors.
    * otherwise survivors are inserted in ascending order into a fresh table of 8 slots
      (slot = v & mask; on collision the next 9 slots are probed when ``i + 9 <= mask``, then
      ``perturb >>= 5; i = (5 i + 1 + perturb) & mask``); when ``fill * 5 >= mask * 3`` after an
      insert the table is rebuilt at the smallest power of two > 4 * used, re-inserting in old
      slot order; the result is read out in slot order.
    """
    matched = set(int(m) for m in matched)
    survivors = [k for k in range(n) if k not in matched]
    if (n >> 2) > len(matched):
        return survivors
    mask = 7
    table = [None] * 8
    used = 0

    def insert_clean(tbl, msk, key):
        perturb = key
        i = key & msk
        while True:
            if tbl[i] is None:
                tbl[i] = key
                return
            if i + 9 <= msk:
                for j in range(1, 10):
                    if tbl[i + j] is None:
                        tbl[i + j] = key
                        return
            perturb >>= 5
            i = (i * 5 + 1 + perturb) & msk

    for key in survivors:
        insert_clean(table, mask, key)     # no equal keys, no dummies: add == clean insert
        used += 1
        if used * 5 >= mask * 3:
            newsize = 8
            while newsize <= used * 4:
                newsize <<= 1
            newtable = [None] * newsize
            for k in table:
                if k is not None:
                    insert_clean(newtable, newsize - 1, k)
            table, mask = newtable, newsize - 1
    return [k for k in table if k is not None]


def min_cost_matching(cost_fn, max_distance, tracks, dets, track_indices, det_indices, lsap=None):
    """deep_sort/linear_assignment.py:11-75 -- clip, solve, split; ORDER of the outputs matters."""
    if len(det_indices) == 0 or len(track_indices) == 0:
        return [], track_indices, det_indices
    cost = cost_fn(tracks, dets, track_indices, det_indices)
    cost[cost > max_distance] = max_distance + 1e-5
    rows, cols = (lsap or linear_sum_assignment)(cost)
    rows, cols = list(rows), list(cols)
    matches, un_t, un_d = [], [], []
    colset, rowset = set(cols), set(rows)
    for c, d in enumerate(det_indices):
        if c not in colset:
            un_d.append(d)
    for r, t in enumerate(track_indices):
        if r not in rowset:
            un_t.append(t)
    for r, c in zip(rows, cols):
        t, d = track_indices[r], det_indices[c]
        if cost[r, c] > max_distance:
            un_t.append(t)
            un_d.append(d)
        else:
            matches.append((t, d))
    return matches, un_t, un_d


def matching_cascade(cost_fn, max_distance, depth, tracks, dets, track_indices, lsap=None,
                     set_order=None):
    """deep_sort/linear_assignment.py:78-141."""
    un_d = list(range(len(dets)))
    matches = []
    for level in range(depth):
        if len(un_d) == 0:
            break
        lvl = [k for k in track_indices if tracks[k].time_since_update == 1 + level]
        if len(lvl) == 0:
            continue
        m, _, un_d = min_cost_matching(cost_fn, max_distance, tracks, dets, lvl, un_d, lsap)
        matches += m
    if set_order is None:
        un_t = list(set(track_indices) - set(k for k, _ in matches))
    else:   # emulated CPython order; valid because confirmed indices are 0..n-1
        assert list(track_indices) == list(range(len(track_indices)))
        un_t = set_order(len(track_indices), [k for k, _ in matches])
    return matches, un_t, un_d


def gate_cost_matrix(cost, tracks, dets, track_indices, det_indices):
    """deep_sort/linear_assignment.py:144-190 (4-dof gate)."""
    meas = np.asarray([dets[i].to_xyah() for i in det_indices])
    for r, ti in enumerate(track_indices):
        g = kf_gating_distance(tracks[ti].mean, tracks[ti].covariance, meas)
        cost[r, g > CHI2INV95_4DOF] = INFTY_COST
    return cost


# ----------------------------------------------------------------------------- track + tracker
class Trk:
    """deep_sort/track.py:67-196."""

    def __init__(self, mean, cov, track_id, n_init, max_age, det):
        self.mean, self.covariance = mean, cov
        self.track_id = track_id
        self.hits, self.age, self.time_since_update = 1, 1, 0
        self.state = TENTATIVE
        self.features = [det.feature]
        self.labels = [det.label]
        self.dist = {det.label: [det.confidence]}
        self._n_init, self._max_age = n_init, max_age

    def predict(self):
        """track.py:113-125."""
        self.mean, self.covariance = kf_predict(self.mean, self.covariance)
        self.age += 1
        self.time_since_update += 1

    def update(self, det):
        """track.py:127-152."""
        self.mean, self.covariance = kf_update(self.mean, self.covariance, det.to_xyah())
        self.features.append(det.feature)
        self.hits += 1
        self.time_since_update = 0
        if self.state == TENTATIVE and self.hits >= self._n_init:
            self.state = CONFIRMED
        self.labels.append(det.label)
        self.dist.setdefault(det.label, []).append(det.confidence)

    def mark_missed(self):
        """track.py:190-196."""
        if self.state == TENTATIVE:
            self.state = DELETED
        elif self.time_since_update > self._max_age:
            self.state = DELETED

    def is_confirmed(self):
        return self.state == CONFIRMED

    def is_deleted(self):
        return self.state == DELETED

    def to_tlwh(self):
        return track_tlwh(self.mean)

    def to_tlbr(self):
        return track_tlbr(self.mean)

    def get_label(self, return_confidence=False):
        """track.py:154-188 -- Dirichlet-expected label vote with the motorbike/bicycle rule."""
        if not self.labels:
            return (None, 0) if return_confidence else None
        stats = [(lbl, len(s), np.average(s)) for lbl, s in self.dist.items()]
        alphas = np.array([a for _, _, a in stats])
        cnt = np.array([c for _, c, _ in stats])
        names = [l for l, _, _ in stats]
        ranked = list(reversed(sorted(zip((alphas + cnt) / (cnt.sum() + alphas.sum()), names))))
        pick = ranked[0][1]
        if len(ranked) > 1 and ranked[0][1] == 'motorbike' and ranked[1][1] == 'bicycle':
            pick = 'motorbike' if ranked[0][0] > ranked[1][0] * 4 else 'bicycle'
        if return_confidence:
            return pick, np.average(self.dist[pick])
        return pick


class Trkr:
    """deep_sort/tracker.py:40-138.  ``trace`` (optional dict) receives per-update internals
    (matches, unmatched lists in their original order) for the parity tests."""

    def __init__(self, metric, max_iou_distance=0.7, max_age=30, n_init=3, lsap=None,
                 set_order=None):
        self.metric = metric
        self.max_iou_distance, self.max_age, self.n_init = max_iou_distance, max_age, n_init
        self.tracks, self.deleted_tracks = [], []
        self._next_id = 1
        self._lsap, self._set_order = lsap, set_order
        self.trace = None

    def predict(self):
        """tracker.py:51-57."""
        for t in self.tracks:
            t.predict()

    def _gated_metric(self, tracks, dets, track_indices, det_indices):
        """tracker.py:97-105."""
        feats = np.array([dets[i].feature for i in det_indices])
        targets = np.array([tracks[i].track_id for i in track_indices])
        cost = self.metric.distance(feats, targets)
        cost = gate_cost_matrix(cost, tracks, dets, track_indices, det_indices)
        if self.trace is not None:
            self.trace.setdefault("gated_costs", []).append(
                ([int(t) for t in targets], [int(d) for d in det_indices], cost.copy()))
        return cost

    def _match(self, dets):
        """tracker.py:95-133."""
        confirmed = [i for i, t in enumerate(self.tracks) if t.is_confirmed()]
        unconfirmed = [i for i, t in enumerate(self.tracks) if not t.is_confirmed()]
        m_a, un_t_a, un_d = matching_cascade(
            self._gated_metric, self.metric.matching_threshold, self.max_age, self.tracks, dets,
            confirmed, self._lsap, self._set_order)
        iou_cand = unconfirmed + [k for k in un_t_a if self.tracks[k].time_since_update == 1]
        un_t_a = [k for k in un_t_a if self.tracks[k].time_since_update != 1]
        m_b, un_t_b, un_d = min_cost_matching(
            iou_cost, self.max_iou_distance, self.tracks, dets, iou_cand, un_d, self._lsap)
        if self.trace is not None:
            self.trace.update(matches_a=list(m_a), matches_b=list(m_b), iou_rows=list(iou_cand))
        return m_a + m_b, list(set(un_t_a + list(un_t_b))), un_d

    def update(self, dets):
        """tracker.py:59-93."""
        matches, un_t, un_d = self._match(dets)
        if self.trace is not None:
            self.trace.update(matches=list(matches), unmatched_tracks=list(un_t),
                              unmatched_detections=list(un_d),
                              match_ids=[(self.tracks[t].track_id, d) for t, d in matches])
        for t, d in matches:
            self.tracks[t].update(dets[d])
        for t in un_t:
            self.tracks[t].mark_missed()
        for d in un_d:
            mean, cov = kf_initiate(dets[d].to_xyah())
            self.tracks.append(Trk(mean, cov, self._next_id, self.n_init, self.max_age, dets[d]))
            self._next_id += 1
        self.deleted_tracks = [t for t in self.tracks if t.is_deleted()]
        self.tracks = [t for t in self.tracks if not t.is_deleted()]
        active = [t.track_id for t in self.tracks if t.is_confirmed()]
        feats, targets = [], []
        for t in self.tracks:
            if not t.is_confirmed():
                continue
            feats += t.features
            targets += [t.track_id for _ in t.features]
            t.features = []
        self.metric.partial_fit(np.asarray(feats), np.asarray(targets), active)
