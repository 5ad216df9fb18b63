"""CPU restatement of the count-line crossing test and the per-label counters.  TEST INFRASTRUCTURE.

Follows tools/intersection.py:4-30 and deepdish.py:1035-1114,1303-1312 (paths relative to
/root/reference).  Pinned: the six known-answer asserts of tools/intersection.py:35-57 are replayed
in tests/test_oracle_intersection.py; the counters are pinned against the reference's own
``Pipeline.process_results`` driven in the build container (fixtures in tests/golden/).
"""
import sys

import numpy as np

_EPS = sys.float_info.epsilon


def _cross2(a, b):
    # numpy's 2-D np.cross (tools/intersection.py:8,10,22): a0*b1 - a1*b0, two roundings + one
    return a[0] * b[1] - a[1] * b[0]


def intersection(p, pr, q, qs):
    """tools/intersection.py:4-24 -- do segments p->pr and q->qs intersect?"""
    r = pr - p
    s = qs - q
    rxs = _cross2(r, s)
    qmp = q - p
    qpxr = _cross2(qmp, r)
    if abs(rxs) < _EPS:
        if abs(qpxr) < _EPS:          # collinear: overlap of the projected interval with [0, 1]
            rdrr = r / np.dot(r, r)
            t0 = np.dot(qmp, rdrr)
            t1 = t0 + np.dot(s, rdrr)
            if t0 > t1:
                t0, t1 = t1, t0
            return not (t1 < 0 or t0 > 1)
        return False                  # parallel
    t = _cross2(qmp, s) / rxs
    u = qpxr / rxs
    return bool(0.0 <= t and t <= 1.0 and 0.0 <= u and u <= 1.0)


def any_intersection(p1, q1, pts):
    """tools/intersection.py:26-30 -- any consecutive pair of the polyline crosses p1->q1."""
    for a, b in zip(pts, pts[1:]):
        if intersection(p1, q1, a, b):
            return True
    return False


def default_line(width, height):
    """deepdish.py:739-744 -- vertical mid-line, integer-truncated then float."""
    return np.array([[width / 2, 0], [width / 2, height]], dtype=int).astype(float)


class LineCounter:
    """deepdish.py:1035-1114 (process_results, counting part) + :1303-1312 (check_deleted_track).

    Reference quirk kept: ``delcounts`` is overwritten per deleted track (deepdish.py:1041-1044), so
    only the LAST deleted track of a frame can contribute to ``delcount``.
    """

    def __init__(self, line, labels):
        self.line = np.asarray(line, dtype=float)
        self.db = {}
        self.pos = {l: 0 for l in labels}
        self.neg = {l: 0 for l in labels}
        self.int = {l: 0 for l in labels}
        self.dele = {l: 0 for l in labels}
        self.events = []           # per step: list of (track_id, label, direction) for the tests

    def _check_deleted(self, trk):
        out = {}
        i = trk.track_id
        if i in self.db and len(self.db[i]) > 1:
            if any_intersection(self.line[0], self.line[1], np.array(self.db[i])):
                l = trk.get_label()
                out[l] = out.get(l, 0) + 1
            self.db[i] = []
        return out

    def step(self, tracker):
        delcounts = {}
        for trk in tracker.deleted_tracks:
            if trk.is_deleted():
                delcounts = self._check_deleted(trk)
        hits = []
        p1, q1 = self.line[0], self.line[1]
        for trk in tracker.tracks:
            lbl = trk.get_label()
            if not trk.is_confirmed() or trk.time_since_update > 1:
                continue
            path = self.db.setdefault(trk.track_id, [])
            bb = trk.to_tlbr()
            path.append(np.array([(bb[0] + bb[2]) / 2.0, bb[3]]))
            if len(path) > 1:
                p2, q2 = np.array(path[-1]), np.array(path[-2])
                cp = _cross2(q1 - p1, q2 - p2)
                if intersection(p1, q1, p2, q2):
                    hits.append((trk.track_id, lbl, cp))
        ev = []
        for tid, lbl, cp in hits:
            if cp >= 0:
                self.pos[lbl] += 1
            else:
                self.neg[lbl] += 1
            self.int[lbl] += 1
            ev.append((tid, lbl, 1 if cp >= 0 else -1))
        for lbl, d in delcounts.items():
            self.dele[lbl] += d
        self.events.append(ev)

    def counts(self, labels):
        """[C, 4] int64: pos, neg, int, del per label."""
        return np.array([[self.pos[l], self.neg[l], self.int[l], self.dele[l]] for l in labels],
                        dtype=np.int64)
