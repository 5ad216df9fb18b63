"""Shared parity harness: drive the oracle (one Trkr per stream) next to a batched implementation
(CUDA BatchedTracker or the host emulation) on the same Scene and compare every tick."""
import numpy as np

from oracle import deepsort as od, countline as oc

LABELS3 = ["person", "bicycle", "car"]


class OracleStreams:
    def __init__(self, n_streams, labels, budget=100, max_age=30, n_init=3, max_cos=0.2, max_iou=0.7,
                 line=None, port=False):
        self.labels = list(labels)
        kw = dict(lsap=od.lsap_port, set_order=od.cpython_set_difference_order) if port else {}
        self.trk = [od.Trkr(od.Metric("cosine", max_cos, budget), max_iou, max_age, n_init, **kw)
                    for _ in range(n_streams)]
        if line is None:
            lines = [oc.default_line(640, 480)] * n_streams
        else:
            line = np.asarray(line, float)
            lines = [line.reshape(2, 2)] * n_streams if line.size == 4 else [l.reshape(2, 2) for l in line]
        self.cnt = [oc.LineCounter(l, self.labels) for l in lines]
        self.det_ids = None

    def step(self, batch, streams=None):
        """batch: deepdish_b200.scene.SceneBatch on the host.  Returns per-stream det->track id lists."""
        out = []
        for s, (t, c) in enumerate(zip(self.trk, self.cnt)):
            tlwh, conf, lab, feat = batch.stream(s if streams is None else streams[s])
            dets = [od.Det(tlwh[i], self.labels[lab[i]], conf[i], feat[i]) for i in range(len(conf))]
            t.trace = {}
            t.predict()
            t.update(dets)
            c.step(t)
            ids = [-1] * len(dets)
            for tid, d in t.trace["match_ids"]:
                ids[d] = tid
            nxt = t._next_id - len(t.trace["unmatched_detections"])
            for k, d in enumerate(t.trace["unmatched_detections"]):
                ids[d] = nxt + k
            out.append(ids)
        self.det_ids = out
        return out


def compare_stream(o_trk, o_cnt, view, s, labels, rtol=1e-4, gallery=True, gallery_rows=None):
    """Compare oracle tracker/counter of one stream with the SoA state `view` (dict of numpy arrays).
    gallery_rows(s, slot) -> [len,128] float32 (optional): the implementation's gallery of a slot, oldest row first,
    compared with metric.samples (the implementation stores the rows unit-normalised)."""
    n = int(view["n_tracks"][s])
    slots = view["order"][s, :n]
    assert n == len(o_trk.tracks), (s, n, len(o_trk.tracks))
    assert [int(x) for x in view["track_id"][s, slots]] == [t.track_id for t in o_trk.tracks]
    assert [int(x) for x in view["state"][s, slots]] == [t.state for t in o_trk.tracks]
    assert [int(x) for x in view["hits"][s, slots]] == [t.hits for t in o_trk.tracks]
    assert [int(x) for x in view["age"][s, slots]] == [t.age for t in o_trk.tracks]
    assert [int(x) for x in view["tsu"][s, slots]] == [t.time_since_update for t in o_trk.tracks]
    nd = int(view["n_deleted"][s])
    dsl = view["deleted"][s, :nd]
    assert [int(x) for x in view["track_id"][s, dsl]] == [t.track_id for t in o_trk.deleted_tracks]
    assert int(view["next_id"][s]) == o_trk._next_id
    if n:
        om = np.stack([t.mean for t in o_trk.tracks])
        ocv = np.stack([t.covariance for t in o_trk.tracks])
        np.testing.assert_allclose(view["mean"][s, slots], om, rtol=rtol, atol=1e-9)
        np.testing.assert_allclose(view["cov"][s, slots], ocv, rtol=rtol, atol=1e-12)
    if gallery:
        for k, t in enumerate(o_trk.tracks):
            sl = slots[k]
            if t.is_confirmed():
                g = np.asarray(o_trk.metric.samples[t.track_id])
                assert int(view["gal_len"][s, sl]) == len(g)
                if gallery_rows is not None:
                    got = np.asarray(gallery_rows(s, int(sl)))
                    unit = g / np.sqrt(np.sum(g * g, axis=1, dtype=np.float32))[:, None]
                    np.testing.assert_allclose(got, unit, rtol=2e-6, atol=1e-7)
    if o_cnt is not None:
        np.testing.assert_array_equal(view["counts"][s], o_cnt.counts(labels))


def compare_costs(o_trk, view_pre, view_post, s, rtol=1e-4, atol=2e-6):
    """Gated appearance costs of the last update: oracle (per cascade level, before clipping) against the
    CUDA gate bits + f32 min-cosine costs.  view_pre: track_id/order/n_tracks BEFORE the update (slot of a
    track id); view_post: gate / cost arrays after it.  Returns the number of gate-passing pairs checked."""
    n = int(view_pre["n_tracks"][s])
    slot_of = {int(view_pre["track_id"][s, sl]): int(sl) for sl in view_pre["order"][s, :n]}
    checked = 0
    for tids, dets, cost in o_trk.trace.get("gated_costs", []):
        for r, tid in enumerate(tids):
            sl = slot_of[tid]
            for c, d in enumerate(dets):
                bit = (int(view_post["gate"][s, sl, d >> 5]) >> (d & 31)) & 1
                if cost[r, c] >= 1e5:
                    assert bit == 0, (tid, d)
                else:
                    assert bit == 1, (tid, d)
                    got = float(view_post["cost"][s, sl, d])
                    assert abs(got - cost[r, c]) <= atol + rtol * abs(cost[r, c]), (tid, d, got, cost[r, c])
                    checked += 1
    return checked
