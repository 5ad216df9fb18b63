"""FrameRecords against outputs of the unmodified reference (tests/golden/framerecords.json.gz, made by
oracle/make_golden.py golden_framerecords): process_boxes on the CPU; the whole per-frame sequence
process_boxes -> process_detections -> Tracker.predict / update -> tracker.tracks = process_tracking on the GPU, where
process_tracking's host edits (force-updated, confirmed tracks; dropped duplicates) must reach the device state."""
import gzip
import json
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "framerecords.json.gz")


def _load():
    with gzip.open(GOLDEN, "rt") as fh:
        return json.load(fh)


def _records(g):
    from deepdish_b200.framerecords import FrameRecords
    names = {int(k): v for k, v in g["names"].items()}
    fr = FrameRecords(names)
    for n in names.values():
        fr.add_annotation_label_info(n, fr.detector_labelname_to_id[n], "#000000")
    fr.add_annotation_label_info("unicorn", None, "#ffffff")
    return fr


def _annotate(fr, g, f):
    for a in g["annotations"]:
        if a[0] == f:
            fr.add_annotated_track(a[0], a[1], a[2], np.array(a[3]), False, False, True, 0)


def test_process_boxes_matches_reference():
    g = _load()
    fr = _records(g)
    n_extra = 0
    for f, fx in enumerate(g["frames"]):
        _annotate(fr, g, f)
        b, l, s = fr.process_boxes(f, np.array(fx["boxes"], dtype=np.int64).reshape(-1, 4), fx["labels"],
                                   np.array(fx["scores"], dtype=np.float32))
        assert [[float(v) for v in x] for x in b] == fx["out_boxes"], f
        assert list(l) == fx["out_labels"] and [float(x) for x in s] == fx["out_scores"], f
        n_extra += len(b) - len(fx["boxes"])
    assert n_extra > 10            # annotations without a detection entered as detections


@pytest.mark.gpu
def test_frame_loop_with_forced_updates_and_dropped_tracks():
    from deepdish_b200.deep_sort import nn_matching
    from deepdish_b200.deep_sort.detection import Detection
    from deepdish_b200.deep_sort.tracker import Tracker
    g = _load()
    fr = _records(g)
    trk = Tracker(nn_matching.NearestNeighborDistanceMetric("cosine", 0.2, 50), max_iou_distance=0.7, max_age=30, n_init=3)
    forced = 0
    for f, fx in enumerate(g["frames"]):
        _annotate(fr, g, f)
        b, l, s = fr.process_boxes(f, np.array(fx["boxes"], dtype=np.int64).reshape(-1, 4), fx["labels"],
                                   np.array(fx["scores"], dtype=np.float32))
        dets = [Detection(bx, lb, sc, np.array(ft, np.float32)) for bx, lb, sc, ft in zip(b, l, s, fx["feats"])]
        dets = fr.process_detections(f, dets)
        trk.predict()
        trk.update(dets)
        before = [(t.track_id, t.time_since_update) for t in trk.tracks]
        trk.tracks = fr.process_tracking(f, trk)
        forced += sum(1 for (tid, tsu), t in zip(before, trk.tracks) if t.track_id == tid and tsu > 0 and t.time_since_update == 0)
        assert [t.track_id for t in trk.tracks] == fx["track_ids"], f
        assert [int(t.state) for t in trk.tracks] == fx["states"], f
        assert [int(t.time_since_update) for t in trk.tracks] == fx["tsu"], f
        assert [int(t.hits) for t in trk.tracks] == fx["hits"], f
        assert trk._next_id == fx["next_id"], f
        if trk.tracks:
            np.testing.assert_allclose(np.stack([t.mean for t in trk.tracks]), np.array(fx["means"]), rtol=1e-4, atol=1e-9)
    assert forced >= 2
