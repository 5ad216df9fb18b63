"""The drop-in deep_sort API (single-stream Tracker facade + per-function operators) against the
oracle, written the way a test-suite of the reference would read."""
import numpy as np
import pytest

from oracle import deepsort as od
from deepdish_b200.scene import Scene
from tests import goldens
from tests.goldens import LABELS3

pytestmark = pytest.mark.gpu


def test_install_as_deep_sort_and_track_lifecycle():
    import deepdish_b200
    deepdish_b200.install_as_deep_sort(box_encoder=True, framerecords=True)
    from tools import generate_detections as gdet
    from deepdish.framerecords import FrameRecords
    assert gdet.create_box_encoder("constant")(np.zeros((32, 32, 3), np.uint8), [np.array([2, 2, 8, 16])]).shape == (1, 128)
    assert FrameRecords({0: "person"}).process_boxes(0, np.zeros((0, 4), np.int64), [], np.zeros(0)) == ([], [], [])
    from deep_sort import nn_matching, preprocessing
    from deep_sort.detection import Detection
    from deep_sort.tracker import Tracker
    from tools.intersection import intersection, any_intersection
    with pytest.raises(ValueError):
        nn_matching.NearestNeighborDistanceMetric("manhattan", 0.2)
    metric = nn_matching.NearestNeighborDistanceMetric("cosine", 0.2, 20)
    tracker = Tracker(metric, max_iou_distance=0.7, max_age=30)
    om = od.Metric("cosine", 0.2, 20)
    ot = od.Trkr(om, 0.7, 30, 3)
    sc = Scene(1, 10, 12, n_labels=3, seed=77)
    for f in range(60):
        tlwh, conf, lab, feat = sc.step().stream(0)
        dets = [Detection(tlwh[i], LABELS3[lab[i]], conf[i], feat[i]) for i in range(len(conf))]
        odets = [od.Det(tlwh[i], LABELS3[lab[i]], conf[i], feat[i]) for i in range(len(conf))]
        tracker.predict(); tracker.update(dets)
        ot.predict(); ot.update(odets)
        assert [t.track_id for t in tracker.tracks] == [t.track_id for t in ot.tracks]
        assert [t.state for t in tracker.tracks] == [t.state for t in ot.tracks]
        assert [t.is_confirmed() for t in tracker.tracks] == [t.is_confirmed() for t in ot.tracks]
        assert [t.time_since_update for t in tracker.tracks] == [t.time_since_update for t in ot.tracks]
        assert [t.track_id for t in tracker.deleted_tracks] == [t.track_id for t in ot.deleted_tracks]
        assert [t.get_label() for t in tracker.tracks] == [t.get_label() for t in ot.tracks]
        assert tracker._next_id == ot._next_id
        for a, b in zip(tracker.tracks, ot.tracks):
            np.testing.assert_allclose(a.mean, b.mean, rtol=1e-4, atol=1e-9)
            np.testing.assert_allclose(a.to_tlbr(), b.to_tlbr(), rtol=1e-4, atol=1e-9)
    assert sorted(metric.samples.keys()) == sorted(om.samples.keys())
    for k in om.samples:
        ref = np.asarray(om.samples[k]); ref = ref / np.linalg.norm(ref, axis=1, keepdims=True)
        np.testing.assert_allclose(np.asarray(metric.samples[k]), ref, rtol=1e-5, atol=1e-7)
    # intersection asserts of tools/intersection.py:35-57 through the mirror
    from tests.test_oracle_intersection import GOLDEN_SEGMENTS, GOLDEN_POLYLINES
    for p, pr, q, qs, exp in GOLDEN_SEGMENTS:
        assert intersection(p, pr, q, qs) == exp
    for p, q, pts, exp in GOLDEN_POLYLINES:
        assert any_intersection(p, q, pts) == exp
    assert preprocessing.non_max_suppression(np.zeros((0, 4)), 0.6, np.zeros(0)) == []


def test_per_function_api():
    from deepdish_b200.deep_sort import kalman_filter, nn_matching, iou_matching, linear_assignment, preprocessing
    from deepdish_b200.deep_sort.detection import Detection
    rng = np.random.default_rng(5)
    kf = kalman_filter.KalmanFilter()
    z = np.array([320., 240., 0.5, 80.])
    mean, cov = kf.initiate(z)
    em, ec = od.kf_initiate(z)
    np.testing.assert_array_equal(mean, em); np.testing.assert_array_equal(cov, ec)
    mean, cov = kf.predict(mean, cov); em, ec = od.kf_predict(em, ec)
    np.testing.assert_allclose(cov, ec, rtol=1e-12)
    mean, cov = kf.update(mean, cov, z + 1); em, ec = od.kf_update(em, ec, z + 1)
    np.testing.assert_allclose(mean, em, rtol=1e-9); np.testing.assert_allclose(cov, ec, rtol=1e-7, atol=1e-13)
    meas = z + rng.normal(0, 3, (7, 4)) * [1, 1, 0.01, 1]
    np.testing.assert_allclose(kf.gating_distance(mean, cov, meas), od.kf_gating_distance(em, ec, meas), rtol=1e-7)
    assert kalman_filter.chi2inv95[4] == 9.4877
    # metric
    m = nn_matching.NearestNeighborDistanceMetric("cosine", 0.2, 3)
    o = od.Metric("cosine", 0.2, 3)
    f = rng.normal(size=(9, 128)).astype(np.float32)
    t = np.array([1, 1, 2, 1, 2, 1, 3, 3, 1])
    m.partial_fit(f, t, [1, 2]); o.partial_fit(f, t, [1, 2])
    assert sorted(m.samples) == [1, 2] and len(m.samples[1]) == 3
    q = rng.normal(size=(4, 128)).astype(np.float32)
    np.testing.assert_allclose(m.distance(q, [1, 2]), o.distance(q, [1, 2]), rtol=1e-4, atol=2e-6)
    # NMS + IoU + min_cost_matching ordering
    g = goldens.load("nms.npz")
    n = int(g["counts"][3])
    assert preprocessing.non_max_suppression(g["boxes"][3, :n].astype(np.int64), float(g["thr"][3]), g["scores"][3, :n]) \
        == list(g["keep"][3, :g["nkeep"][3]])
    np.testing.assert_allclose(iou_matching.iou(np.array([10., 10, 20, 40]), np.array([[12., 8, 20, 40], [200, 200, 5, 5]])),
                               od.iou(np.array([10., 10, 20, 40]), np.array([[12., 8, 20, 40], [200, 200, 5, 5]])))

    class T:
        def __init__(s, tsu): s.time_since_update = tsu
    cost = rng.random((6, 9)); cost[cost > 0.5] = 5.0

    def metric_fn(tracks, dets, ti, di):
        return cost[np.ix_(ti, di)].copy()
    got = linear_assignment.min_cost_matching(metric_fn, 0.5, [T(1)] * 6, [None] * 9, list(range(6)), list(range(9)))
    exp = od.min_cost_matching(metric_fn, 0.5, [T(1)] * 6, [None] * 9, list(range(6)), list(range(9)))
    assert got == exp
    assert linear_assignment.INFTY_COST == 1e5
    assert Detection([1, 2, 3, 4], "person", 0.5, np.zeros(128)).to_xyah().tolist() == [2.5, 4.0, 0.75, 4.0]


def test_detector_adapter_facades_vs_reference_fixture():
    """tools.yolov5.YOLOV5 / tools.ssd_mobilenet.SSD_MOBILENET mirrors: detect_image's own output (float
    boxes, label names, scores, order) against what the unmodified reference adapters returned."""
    from PIL import Image
    from deepdish_b200.tools.yolov5 import YOLOV5
    from deepdish_b200.tools.ssd_mobilenet import SSD_MOBILENET
    from oracle import detect as odet
    g = goldens.load("yolo.npz")
    names, wanted = list(g["names"]), list(g["wanted"])
    img = Image.new("RGB", (640, 480))
    for f in range(g["head"].shape[0]):
        det = YOLOV5(wanted_labels=wanted, labels=names, head_fn=lambda im, f=f: g["head"][f:f + 1])
        boxes, labels, scores = det.detect_image(img)
        np.testing.assert_array_equal(np.array(boxes, np.float32).reshape(-1, 4), g["tlwh%d" % f])
        assert [names.index(l) for l in labels] == list(g["cls%d" % f])
        np.testing.assert_array_equal(np.array(scores, np.float32), g["score%d" % f])
    # SSD: the op's decode is unpinned (third party); feed heads whose decoded top boxes are known via the oracle
    rng = np.random.default_rng(2)
    anchors = odet.ssd_anchors()
    names = ["???"] + ["c%02d" % i for i in range(1, 91)]
    names[1], names[3] = "person", "car"
    rb = rng.normal(0, 0.5, (4, 1917, 4)).astype(np.float32)
    sc = rng.beta(0.4, 12, (4, 1917, 91)).astype(np.float32)
    for b in range(4):
        sc[b, rng.choice(1917, 10, replace=False), 1 + rng.choice([0, 2], 10)] = rng.uniform(0.55, 1, 10).astype(np.float32)
    det = SSD_MOBILENET(wanted_labels=["person", "car"], labels=names, anchors=anchors)
    got = det.detect_heads(rb, sc, (640, 480))
    for b in range(4):
        ob, oc, os_, _ = odet.tflite_detection_postprocess(rb[b], sc[b], anchors)
        tlwh, labels, scores = odet.ssd_postprocess(ob, oc, os_, 640, 480, names, ["person", "car"])
        assert got[b][1] == labels
        np.testing.assert_array_equal(np.array(got[b][2], np.float32), scores)
        np.testing.assert_allclose(np.array(got[b][0]).reshape(-1, 4), tlwh, rtol=1e-5, atol=1e-3)
