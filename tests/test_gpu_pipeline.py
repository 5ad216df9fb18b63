"""Mixed YOLOv5 / SSD-MobileNet streams through the whole device pipeline (decode -> box filter -> NMS ->
gather -> tracker tick -> count-line) against the oracle chain run per stream on the host (the shape of
BASELINE config 4, at test size)."""
import numpy as np
import pytest
import torch

from oracle import detect as odet, deepsort as od, countline as oc
from deepdish_b200.scene import Scene
from tests.test_gpu_detect import COCO

pytestmark = pytest.mark.gpu

TRACK_LABELS = ["person", "bicycle", "car", "bus"]


def _yolo_head_from_scene(b, s, rng, na=600):
    """Encode one stream's scene boxes into a [na, 85] head (half-pixel offsets survive the int truncation)."""
    h = np.zeros((na, 85), np.float32)
    h[:, 0:2] = rng.uniform(0.1, 0.9, (na, 2)); h[:, 2:4] = rng.uniform(0.02, 0.1, (na, 2))
    h[:, 4] = rng.uniform(0.0, 0.2, na); h[:, 5:] = rng.uniform(0.0, 0.5, (na, 80))
    tlwh, conf, lab, _ = b.stream(s)
    rows = rng.choice(na, len(conf), replace=False)
    for r, t, c, l in zip(rows, tlwh, conf, lab):
        x, y, w, hh = t
        h[r, 0] = (x + 0.5 + (w + 0.5) / 2) / 640; h[r, 1] = (y + 0.5 + (hh + 0.5) / 2) / 480
        h[r, 2] = (w + 0.5) / 640; h[r, 3] = (hh + 0.5) / 480
        h[r, 4] = c; h[r, 5:] = 0.01
        h[r, 5 + [0, 1, 2, 5][l]] = 1.0                       # person, bicycle, car, bus
    return h


def test_mixed_yolo_ssd_pipeline_matches_oracle_chain():
    from deepdish_b200.batched import BatchedTracker
    from deepdish_b200.pipeline import DetectTrackPipeline, YoloFrontEnd, SsdFrontEnd
    SY, SS, D = 4, 3, 24
    S = SY + SS
    rng = np.random.default_rng(3)
    wanted = ["person", "bicycle", "car", "bus"]
    ssd_names = ["???"] + ["c%02d" % i for i in range(1, 91)]
    ssd_names[1], ssd_names[2], ssd_names[3], ssd_names[6] = "person", "bicycle", "car", "bus"
    anchors = odet.ssd_anchors()
    bt = BatchedTracker(S, TRACK_LABELS, max_tracks=96, max_dets=D, budget=30, max_age=20, n_chunks=2)
    pipe = DetectTrackPipeline(bt, [
        YoloFrontEnd(0, SY, COCO, wanted, TRACK_LABELS, ncap=256),
        SsdFrontEnd(SY, S, ssd_names, wanted, TRACK_LABELS, torch.from_numpy(anchors))])
    scene = Scene(SY, 14, D, n_labels=4, seed=8, clutter_mean=1.0)
    trk = [od.Trkr(od.Metric("cosine", 0.2, 30), 0.7, 20, 3) for _ in range(S)]
    cnt = [oc.LineCounter(oc.default_line(640, 480), TRACK_LABELS) for _ in range(S)]
    ssd_hot = [rng.choice(1917, 12, replace=False) for _ in range(SS)]
    for f in range(40):
        b = scene.step()
        head = np.stack([_yolo_head_from_scene(b, s, rng) for s in range(SY)])
        rb = rng.normal(0, 0.3, (SS, 1917, 4)).astype(np.float32)
        sc = rng.beta(0.4, 12, (SS, 1917, 91)).astype(np.float32)
        for s in range(SS):                       # persistent confident anchors -> persistent SSD tracks
            sc[s, ssd_hot[s], 1 + rng.choice([0, 1, 2, 5], 12)] = rng.uniform(0.55, 1.0, 12).astype(np.float32)
        feats = rng.normal(size=(S, D, 128)).astype(np.float32)
        ids = pipe.step([torch.from_numpy(head).cuda(), (torch.from_numpy(rb).cuda(), torch.from_numpy(sc).cuda())],
                        torch.from_numpy(feats).cuda()).cpu().numpy()
        pipe.check()
        got_count = pipe.det_count.cpu().numpy()
        got_tlwh = pipe.det_tlwh.cpu().numpy()
        for s in range(S):
            if s < SY:
                tlwh, cls, score, _ = odet.yolo_decode(head[s], 640, 480, COCO, wanted, 0.25)
                labels = [COCO[c] for c in cls]
                boxes = list(tlwh)
            else:
                ob, ocl, osc, _ = odet.tflite_detection_postprocess(rb[s - SY], sc[s - SY], anchors)
                t64, labels, score = odet.ssd_postprocess(ob, ocl, osc, 640, 480, ssd_names, wanted)
                boxes = [tuple(r) for r in t64]
            ib, kept = odet.box_filter(boxes, 640, 480)
            score = np.asarray(score, np.float32)[kept]
            labels = [labels[i] for i in kept]
            keep = odet.non_max_suppression(ib, 0.6, score) if len(ib) else []
            assert got_count[s] == len(keep), (f, s)
            np.testing.assert_array_equal(got_tlwh[s, :len(keep)], ib[keep].astype(float).reshape(-1, 4))
            use = ib[keep].astype(float).reshape(-1, 4)
            dets = [od.Det(use[k], labels[i], score[i], feats[s, k]) for k, i in enumerate(keep)]
            trk[s].trace = {}
            trk[s].predict(); trk[s].update(dets); cnt[s].step(trk[s])
            exp = [-1] * len(dets)
            for tid, d in trk[s].trace["match_ids"]:
                exp[d] = tid
            nxt = trk[s]._next_id - len(trk[s].trace["unmatched_detections"])
            for k, d in enumerate(trk[s].trace["unmatched_detections"]):
                exp[d] = nxt + k
            assert list(ids[s, :len(dets)]) == exp, (f, s)
    total = bt.total_counts.cpu().numpy()
    np.testing.assert_array_equal(total, sum(c.counts(TRACK_LABELS) for c in cnt))
    assert sum(t._next_id for t in trk) > 100


def test_frames_to_tracks_with_the_dummy_box_encoder():
    """The whole per-frame chain of deepdish run with `--encoder-model dummy`, on the device: YOLOv5 head decode + box
    filter + NMS -> extract_image_patch (16 x 8) + DummyImageEncoder on the camera frame -> tracker tick + count-line,
    against the oracle chain per stream (tools/yolov5.py, deepdish.py:946-955, preprocessing.py,
    tools/generate_detections.py:40-105,180-215, deep_sort)."""
    from oracle import patches as op
    from deepdish_b200.batched import BatchedTracker
    from deepdish_b200.pipeline import DetectTrackPipeline, YoloFrontEnd
    from deepdish_b200.tools import generate_detections as gd
    S, D, H, W = 5, 24, 480, 640
    rng = np.random.default_rng(13)
    wanted = ["person", "bicycle", "car", "bus"]
    yy, xx = np.mgrid[0:H, 0:W]
    frames = np.zeros((S, H, W, 3), np.uint8)
    for s in range(S):                      # smooth frames: the dummy feature of a box changes slowly as it moves
        for c in range(3):
            a, b_, ph = rng.uniform(0.004, 0.02, 2).tolist() + [rng.uniform(0, 6)]
            frames[s, :, :, c] = (127 + 100 * np.sin(a * xx + b_ * yy + ph) + rng.integers(-3, 4, (H, W))).clip(0, 255)
    frames_dev = torch.from_numpy(frames).cuda()
    enc = gd.create_box_encoder("dummy")
    bt = BatchedTracker(S, TRACK_LABELS, max_tracks=96, max_dets=D, budget=30, max_age=20)
    pipe = DetectTrackPipeline(bt, [YoloFrontEnd(0, S, COCO, wanted, TRACK_LABELS, ncap=256)])
    scene = Scene(S, 14, D, n_labels=4, seed=9, clutter_mean=1.0)
    trk = [od.Trkr(od.Metric("cosine", 0.2, 30), 0.7, 20, 3) for _ in range(S)]
    cnt = [oc.LineCounter(oc.default_line(640, 480), TRACK_LABELS) for _ in range(S)]
    got_feats = {}

    def features(tlwh, count):
        f, valid = enc.batch(frames_dev, tlwh, count)
        got_feats["f"], got_feats["valid"] = f, valid
        return f

    appearance_matches = 0
    for fidx in range(40):
        b = scene.step()
        head = np.stack([_yolo_head_from_scene(b, s, rng) for s in range(S)])
        ids = pipe.step([torch.from_numpy(head).cuda()], features).cpu().numpy()
        pipe.check()
        got_count = pipe.det_count.cpu().numpy()
        gf = got_feats["f"].cpu().numpy()
        for s in range(S):
            tlwh, cls, score, _ = odet.yolo_decode(head[s], 640, 480, COCO, wanted, 0.25)
            ib, kept = odet.box_filter(list(tlwh), 640, 480)
            score = np.asarray(score, np.float32)[kept]
            labels = [COCO[c] for c in cls[kept]]
            keep = odet.non_max_suppression(ib, 0.6, score) if len(ib) else []
            assert got_count[s] == len(keep), (fidx, s)
            boxes = ib[keep].reshape(-1, 4)
            patches = [op.extract_image_patch(frames[s], bx, (16, 8)) for bx in boxes]
            assert all(p is not None for p in patches)
            feats = op.dummy_encode(np.stack(patches)) if len(patches) else np.zeros((0, 128), np.float32)
            np.testing.assert_array_equal(gf[s, :len(keep)], feats, err_msg="features %d %d" % (fidx, s))   # bit-exact
            dets = [od.Det(boxes[k].astype(float), labels[i], score[i], feats[k]) for k, i in enumerate(keep)]
            trk[s].trace = {}
            trk[s].predict(); trk[s].update(dets); cnt[s].step(trk[s])
            appearance_matches += len(trk[s].trace["matches_a"])
            exp = [-1] * len(dets)
            for tid, d in trk[s].trace["match_ids"]:
                exp[d] = tid
            nxt = trk[s]._next_id - len(trk[s].trace["unmatched_detections"])
            for k, d in enumerate(trk[s].trace["unmatched_detections"]):
                exp[d] = nxt + k
            assert list(ids[s, :len(dets)]) == exp, (fidx, s)
    np.testing.assert_array_equal(bt.total_counts.cpu().numpy(), sum(c.counts(TRACK_LABELS) for c in cnt))
    assert appearance_matches > 200          # the cascade (appearance) stage did real work, not only the IoU stage
