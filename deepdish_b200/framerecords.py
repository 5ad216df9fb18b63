"""``deepdish.framerecords`` mirror (reference deepdish/framerecords.py): the per-frame records that merge CVAT
annotations with detections around the tracker.  ``Pipeline`` calls, every frame,
``process_boxes`` before the encoder (deepdish.py:1001), ``process_detections`` before ``Tracker.update`` (:1017) and
``tracker.tracks = process_tracking(frame, tracker)`` after it (:1047).

Host-side bookkeeping only -- but ``process_tracking`` reaches into the tracker: it force-updates a missed track with the
annotation's detection (``t.update(tracker.kf, det)``, state Confirmed, time_since_update 0) and drops duplicate tracks
by assigning ``tracker.tracks``.  With ``deepdish_b200.deep_sort.tracker.Tracker`` both are written back to the device
state (``Tracker._flush``).  The CVAT XML import / export of the reference (``xml_output``, ``indent``) is file I/O and
is not mirrored.
"""
from copy import copy

from .deep_sort.track import TrackState


class FrameRecord:
    """framerecords.py:12-16."""

    def __init__(self, tlbr, label_id, order=None):
        self.tlbr, self.label_id, self.order = tlbr, label_id, order


class AnnotationRecord(FrameRecord):
    """framerecords.py:18-28: one annotated box of one annotation track in one frame."""

    def __init__(self, annot_track_id, lbl, det_lbl_id, tlbr, outside, occluded, keyframe, z_order, order=None):
        super().__init__(tlbr, det_lbl_id, order=order)
        self.annotation_track_id, self.annotation_label = annot_track_id, lbl
        self.is_outside, self.is_occluded, self.is_keyframe, self.z_order = outside, occluded, keyframe, z_order
        self.tentative_matches = {}
        self.score = 1.0


class TentativeRecord(FrameRecord):
    """framerecords.py:30-33: a detector box."""

    def __init__(self, tlbr, label_id, score, order=None):
        super().__init__(tlbr, label_id, order=order)
        self.score = score


def overlap(a, b):
    """framerecords.py:36-41: intersection area over the smaller of the two box areas."""
    ax1, ay1, ax2, ay2 = list(a.tlbr)
    bx1, by1, bx2, by2 = list(b.tlbr)
    inter = max(0, min(ax2, bx2) - max(ax1, bx1)) * max(0, min(ay2, by2) - max(ay1, by1))
    return inter / min(abs(ax2 - ax1) * abs(ay2 - ay1), abs(bx2 - bx1) * abs(by2 - by1))


class FrameRecords:
    def __init__(self, detector_id_to_labelname, overlap_threshold=0.9, override_tentative_detections=True,
                 minimum_track_frames=3):
        self.frames = {}
        self.labels = {}
        self.detector_id_to_labelname = detector_id_to_labelname
        self.detector_labelname_to_id = {v: k for k, v in detector_id_to_labelname.items()}
        self.overlap_threshold = overlap_threshold
        self.override_tentative_detections = override_tentative_detections
        self.minimum_track_frames = minimum_track_frames

    def add_annotation_label_info(self, annotlabelname, detectorlabelid, annotlabelcolor):
        self.labels[annotlabelname] = {'detector_id': detectorlabelid, 'color': annotlabelcolor}

    def add_annotated_track(self, frame, annot_track_id, lbl, pts, outside, occluded, keyframe, z_order):
        rec = AnnotationRecord(annot_track_id, lbl, self.labels[lbl]['detector_id'], pts, outside, occluded, keyframe,
                               z_order)
        self.frames.setdefault(frame, []).append(rec)

    def process_boxes(self, frame, boxes_in, labelnames_in, scores_in):
        """framerecords.py:63-123: detector boxes (tlwh) + this frame's annotations -> the boxes the encoder and
        the tracker see, in the order [annotations that overlap a detection, detections without annotation,
        annotations without detection]."""
        free = []                                             # detections not yet claimed by an annotation
        for i, (tlwh, name, score) in enumerate(zip(boxes_in, labelnames_in, scores_in)):
            tlbr = copy(tlwh)
            tlbr[2:] = tlbr[:2] + tlbr[2:]
            free.append(TentativeRecord(tlbr, self.detector_labelname_to_id[name], score, order=i))
        claimed, lonely, unknown = [], [], []
        for rec in self.frames.setdefault(frame, []):
            if not isinstance(rec, AnnotationRecord):
                continue
            hit = False
            for k, det in enumerate(free):
                if overlap(rec, det) >= self.overlap_threshold and (rec.label_id == det.label_id or rec.label_id is None):
                    rec.tentative_matches[frame] = copy(det)
                    claimed.append(rec)
                    del free[k]
                    hit = True
                    break
            if not hit and rec.label_id is not None:
                lonely.append(rec)                            # enters the tracker as a detection with score 1.0
            elif rec.label_id is None:
                unknown.append(rec)                           # kept on record only (also when it claimed a detection)
        result = claimed + free + lonely
        boxes_out, labels_out, scores_out = [], [], []
        for i, rec in enumerate(result):
            rec.order = i
            tlwh = copy(rec.tlbr)
            tlwh[2:] = tlwh[2:] - tlwh[:2]
            boxes_out.append(tlwh)
            labels_out.append(self.detector_id_to_labelname[rec.label_id])
            scores_out.append(rec.score)
        self.frames[frame] = result + unknown
        return boxes_out, labels_out, scores_out

    def process_detections(self, frame, detections):
        """framerecords.py:125-129."""
        for det, rec in zip(detections, self.frames[frame]):
            rec.detection = det
            det.record = rec
        return detections

    def process_tracking(self, frame, tracker, tracks=None):
        """framerecords.py:131-184: tie tracks to annotation tracks through the records of their detections; a track
        that follows exactly one annotation track and missed this frame is updated with the annotation's detection
        and confirmed; of several tracks following the same annotation track only the one with the most recorded
        detections survives."""
        if tracks is None:
            tracks = tracker.tracks
        followers = {}                                        # annotation track id -> [(tracker id, #recorded dets)]
        for t in tracks:
            with_rec = [d for d in t.detections if hasattr(d, 'record')]
            ann_ids = {d.record.annotation_track_id for d in with_rec if isinstance(d.record, AnnotationRecord)}
            if len(ann_ids) == 1:
                i = ann_ids.pop()
                r = next((r for r in self.frames[frame]
                          if isinstance(r, AnnotationRecord) and r.annotation_track_id == i), None)
                if r is not None:
                    followers.setdefault(i, []).append((t.track_id, len(with_rec)))
                    if t.time_since_update > 0:
                        t.update(tracker.kf, r.detection)
                        t.state = TrackState.Confirmed
                        t.time_since_update = 0
            for d in with_rec:
                d.record.track = t
        drop = set()
        for entries in followers.values():
            most = max(entries, key=lambda e: e[1])[1]
            drop.update(tid for tid, n in entries if n < most)
        return [t for t in tracks if t.track_id not in drop]
