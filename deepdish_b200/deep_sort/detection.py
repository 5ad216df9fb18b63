"""``deep_sort.detection`` mirror (reference deep_sort/detection.py:5-50).  A Detection is a small host
value object; the batched tracker packs lists of them into padded device tensors."""
import numpy as np


class Detection(object):
    """Bounding box detection in a single image: ``tlwh`` f64[4], ``label``, ``confidence`` float,
    ``feature`` f32[d] (this fork's signature, detection.py:29)."""

    def __init__(self, tlwh, label, confidence, feature):
        self.tlwh = np.asarray(tlwh, dtype=float)
        self.label = label
        self.confidence = float(confidence)
        self.feature = np.asarray(feature, dtype=np.float32)

    def to_tlbr(self):
        """(min x, min y, max x, max y) -- detection.py:35-41."""
        box = self.tlwh.copy()
        box[2:] += box[:2]
        return box

    def to_xyah(self):
        """(centre x, centre y, aspect w/h, height) -- detection.py:43-50."""
        box = self.tlwh.copy()
        box[:2] += box[2:] / 2
        box[2] /= box[3]
        return box
