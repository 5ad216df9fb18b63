"""``tools.tflite`` mirror (reference tools/tflite.py + tools/tflite_object_detector.py): the ``TFLITE`` detector adapter
(EfficientDet-style models whose graph ends in the detection post-process op).  ``ObjectDetector._postprocess``
(tflite_object_detector.py:234-295) and ``TFLITE.detect_image``'s filtering (tflite.py:26-41) run in the CUDA kernel
``k_tflite_post``.

The CNN is out of scope: instead of an interpreter the adapter takes ``outputs_fn``: a callable
``ndarray image [H,W,3] -> (boxes [n,4] (ymin,xmin,ymax,xmax normalised), classes [n], scores [n], count)`` -- the four
tensors ``ObjectDetector.detect`` reads (tflite_object_detector.py:213-218).
"""
from typing import List, NamedTuple

import numpy as np
import torch

from .. import ops


class ObjectDetectorOptions(NamedTuple):
    """tools/tflite_object_detector.py:42-61."""
    enable_edgetpu: bool = False
    label_allow_list: List[str] = None
    label_deny_list: List[str] = None
    max_results: int = -1
    num_threads: int = 1
    score_threshold: float = 0.0


class TFLITE:
    def __init__(self, wanted_labels=None, model_file=None, label_file=None, num_threads=None, edgetpu=False,
                 libedgetpu=None, score_threshold=0.5, outputs_fn=None, label_list=None, input_size=(320, 320),
                 options=None):
        self.opts = options or ObjectDetectorOptions(num_threads=num_threads or 1, score_threshold=score_threshold,
                                                     enable_edgetpu=edgetpu)
        self.use_edgetpu = edgetpu
        self.num_threads = num_threads or 1
        if wanted_labels is None:
            wanted_labels = ['person']
        self.wanted_labels = wanted_labels
        if label_list is None:
            if label_file is None:
                raise ValueError("TFLITE needs `label_list` or `label_file` (the model metadata is out of scope)")
            with open(label_file) as f:
                label_list = [line.strip() for line in f.readlines()]
        self.label_list = list(label_list)
        self.labels = {i + 1: self.label_list[i] for i in range(0, len(self.label_list))}     # tflite.py:23
        self.width, self.height = input_size
        self.outputs_fn = outputs_fn
        o = self.opts
        ok = [1] * len(self.label_list)
        for i, n in enumerate(self.label_list):
            if o.label_deny_list is not None and n in o.label_deny_list:
                ok[i] = 0
            if o.label_allow_list is not None and n not in o.label_allow_list:
                ok[i] = 0
        self._ok = torch.tensor(ok, dtype=torch.uint8, device="cuda")
        self._wanted = torch.tensor([1 if n in self.wanted_labels else 0 for n in self.label_list], dtype=torch.uint8,
                                    device="cuda")

    def detect_outputs(self, boxes, classes, scores, count, img_size):
        """Batched form: boxes [b,n,4], classes [b,n], scores [b,n], count [b] for frames of one (w, h) size.
        Returns per frame (boxes [[left, top, w, h], ...] ints, label names, scores) like detect_image."""
        bx = ops._dev(boxes, torch.float32)
        out = ops.tflite_postprocess(bx, ops._dev(classes, torch.float32), ops._dev(scores, torch.float32),
                                     ops._dev(count, torch.int32), self._ok, self._wanted, self.opts.score_threshold,
                                     img_size, self.opts.max_results)
        if int(out["flags"].max()) != 0:
            raise IndexError("class id outside the label list")
        cnt = out["count"].cpu().numpy()
        tl, sc, lb = out["tlwh"].cpu().numpy(), out["score"].cpu().numpy(), out["label"].cpu().numpy()
        res = []
        for f in range(bx.shape[0]):
            n = int(cnt[f])
            res.append(([[int(v) for v in b] for b in tl[f, :n]], [self.label_list[int(c)] for c in lb[f, :n]],
                        list(sc[f, :n])))
        return res

    def detect_image(self, img):
        """tools/tflite.py:26-41."""
        if self.outputs_fn is None:
            raise RuntimeError("TFLITE.detect_image needs outputs_fn (the CNN is out of scope of deepdish_b200)")
        arr = np.array(img)[..., :3]
        boxes, classes, scores, count = self.outputs_fn(arr)
        boxes = np.asarray(boxes, np.float32).reshape(1, -1, 4)
        return self.detect_outputs(boxes, np.asarray(classes, np.float32).reshape(1, -1),
                                   np.asarray(scores, np.float32).reshape(1, -1), np.asarray([int(count)], np.int32),
                                   (arr.shape[1], arr.shape[0]))[0]
