"""``tools.yolov5`` mirror (reference tools/yolov5.py): the ``YOLOV5`` detector adapter whose
post-processing (yolov5.py:115-146) runs in the CUDA kernel ``k_yolo_decode``.

The CNN itself is out of scope, so instead of a TFLite interpreter the adapter takes ``head_fn``: a callable
``PIL.Image -> ndarray | tensor [1, N, 5+C]`` producing the exported model's normalised head (the tensor
``interpreter.get_tensor(output_details[0]['index'])`` returns at yolov5.py:109), optionally quantised
(uint8 + ``quantization=(scale, zero_point)``, yolov5.py:115-118).
"""
import numpy as np
import torch

from .. import ops


class YOLOV5:
    def __init__(self, wanted_labels=None, head_fn=None, labels=None, label_file=None, score_threshold=0.25,
                 input_size=(640, 640), quantization=None, num_threads=None, edgetpu=False):
        if wanted_labels is None:
            wanted_labels = ['person']
        self.wanted_labels = wanted_labels
        self.score_threshold = score_threshold
        if labels is None:
            if label_file is None:
                raise ValueError("YOLOV5 needs `labels` (index -> name) or `label_file`")
            with open(label_file) as f:                                   # yolov5.py:91-95
                labels = {i: line.strip() for i, line in enumerate(f.readlines())}
        self.labels = dict(labels) if isinstance(labels, dict) else {i: n for i, n in enumerate(labels)}
        self.width, self.height = input_size
        self.use_edgetpu, self.num_threads = edgetpu, num_threads
        self.int8 = quantization is not None
        self.quantization = quantization
        self.head_fn = head_fn
        names = [self.labels[i] for i in range(len(self.labels))]
        self._mask = torch.tensor([1 if n in self.wanted_labels else 0 for n in names], dtype=torch.uint8,
                                  device="cuda")

    def detect_heads(self, heads, img_sizes):
        """Batched form: heads [B, N, 5+C] (f32, or uint8 with quantization), one (w, h) image size for all.
        Returns per frame (boxes tlwh f32 lists, label names, scores) like detect_image."""
        head = ops._dev(heads, torch.uint8 if self.int8 else torch.float32)
        ncap = head.shape[1]
        out = ops.yolo_decode(head, self._mask, self.score_threshold, img_sizes, (0, 0), ncap=min(ncap, 4096),
                              quant=self.quantization)
        if int(out["flags"].max()) != 0:
            raise RuntimeError("more than %d detections in a frame" % min(ncap, 4096))
        res = []
        cnt = out["count"].cpu().numpy()
        tl, sc, cl = out["tlwh"].cpu().numpy(), out["score"].cpu().numpy(), out["cls"].cpu().numpy()
        for f in range(head.shape[0]):
            n = int(cnt[f])
            res.append(([list(b) for b in tl[f, :n].astype(np.float32)], [self.labels[int(c)] for c in cl[f, :n]],
                        list(sc[f, :n])))
        return res

    def detect_image(self, img):
        """tools/yolov5.py:97-146 -> (boxes [[x, y, w, h], ...], label names, scores), ascending anchor order."""
        if self.head_fn is None:
            raise RuntimeError("YOLOV5.detect_image needs head_fn (the CNN is out of scope of deepdish_b200)")
        head = self.head_fn(img)
        head = head if isinstance(head, torch.Tensor) else np.asarray(head)
        return self.detect_heads(head.reshape(1, head.shape[-2], head.shape[-1]), img.size)[0]
