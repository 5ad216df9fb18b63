"""``tools.ssd_mobilenet`` mirror (reference tools/ssd_mobilenet.py): the ``SSD_MOBILENET`` detector adapter
whose post-processing (the TFLite_Detection_PostProcess op restated + ssd_mobilenet.py:59-150,198-213) runs
in the CUDA kernel ``k_ssd_decode``.

The CNN is out of scope: ``head_fn`` maps a ``PIL.Image`` to the raw SSD head ``(raw_boxes [1917,4],
raw_scores [1917,91])`` that feeds the post-processing op inside the reference's .tflite file.
"""
import numpy as np
import torch

from .. import ops


class SSD_MOBILENET:
    def __init__(self, wanted_labels=None, head_fn=None, labels=None, label_file=None, anchors=None,
                 score_threshold=0.5, input_size=(300, 300), num_threads=None, edgetpu=False):
        if wanted_labels is None:
            wanted_labels = ['person']
        self.wanted_labels = wanted_labels
        self.score_threshold = score_threshold
        if labels is None:
            if label_file is None:
                raise ValueError("SSD_MOBILENET needs `labels` or `label_file`")
            with open(label_file, 'r') as f:                              # ssd_mobilenet.py:45-47
                labels = {i: line.strip() for i, line in enumerate(f.readlines())}
        self.labels = dict(labels) if isinstance(labels, dict) else {i: n for i, n in enumerate(labels)}
        self.width, self.height = input_size
        self.use_edgetpu, self.num_threads = edgetpu, num_threads
        self.head_fn = head_fn
        if anchors is None:
            raise ValueError("SSD_MOBILENET needs the model's anchor boxes [A,4] (ycenter, xcenter, h, w)")
        self._anchors = ops._dev(anchors, torch.float32)
        n = len(self.labels)
        c2l = [(c + 1) if (c + 1 < n and self.labels[c + 1] in self.wanted_labels) else -1 for c in range(n - 1)]
        self._c2l = torch.tensor(c2l, dtype=torch.int32, device="cuda")

    def detect_heads(self, raw_boxes, raw_scores, img_size):
        rb, rs = ops._dev(raw_boxes, torch.float32), ops._dev(raw_scores, torch.float32)
        out = ops.ssd_decode(rb, rs, self._anchors, self._c2l, self.score_threshold, 0.5, img_size, (0, 0))
        res = []
        cnt = out["count"].cpu().numpy()
        tl, sc, lb = out["tlwh"].cpu().numpy(), out["score"].cpu().numpy(), out["label"].cpu().numpy()
        for f in range(rb.shape[0]):
            n = int(cnt[f])
            res.append(([list(b) for b in tl[f, :n]], [self.labels[int(l)] for l in lb[f, :n]], list(sc[f, :n])))
        return res

    def detect_image(self, img):
        """tools/ssd_mobilenet.py:198-213 -> (boxes [[x, y, w, h], ...], label names, scores)."""
        if self.head_fn is None:
            raise RuntimeError("SSD_MOBILENET.detect_image needs head_fn (the CNN is out of scope of deepdish_b200)")
        rb, rs = self.head_fn(img)
        rb, rs = np.asarray(rb, np.float32), np.asarray(rs, np.float32)
        return self.detect_heads(rb.reshape(1, -1, 4), rs.reshape(1, rb.reshape(-1, 4).shape[0], -1), img.size)[0]
