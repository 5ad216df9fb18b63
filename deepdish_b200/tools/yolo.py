"""``tools.yolo`` mirror (reference tools/yolo.py): the Keras YOLOv3 detector adapter ``YOLO`` whose post-processing --
decode_netout, correct_yolo_boxes, do_nms, get_boxes and the tail of detect_image (yolo.py:48-153,207-237) -- runs in the
CUDA kernel ``k_yolo3_post`` (dd_yolo3_decode), quirks included: the returned boxes are transposed (x = box[1],
y = box[0], yolo.py:222-225), a box with two labels above the threshold comes back twice, results are in reversed
get_boxes order.

The CNN is out of scope, so instead of ``keras.models.load_model`` the adapter takes ``model``: an object with
``predict(image_data [1, H, W, 3] float32) -> [map0, map1, map2]`` returning the three raw output maps
``[1, g, g, 255]`` (what ``self.model.predict`` returns at yolo.py:200).
"""
import numpy as np
import torch

from .. import ops

COCO_NAMES = ["person", "bicycle", "car", "motorbike", "aeroplane", "bus", "train", "truck", "boat", "traffic light",
              "fire hydrant", "stop sign", "parking meter", "bench", "bird", "cat", "dog", "horse", "sheep", "cow", "elephant",
              "bear", "zebra", "giraffe", "backpack", "umbrella", "handbag", "tie", "suitcase", "frisbee", "skis", "snowboard",
              "sports ball", "kite", "baseball bat", "baseball glove", "skateboard", "surfboard", "tennis racket", "bottle",
              "wine glass", "cup", "fork", "knife", "spoon", "bowl", "banana", "apple", "sandwich", "orange", "broccoli",
              "carrot", "hot dog", "pizza", "donut", "cake", "chair", "sofa", "pottedplant", "bed", "diningtable", "toilet",
              "tvmonitor", "laptop", "mouse", "remote", "keyboard", "cell phone", "microwave", "oven", "toaster", "sink",
              "refrigerator", "book", "clock", "vase", "scissors", "teddy bear", "hair drier", "toothbrush"]


def letterbox_image(image, size):
    """yolo.py:155-166 (host-side image preparation for the CNN; PIL)."""
    from PIL import Image
    image_w, image_h = image.size
    w, h = size
    new_w = int(image_w * min(w * 1.0 / image_w, h * 1.0 / image_h))
    new_h = int(image_h * min(w * 1.0 / image_w, h * 1.0 / image_h))
    resized = image.resize((new_w, new_h), Image.BICUBIC)
    boxed = Image.new('RGB', size, (128, 128, 128))
    boxed.paste(resized, ((w - new_w) // 2, (h - new_h) // 2))
    return boxed


class YOLO(object):
    def __init__(self, wanted_labels=None, model=None, label_file=None, num_threads=None, score_threshold=0.5,
                 model_image_size=(416, 416), class_names=None):
        self.use_edgetpu = False
        self.num_threads = 1
        self.model = model
        self.anchors = [[116, 90, 156, 198, 373, 326], [30, 61, 62, 45, 59, 119], [10, 13, 16, 30, 33, 23]]   # yolo.py:166
        self.class_names = list(class_names) if class_names is not None else list(COCO_NAMES)
        self.score_threshold = score_threshold
        self.iou = 0.5
        self.model_image_size = tuple(model_image_size)
        self.height, self.width = self.model_image_size
        self.is_fixed_size = self.model_image_size != (None, None)
        if wanted_labels is None:
            wanted_labels = ['person']
        self.wanted_labels = wanted_labels
        self.labels = dict(enumerate(self.class_names))
        self._mask = torch.tensor([1 if n in self.wanted_labels else 0 for n in self.class_names], dtype=torch.uint8,
                                  device="cuda")

    def detect_maps(self, maps, image_size, ncap=256):
        """Batched form: maps = three arrays / tensors [B, g, g, 3 * (5 + C)] of raw network outputs, one (w, h) camera
        image size for all.  Returns per frame (boxes [[x, y, w, h], ...] ints, label names, scores) like detect_image."""
        dev = [ops._dev(np.asarray(m) if not isinstance(m, torch.Tensor) else m, torch.float32) for m in maps]
        input_w, input_h = self.model_image_size                       # yolo.py:201 (sic)
        out = ops.yolo3_decode(dev, self.anchors, self._mask, self.score_threshold, 0.5, image_size, (input_w, input_h),
                               ncap=ncap)
        flags = int(out["flags"].max())
        if flags & 2:
            raise RuntimeError("more than %d results (or more than 128 boxes above the threshold) in a frame" % ncap)
        if flags & 64:
            raise ZeroDivisionError("float division by zero")          # bbox_iou of two zero-area boxes (yolo.py:116)
        cnt = out["count"].cpu().numpy()
        bx, sc, lb = out["box"].cpu().numpy(), out["score"].cpu().numpy(), out["label"].cpu().numpy()
        res = []
        for f in range(len(cnt)):
            n = int(cnt[f])
            res.append(([[int(v) for v in b] for b in bx[f, :n]], [self.class_names[int(c)] for c in lb[f, :n]],
                        list(sc[f, :n])))
        return res

    def detect_image(self, image):
        """yolo.py:186-237 -> (boxes [[x, y, w, h], ...], label names, scores)."""
        if self.model is None:
            raise RuntimeError("YOLO.detect_image needs `model` (the CNN is out of scope of deepdish_b200)")
        if self.is_fixed_size:
            boxed = letterbox_image(image, tuple(reversed(self.model_image_size)))
        else:
            boxed = letterbox_image(image, (image.width - (image.width % 32), image.height - (image.height % 32)))
        data = np.expand_dims(np.array(boxed, dtype='float32') / 255., 0)
        yhat = self.model.predict(data)
        return self.detect_maps([np.asarray(y) for y in yhat], image.size)[0]

    def close_session(self):
        pass
