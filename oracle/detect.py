"""CPU restatement of the detector post-processing: YOLOv5 head decode, SSD-MobileNet post-process,
the pre-NMS box filter and NMS.  TEST INFRASTRUCTURE.

Follows tools/yolov5.py:115-146, tools/ssd_mobilenet.py:59-150,198-213, deepdish.py:941-960 and
deep_sort/preprocessing.py:6-73 (paths relative to /root/reference).

Pinned: YOLO decode, the box filter (the reference's own loop, driven through Pipeline.detect_objects by
oracle/refload.RefBoxFilter), NMS and the SSD *post*-processing against the unmodified reference
(fixtures in tests/golden/).  PARITY UNPINNED: ``tflite_detection_postprocess`` (the 1917-anchor decode
and first NMS) is the third-party TFLite custom op ``TFLite_Detection_PostProcess``
(tensorflow/lite/kernels/detection_postprocess.cc; pinned runtime tflite_runtime 2.5.0.post1 in the
reference's Dockerfile:32) whose model files are absent from /root/reference; it is restated here from
its published algorithm (fast-NMS mode, standard SSD-MobileNet-v1 attributes) and nothing in the
reference pins its output.
"""
import numpy as np


# ----------------------------------------------------------------------------- YOLOv5
def yolo_decode(head, img_w, img_h, class_names, wanted, thr=0.25):
    """tools/yolov5.py:120-146.  head [N, 5+C] f32 (normalised xywh, obj, cls...).
    Returns (tlwh f32 [K,4], class_idx [K], score f32 [K], anchor_idx [K]) in ascending anchor order."""
    x = np.array(head, dtype=np.float32, copy=True)
    boxes = np.empty_like(x[:, :4])
    boxes[:, 0] = x[:, 0] - x[:, 2] / 2
    boxes[:, 1] = x[:, 1] - x[:, 3] / 2
    boxes[:, 2] = x[:, 0] + x[:, 2] / 2
    boxes[:, 3] = x[:, 1] + x[:, 3] / 2
    x[:, 5:] *= x[:, 4:5]
    best = np.argmax(x[:, 5:], axis=-1)                     # first maximum wins
    conf = np.take_along_axis(x, (best + 5)[:, None], axis=-1)[:, 0]
    keep = np.where(conf >= np.float32(thr))[0]
    b = boxes[keep]
    b *= np.array([img_w, img_h, img_w, img_h])             # f32 * int -> rounded to f32 (:131)
    wanted = set(wanted)
    sel = np.array([class_names[int(c)] in wanted for c in best[keep]], dtype=bool)
    b, k = b[sel], keep[sel]
    tlwh = b.copy()
    tlwh[:, 2] = b[:, 2] - b[:, 0]
    tlwh[:, 3] = b[:, 3] - b[:, 1]
    return tlwh, best[k].astype(np.int64), conf[k], k


def yolo_dequant(q, scale, zero_point):
    """tools/yolov5.py:115-118 -- int8/uint8 head -> f32."""
    return (q.astype(np.float32) - zero_point) * scale


# ----------------------------------------------------------------------------- box filter
def box_filter(boxes, frame_w, frame_h):
    """deepdish.py:946-955 (motion test excluded).  boxes: sequence of (x,y,w,h) numpy scalars.
    Returns (int boxes [K,4] int64, kept input indices)."""
    arr = np.asarray(boxes)
    out, idx = [], []
    if arr.size and np.any(np.isnan(arr)):
        return np.zeros((0, 4), dtype=np.int64), np.zeros(0, dtype=np.int64)
    for n, (x, y, w, h) in enumerate(boxes):
        x, y = int(np.clip(x, 0, frame_w)), int(np.clip(y, 0, frame_h))
        w, h = int(np.clip(w, 0, frame_w - x)), int(np.clip(h, 0, frame_h - y))
        if w * h > 0.9 * frame_w * frame_h:
            continue
        out.append((x, y, w, h))
        idx.append(n)
    return np.array(out, dtype=np.int64).reshape(-1, 4), np.array(idx, dtype=np.int64)


# ----------------------------------------------------------------------------- NMS
def non_max_suppression(boxes, max_overlap, scores=None):
    """deep_sort/preprocessing.py:6-73 -- greedy, overlap = inter / area(other), +1 px convention.
    Returns indices in pick order (descending score).  Equal scores: the reference's np.argsort (preprocessing.py:50)
    is unstable and build-dependent -- numpy's scalar introsort (the reference's ARM targets) is an insertion sort,
    i.e. stable, for n <= 16, numpy's x86 SIMD argsort is not.  The oracle pins the scalar-path behaviour with an
    explicitly stable sort (identical to the reference whenever scores are unique, and for ties at n <= 16 on the
    scalar path: fixture nms_ties.npz); ties at n > 16 are implementation-defined in the reference."""
    if len(boxes) == 0:
        return []
    b = np.asarray(boxes).astype(float)
    x1, y1 = b[:, 0], b[:, 1]
    x2, y2 = b[:, 2] + b[:, 0], b[:, 3] + b[:, 1]
    area = (x2 - x1 + 1) * (y2 - y1 + 1)
    order = np.argsort(scores, kind="stable") if scores is not None else np.argsort(y2, kind="stable")
    pick = []
    while len(order) > 0:
        i = order[-1]
        rest = order[:-1]
        pick.append(int(i))
        w = np.maximum(0, np.minimum(x2[i], x2[rest]) - np.maximum(x1[i], x1[rest]) + 1)
        h = np.maximum(0, np.minimum(y2[i], y2[rest]) - np.maximum(y1[i], y1[rest]) + 1)
        order = rest[~((w * h) / area[rest] > max_overlap)]
    return pick


# ----------------------------------------------------------------------------- SSD-MobileNet
def ssd_anchors():
    """Standard SSD-MobileNet-v1 300x300 anchor set (1917 = 19^2*3 + (10^2+5^2+3^2+2^2+1)*6), rows
    (ycenter, xcenter, h, w) f32.  Anchors are an INPUT to both oracle and kernel, so the exact recipe
    does not affect parity."""
    out = []
    grids = [19, 10, 5, 3, 2, 1]
    scales = [0.2 + (0.95 - 0.2) * i / 5 for i in range(6)] + [1.0]
    for k, g in enumerate(grids):
        if k == 0:
            specs = [(0.1, 1.0), (scales[0], 2.0), (scales[0], 0.5)]
        else:
            specs = [(scales[k], 1.0), (scales[k], 2.0), (scales[k], 0.5), (scales[k], 3.0),
                     (scales[k], 1.0 / 3), (np.sqrt(scales[k] * scales[k + 1]), 1.0)]
        for y in range(g):
            for x in range(g):
                for s, ar in specs:
                    out.append(((y + 0.5) / g, (x + 0.5) / g, s / np.sqrt(ar), s * np.sqrt(ar)))
    return np.array(out, dtype=np.float32)


def _iou_f32(a, b):
    """detection_postprocess.cc ComputeIntersectionOverUnion, f32, boxes (ymin,xmin,ymax,xmax)."""
    f = np.float32
    area_a = (a[2] - a[0]) * (a[3] - a[1])
    area_b = (b[2] - b[0]) * (b[3] - b[1])
    if area_a <= 0 or area_b <= 0:
        return f(0)
    ih = max(min(a[2], b[2]) - max(a[0], b[0]), f(0))
    iw = max(min(a[3], b[3]) - max(a[1], b[1]), f(0))
    inter = f(ih * iw)
    return f(inter / f(f(area_a + area_b) - inter))


def tflite_detection_postprocess(raw_boxes, raw_scores, anchors, max_det=10, score_thr=1e-8,
                                 iou_thr=0.6, scales=(10.0, 10.0, 5.0, 5.0)):
    """PARITY UNPINNED restatement of TFLite_Detection_PostProcess (fast NMS, 1 class/detection).
    raw_boxes [A,4] (ty,tx,th,tw), raw_scores [A,1+K] (col 0 = background), anchors [A,4].
    Returns the op's four outputs: boxes [max_det,4] (ymin,xmin,ymax,xmax), classes, scores, count."""
    f = np.float32
    rb, an = np.asarray(raw_boxes, f), np.asarray(anchors, f)
    yc = rb[:, 0] / f(scales[0]) * an[:, 2] + an[:, 0]
    xc = rb[:, 1] / f(scales[1]) * an[:, 3] + an[:, 1]
    # exp: correctly rounded f32 by declaration (f64 exp, one rounding) -- the CUDA kernel's dd_expf does the same
    hh = f(0.5) * np.exp((rb[:, 2] / f(scales[2])).astype(np.float64)).astype(f) * an[:, 2]
    hw = f(0.5) * np.exp((rb[:, 3] / f(scales[3])).astype(np.float64)).astype(f) * an[:, 3]
    dec = np.stack([yc - hh, xc - hw, yc + hh, xc + hw], axis=1).astype(f)
    cls_scores = np.asarray(raw_scores, f)[:, 1:]
    best = np.argmax(cls_scores, axis=1)
    best_score = cls_scores[np.arange(len(best)), best]
    cand = np.where(best_score >= f(score_thr))[0]
    cand = cand[np.argsort(-best_score[cand], kind="stable")]
    selected = []
    active = np.ones(len(cand), dtype=bool)
    for a in range(len(cand)):
        if len(selected) >= max_det:
            break
        if not active[a]:
            continue
        selected.append(cand[a])
        for b in range(a + 1, len(cand)):
            if active[b] and _iou_f32(dec[cand[a]], dec[cand[b]]) > f(iou_thr):
                active[b] = False
    n = len(selected)
    ob = np.zeros((max_det, 4), f)
    oc = np.zeros(max_det, f)
    os_ = np.zeros(max_det, f)
    ob[:n] = dec[selected]
    oc[:n] = best[selected]
    os_[:n] = best_score[selected]
    return ob, oc, os_, f(n)


def _float_set_order(values):
    """Iteration order of ``set(values)`` for small non-negative integer-valued floats on CPython 3.12
    (hash(float(k)) == k) -- tools/ssd_mobilenet.py:61 iterates classes in this order."""
    mask, table, fill = 7, [None] * 8, 0

    def place(tbl, msk, key, clean):
        perturb = key
        i = key & msk
        while True:
            n_probe = 10 if i + 9 <= msk else 1
            for j in range(n_probe):
                if tbl[i + j] is None:
                    tbl[i + j] = key
                    return True
                if not clean and tbl[i + j] == key:
                    return False
            perturb >>= 5
            i = (i * 5 + 1 + perturb) & msk

    for v in values:
        if place(table, mask, int(v), False):
            fill += 1
            if fill * 5 >= mask * 3:
                size = 8
                while size <= fill * 4:
                    size <<= 1
                new = [None] * size
                for k in table:
                    if k is not None:
                        place(new, size - 1, k, True)
                table, mask = new, size - 1
    return [k for k in table if k is not None]


def ssd_postprocess(op_boxes, op_classes, op_scores, img_w, img_h, label_names, wanted,
                    confidence=0.5, iou_threshold=0.5, score_threshold=0.5, set_order=None):
    """tools/ssd_mobilenet.py:100-150 (predict) + :59-98 (nms_boxes) + :198-213 (detect_image).
    Inputs are the TFLite op's outputs (boxes [10,4] ymin,xmin,ymax,xmax normalised; classes; scores).
    Returns (tlwh f64 [K,4], label names, scores f32)."""
    boxes = np.array(op_boxes, np.float32, copy=True)
    classes = np.array(op_classes, np.float32, copy=True)
    scores = np.array(op_scores, np.float32, copy=True)
    scores[np.reshape(np.where(np.isnan(boxes)), -1)] = 0      # :111-113 (reference quirk: flat
    scores[np.where(np.isnan(scores))] = 0                     #  row+col index list; NaN-free inputs)
    idx = np.where(scores >= confidence)
    b = boxes[idx][:, [1, 0, 3, 2]] * [img_w, img_h, img_w, img_h]   # -> f64 (:127)
    l, s = classes[idx], scores[idx]
    nb, nl, ns = [], [], []
    class_order = set(l) if set_order is None else [np.float32(k) for k in set_order(l)]
    for c in class_order:
        m = np.where(l == c)
        bb, cc, ss = b[m], l[m], s[m]
        x, y = bb[:, 0], bb[:, 1]
        w, h = bb[:, 2] - bb[:, 0], bb[:, 3] - bb[:, 1]
        areas = w * h
        order = ss.argsort()[::-1]
        keep = []
        while order.size > 0:
            i = order[0]
            keep.append(i)
            r = order[1:]
            w1 = np.maximum(0.0, np.minimum(x[i] + w[i], x[r] + w[r]) - np.maximum(x[i], x[r]) + 1)
            h1 = np.maximum(0.0, np.minimum(y[i] + h[i], y[r] + h[r]) - np.maximum(y[i], y[r]) + 1)
            inter = w1 * h1
            ovr = inter / (areas[i] + areas[r] - inter)
            order = r[np.where(ovr <= iou_threshold)[0]]
        keep = np.array(keep)
        nb.append(bb[keep]); nl.append(cc[keep]); ns.append(ss[keep])
    if not nb:
        return np.zeros((0, 4)), [], np.zeros(0, np.float32)
    bb, ll, ss = np.concatenate(nb), np.concatenate(nl).astype(np.uint), np.concatenate(ns)
    names = [label_names[int(k) + 1] for k in ll if 0 <= k < len(label_names) - 1]
    out_b, out_l, out_s = [], [], []
    for i in range(len(bb)):
        if names[i] in wanted and ss[i] >= score_threshold:
            out_b.append([bb[i][0], bb[i][1], bb[i][2] - bb[i][0], bb[i][3] - bb[i][1]])
            out_l.append(names[i]); out_s.append(ss[i])
    return np.array(out_b, dtype=float).reshape(-1, 4), out_l, np.array(out_s, np.float32)


def tflite_adapter_postprocess(op_boxes, op_classes, op_scores, count, img_w, img_h, label_list, wanted,
                               score_thr=0.5, allow=None, deny=None, max_results=-1):
    """tools/tflite_object_detector.py:234-295 (ObjectDetector._postprocess) followed by
    tools/tflite.py:26-41 (TFLITE.detect_image): the outputs of the model's detection post-process op (boxes
    (ymin,xmin,ymax,xmax) normalised, class ids, scores, count) -> (boxes [left, top, w, h] ints, label names,
    scores).  score >= threshold (:254); int() truncation of float32 products (:256-260); STABLE descending sort
    by score (:270-273); deny / allow lists (:276-288); max_results (:291-293); then the adapter keeps the
    detections whose label is wanted (tflite.py:31-40)."""
    res = []
    for i in range(int(count)):
        if op_scores[i] >= score_thr:
            y0, x0, y1, x1 = (np.float32(v) for v in op_boxes[i])
            rect = (int(x0 * np.float32(img_w)), int(y0 * np.float32(img_h)), int(x1 * np.float32(img_w)),
                    int(y1 * np.float32(img_h)))                                   # left, top, right, bottom
            res.append((rect, label_list[int(op_classes[i])], op_scores[i]))
    res = sorted(res, key=lambda d: d[2], reverse=True)
    if deny is not None:
        res = [d for d in res if d[1] not in deny]
    if allow is not None:
        res = [d for d in res if d[1] in allow]
    if max_results > 0:
        res = res[:min(len(res), max_results)]
    boxes, labels, scores = [], [], []
    for (l, t, r, b), lab, sc in res:
        if lab in wanted:
            boxes.append([l, t, r - l, b - t])
            labels.append(lab)
            scores.append(sc)
    return boxes, labels, scores


# ----------------------------------------------------------------------------- Keras YOLOv3 adapter
def _exp_f32(x):
    """float32 exp, correctly rounded by declaration (f64 exp, one rounding): numpy's own float32 exp differs between
    builds (x86 SIMD vs libm) in the last bit; the CUDA kernel uses the same definition (dd_expf)."""
    return np.exp(np.asarray(x, np.float32).astype(np.float64)).astype(np.float32)


def _sigmoid_f32(x):
    f = np.float32
    return (f(1.0) / (f(1.0) + _exp_f32(-np.asarray(x, f)))).astype(f)


def yolo3_detect(netouts, anchors, class_names, wanted, thr, image_w, image_h, net_w, net_h, nms_thresh=0.5):
    """tools/yolo.py: decode_netout (:48-81) for each output map, correct_yolo_boxes (:83-91), do_nms (:122-137),
    get_boxes (:140-153) and the tail of YOLO.detect_image (:207-237).  netouts: list of [g, g, 3*(5+C)] f32 raw maps.
    Quirks kept: every cell emits a box (the `objectness.all() <= obj_thresh` test, :63, only skips an objectness of
    exactly 0); the row used for y is the FRACTIONAL i / grid_w (:59,67); a box with two labels above the threshold is
    returned twice, both times with its arg-max label (:145-152, :210-214); the returned boxes are transposed
    (x = box[1], y = box[0], :222-225); order = reversed get_boxes order.  Scalar arithmetic follows numpy 2 (NEP 50:
    Python ints / floats do not widen float32).  Returns (boxes int64 [K,4], class indices, scores f32)."""
    f = np.float32
    xmin, ymin, xmax, ymax, classes = [], [], [], [], []
    for k, raw in enumerate(netouts):
        gh, gw = raw.shape[:2]
        o = np.array(raw, dtype=f).reshape(gh, gw, 3, -1)
        o[..., :2] = _sigmoid_f32(o[..., :2])
        o[..., 4:] = _sigmoid_f32(o[..., 4:])
        o[..., 5:] = o[..., 4][..., None] * o[..., 5:]
        o[..., 5:] *= o[..., 5:] > f(thr)
        i = np.arange(gh * gw)
        rowf = (i / gw).astype(f)                          # python float row, narrowed to float32 when it meets y (:67)
        ri, ci = (i / gw).astype(np.int64), i % gw
        cell = o[ri, ci]                                   # [gh*gw, 3, 5+C], i-major then b: the reference's append order
        for b in range(3):
            c = cell[:, b]
            keep = c[:, 4] != 0                            # :63
            x = ((ci.astype(f) + c[:, 0]) / f(gw)).astype(f)
            y = ((rowf + c[:, 1]) / f(gh)).astype(f)
            w = (f(anchors[k][2 * b]) * _exp_f32(c[:, 2]) / f(net_w)).astype(f)
            h = (f(anchors[k][2 * b + 1]) * _exp_f32(c[:, 3]) / f(net_h)).astype(f)
            cell_boxes = np.stack([x - w / f(2), y - h / f(2), x + w / f(2), y + h / f(2)], 1).astype(f)
            cell[:, b, :4] = cell_boxes
            cell[:, b, 4] = np.where(keep, f(1), f(0))
        flat = cell.reshape(-1, cell.shape[-1])            # box order: i, then b
        flat = flat[flat[:, 4] != 0]
        xmin.append(flat[:, 0]); ymin.append(flat[:, 1]); xmax.append(flat[:, 2]); ymax.append(flat[:, 3])
        classes.append(flat[:, 5:])
    xmin, ymin, xmax, ymax = (np.concatenate(a) for a in (xmin, ymin, xmax, ymax))
    cls = np.concatenate(classes).copy()
    # correct_yolo_boxes with new_w, new_h = net_w, net_h: offsets 0, scales 1 -> int(v * image size), truncation
    bx = np.stack([(xmin * f(image_w)).astype(f), (ymin * f(image_h)).astype(f), (xmax * f(image_w)).astype(f),
                   (ymax * f(image_h)).astype(f)], 1).astype(np.int64)        # astype(int64) truncates toward zero

    def overlap(a1, a2, b1, b2):
        if b1 < a1:
            return 0 if b2 < a1 else min(a2, b2) - a1
        return 0 if a2 < b1 else min(a2, b2) - b1

    def iou(p, q):
        iw = overlap(bx[p, 0], bx[p, 2], bx[q, 0], bx[q, 2])
        ih = overlap(bx[p, 1], bx[p, 3], bx[q, 1], bx[q, 3])
        inter = iw * ih
        union = (bx[p, 2] - bx[p, 0]) * (bx[p, 3] - bx[p, 1]) + (bx[q, 2] - bx[q, 0]) * (bx[q, 3] - bx[q, 1]) - inter
        return float(inter) / union

    n = len(bx)
    for c in range(cls.shape[1]):                          # do_nms: only boxes with a non-zero score can suppress
        if not np.any(cls[:, c]):
            continue
        order = np.argsort(-cls[:, c], kind="stable")
        for a in range(n):
            p = order[a]
            if cls[p, c] == 0:
                continue
            for q in order[a + 1:]:
                if iou(p, q) >= nms_thresh:
                    cls[q, c] = 0
    v = [(p, c) for p in range(n) for c in np.nonzero(cls[p] > f(thr))[0]]          # get_boxes order
    out_b, out_l, out_s = [], [], []
    for p, _ in reversed(v):
        lab = int(np.argmax(cls[p]))
        score = cls[p, lab]
        if class_names[lab] not in wanted or score < thr:
            continue
        x, y = int(bx[p, 1]), int(bx[p, 0])                # transposed, like the reference (:222-225)
        w, h = int(bx[p, 3] - bx[p, 1]), int(bx[p, 2] - bx[p, 0])
        if x < 0:
            w, x = w + x, 0
        if y < 0:
            h, y = h + y, 0
        out_b.append([x, y, w, h]); out_l.append(lab); out_s.append(score)
    return np.array(out_b, np.int64).reshape(-1, 4), np.array(out_l, np.int32), np.array(out_s, f)
